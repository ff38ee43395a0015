// conv1d as an implicit GEMM on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM), sm_100a.
//
//   D[128 rows x Nblk] (fp32, TMEM) = sum_{tap t} sum_{ci}  A_t[128 x Cin] (bf16, smem) * W_t[Cin x Nblk] (bf16, smem)
//
// Mapping (time on M, output channels on N):
//   * Activations live in HBM as planar bf16 [C/8][R][8].  One plane-slab of a row tile (128 + halo rows x 16 B)
//     is contiguous in HBM *and* is exactly one K-chunk column of the UMMA "K-major, no swizzle" canonical
//     layout (8-row core matrices of 128 contiguous bytes, SBO = 128 B, LBO = slab pitch).  So the A tile is
//     fetched by Cin/8 plain TMA bulk copies (cp.async.bulk, completion on an mbarrier) with no tensor map, and
//     every conv tap is the SAME smem tile addressed with a start offset of tap*dil rows (16 B each):
//     the halo is loaded once, taps cost no extra traffic, and no im2col is ever materialised.
//   * Weights are pre-packed (packing.py) so that one (n-block, tap, 64-channel chunk) slab is one bulk copy that
//     lands in the same canonical layout; they stream through an SB-deep mbarrier ring out of L2.
//   * Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane) + TMEM owner,
//     warps 2..5 = epilogue (tcgen05.ld -> bias/residual/leaky-relu -> bf16 -> coalesced 16 B planar stores).
//     TMEM holds two accumulator buffers so the epilogue of unit i overlaps the MMAs of unit i+1.
//   * Persistent CTAs stride over row tiles; for N > 256 (conv_pre, ConvTranspose phases) the n-blocks loop
//     inside the CTA so the A tile is fetched once.
//   * ConvTranspose1d runs as a polyphase conv: GEMM column gn = phase*Cout + co, taps = input offsets
//     {-1,0,+1}; the epilogue scatters row r, phase ph to output row up*r + ph.
#include "umma_conv.cuh"

namespace vs {
namespace {

constexpr int kTileM = 128;
constexpr int kThreads = 192;
constexpr int kMaxStages = 8;

struct Plan {
  int rows_a, halo_l, planes, KC, n_kc, Nblk, NB, SA, SB, tmem_cols, n_tiles, Cout;
  uint32_t a_bytes, b_bytes, smem_bytes;
  uint32_t off_b, off_bar;
};

struct Params {
  UmmaConv c;
  Plan p;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must abort the kernel (trap) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int tag) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("umma_conv1d: mbarrier timeout tag=%d block=%d thread=%d parity=%u\n", tag, blockIdx.x, threadIdx.x, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers bf16 inputs with fp32 accumulation.
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Shared-memory matrix descriptor, K-major, SWIZZLE_NONE (cute::UMMA::SmemDescriptor bit layout):
// [0,14) start>>4, [16,30) LBO>>4 (pitch between the two 16-byte K chunks of one MMA),
// [32,46) SBO>>4 (pitch between 8-row groups), [46,48) version = 1 on sm_100, [61,64) layout = 0.
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 = 1 @4, a/b_format BF16 = 1 @7/@10,
// a/b major K = 0 @15/@16, N>>3 @17, M>>4 @24.
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void unpack_bf16x8(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
  }
}

__global__ void __launch_bounds__(kThreads) umma_conv1d_kernel(const __grid_constant__ Params prm) {
  extern __shared__ __align__(128) uint8_t smem[];
  const UmmaConv& c = prm.c;
  const Plan& p = prm.p;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;

  const uint32_t smem_base = smem_u32(smem);
  const uint32_t a_base = smem_base;
  const uint32_t b_base = smem_base + p.off_b;
  const uint32_t bar_base = smem_base + p.off_bar;
  // barrier table (8 B each): a_full[SA] a_empty[SA] b_full[SB] b_empty[SB] acc_full[2] acc_empty[2]
  auto a_full = [&](int i) { return bar_base + 8u * i; };
  auto a_empty = [&](int i) { return bar_base + 8u * (kMaxStages + i); };
  auto b_full = [&](int i) { return bar_base + 8u * (2 * kMaxStages + i); };
  auto b_empty = [&](int i) { return bar_base + 8u * (3 * kMaxStages + i); };
  auto acc_full = [&](int i) { return bar_base + 8u * (4 * kMaxStages + i); };
  auto acc_empty = [&](int i) { return bar_base + 8u * (4 * kMaxStages + 2 + i); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + p.off_bar + 8 * (4 * kMaxStages + 4));

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.SA; ++i) { mbar_init(a_full(i), 1); mbar_init(a_empty(i), 1); }
    for (int i = 0; i < p.SB; ++i) { mbar_init(b_full(i), 1); mbar_init(b_empty(i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(acc_full(i), 1); mbar_init(acc_empty(i), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_units_per_tile = p.NB;
  const int stages_per_unit = c.taps * p.n_kc;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    uint32_t a_it = 0, b_it = 0;
    auto issue_a = [&](int tile) {
      const int sa = a_it % p.SA;
      const uint32_t ph = (a_it / p.SA) & 1;
      mbar_wait(a_empty(sa), ph ^ 1, 1);
      const int row_lo = tile * kTileM - p.halo_l, row_hi = row_lo + p.rows_a;
      const int c_lo = row_lo < 0 ? 0 : row_lo, c_hi = row_hi > c.R ? c.R : row_hi;
      const uint32_t stage = a_base + sa * p.a_bytes;
      const int n_zero_lo = c_lo - row_lo, n_zero_hi = row_hi - c_hi;
      if (n_zero_lo > 0 || n_zero_hi > 0) {   // rows outside [0,R): zero padding of the conv
        const int per_plane = n_zero_lo + n_zero_hi;
        for (int i = lane; i < p.planes * per_plane; i += 32) {
          const int pl = i / per_plane, j = i % per_plane;
          const int row = j < n_zero_lo ? j : (p.rows_a - n_zero_hi + (j - n_zero_lo));
          asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(stage + (uint32_t)(pl * p.rows_a + row) * 16u), "r"(0)
                       : "memory");
        }
        fence_proxy_async();
      }
      __syncwarp();
      const uint32_t bytes = (uint32_t)(c_hi - c_lo) * 16u;
      if (lane == 0) mbar_arrive_expect_tx(a_full(sa), bytes * p.planes);
      __syncwarp();
      for (int pl = lane; pl < p.planes; pl += 32)
        bulk_g2s(stage + (uint32_t)(pl * p.rows_a + n_zero_lo) * 16u, c.in + ((size_t)pl * c.R + c_lo) * 8, bytes, a_full(sa));
      ++a_it;
    };
    const int look = p.SA - 1;
    int next_a = blockIdx.x;
    for (int i = 0; i < look && next_a < p.n_tiles; ++i, next_a += gridDim.x) issue_a(next_a);
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      if (next_a < p.n_tiles) { issue_a(next_a); next_a += gridDim.x; }
      if (lane == 0) {
        for (int nb = 0; nb < p.NB; ++nb)
          for (int s = 0; s < stages_per_unit; ++s) {
            const int sb = b_it % p.SB;
            const uint32_t ph = (b_it / p.SB) & 1;
            mbar_wait(b_empty(sb), ph ^ 1, 2);
            mbar_arrive_expect_tx(b_full(sb), p.b_bytes);
            bulk_g2s(b_base + sb * p.b_bytes,
                     reinterpret_cast<const uint8_t*>(c.w) + ((size_t)nb * stages_per_unit + s) * p.b_bytes, p.b_bytes,
                     b_full(sb));
            ++b_it;
          }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one lane)
    if (lane == 0) {
      uint32_t a_it = 0, b_it = 0, acc_it = 0;
      const uint32_t idesc = make_idesc(p.Nblk);
      const uint32_t a_lbo = (uint32_t)p.rows_a * 16u, b_lbo = (uint32_t)p.Nblk * 16u;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const int sa = a_it % p.SA;
        mbar_wait(a_full(sa), (a_it / p.SA) & 1, 3);
        tc_fence_after();
        const uint32_t a_stage = a_base + sa * p.a_bytes;
        for (int nb = 0; nb < n_units_per_tile; ++nb) {
          const int ab = acc_it & 1;
          mbar_wait(acc_empty(ab), ((acc_it >> 1) & 1) ^ 1, 4);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(ab * p.Nblk);
          uint32_t accumulate = 0;
          for (int t = 0; t < c.taps; ++t)
            for (int kc = 0; kc < p.n_kc; ++kc) {
              const int sb = b_it % p.SB;
              mbar_wait(b_full(sb), (b_it / p.SB) & 1, 5);
              tc_fence_after();
              const uint32_t b_stage = b_base + sb * p.b_bytes;
              for (int k16 = 0; k16 < p.KC / 16; ++k16) {
                const uint32_t plane0 = (uint32_t)(kc * (p.KC / 8) + 2 * k16);
                const uint64_t ad = make_desc(a_stage + (plane0 * p.rows_a + (uint32_t)(t * c.dil)) * 16u, a_lbo, 128u);
                const uint64_t bd = make_desc(b_stage + (uint32_t)(2 * k16 * p.Nblk) * 16u, b_lbo, 128u);
                tc_mma_bf16(d_tmem, ad, bd, idesc, accumulate);
                accumulate = 1;
              }
              tc_commit(b_empty(sb));      // frees the weight slot once these MMAs have read it
              ++b_it;
            }
          tc_commit(acc_full(ab));         // accumulator complete -> epilogue
          ++acc_it;
        }
        tc_commit(a_empty(sa));            // every MMA of the tile has read the A stage
        ++a_it;
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int q = warp & 3;                // TMEM lane quarter this warp may access
    uint32_t acc_it = 0;
    const int R_out = c.R * c.up;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      const int r = tile * kTileM + q * 32 + lane;
      for (int nb = 0; nb < n_units_per_tile; ++nb) {
        const int ab = acc_it & 1;
        mbar_wait(acc_full(ab), (acc_it >> 1) & 1, 6);
        tc_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * p.Nblk);
        for (int col0 = 0; col0 < p.Nblk; col0 += 32) {
          uint32_t v[32];
          tmem_ld32(t_row + (uint32_t)col0, v);
          if (r < c.R) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int gn0 = nb * p.Nblk + col0 + 8 * g;
              const int phs = gn0 / p.Cout, co0 = gn0 - phs * p.Cout;
              const int orow = c.up * r + phs;
              const int utt = c.row_utt ? c.row_utt[orow / c.row_div] : 0;
              const size_t o = ((size_t)(co0 >> 3) * R_out + orow) * 8;
              uint4 raw = make_uint4(0, 0, 0, 0), act = make_uint4(0, 0, 0, 0);
              if (utt >= 0) {
                float y[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) y[e] = __uint_as_float(v[8 * g + e]);
                if (c.bias) {
#pragma unroll
                  for (int e = 0; e < 8; ++e) y[e] += __ldg(c.bias + co0 + e);
                }
                if (c.ubias) {
                  const float* ub = c.ubias + (size_t)(c.ubias_idx ? c.ubias_idx[utt] : utt) * c.N + gn0;
#pragma unroll
                  for (int e = 0; e < 8; ++e) y[e] += __ldg(ub + e);
                }
                if (c.res) {
                  float f[8];
                  unpack_bf16x8(*reinterpret_cast<const uint4*>(c.res + o), f);
#pragma unroll
                  for (int e = 0; e < 8; ++e) y[e] += f[e];
                }
                if (c.res2) {
                  float f[8];
                  unpack_bf16x8(*reinterpret_cast<const uint4*>(c.res2 + o), f);
#pragma unroll
                  for (int e = 0; e < 8; ++e) y[e] += f[e];
                }
                raw = make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]),
                                 pack_bf16x2(y[6], y[7]));
                float z[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) z[e] = lrelu(y[e] * c.act_scale, c.act_slope);
                act = make_uint4(pack_bf16x2(z[0], z[1]), pack_bf16x2(z[2], z[3]), pack_bf16x2(z[4], z[5]),
                                 pack_bf16x2(z[6], z[7]));
              }
              if (c.out_raw) *reinterpret_cast<uint4*>(c.out_raw + o) = raw;
              if (c.out_act) *reinterpret_cast<uint4*>(c.out_act + o) = act;
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty(ab));
        ++acc_it;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

int make_plan(const UmmaConv& c, Plan* out) {
  Plan p{};
  VS_REQUIRE(c.Cin % 16 == 0 && c.Cin >= 16, "umma_conv1d: Cin=%d must be a multiple of 16", c.Cin);
  VS_REQUIRE(c.N % 32 == 0, "umma_conv1d: N=%d must be a multiple of 32", c.N);
  VS_REQUIRE(c.up >= 1 && c.N % c.up == 0 && (c.N / c.up) % 8 == 0, "umma_conv1d: bad upsample factor");
  VS_REQUIRE(c.R > 0 && c.taps >= 1 && c.dil >= 1 && c.pad_l >= 0, "umma_conv1d: bad shape");
  p.Cout = c.N / c.up;
  p.Nblk = c.N < 256 ? c.N : 256;
  VS_REQUIRE(c.N % p.Nblk == 0, "umma_conv1d: N=%d not a multiple of the 256-column block", c.N);
  p.NB = c.N / p.Nblk;
  p.KC = c.Cin < 64 ? c.Cin : 64;
  VS_REQUIRE(c.Cin % p.KC == 0, "umma_conv1d: Cin=%d not a multiple of the %d-channel chunk", c.Cin, p.KC);
  p.n_kc = c.Cin / p.KC;
  p.planes = c.Cin / 8;
  p.halo_l = c.pad_l * c.dil;
  p.rows_a = kTileM + (c.taps - 1) * c.dil;
  p.a_bytes = (uint32_t)p.planes * p.rows_a * 16u;
  p.b_bytes = (uint32_t)p.KC * p.Nblk * 2u;
  VS_REQUIRE(p.rows_a * 16 < (1 << 18), "umma_conv1d: halo too large for the descriptor pitch");
  const uint32_t bar_bytes = 8u * (4 * kMaxStages + 4) + 16u;
  const uint32_t budget_small = 96u * 1024, budget_max = 220u * 1024;
  p.SA = (2 * p.a_bytes + 2 * p.b_bytes + bar_bytes <= budget_max) ? 2 : 1;
  VS_REQUIRE(p.SA * p.a_bytes + 2 * p.b_bytes + bar_bytes <= budget_max, "umma_conv1d: tile does not fit in shared memory");
  uint32_t budget = budget_small;
  if (p.SA * p.a_bytes + 3 * p.b_bytes + bar_bytes > budget) budget = budget_max;
  int sb = (int)((budget - bar_bytes - p.SA * p.a_bytes) / p.b_bytes);
  p.SB = sb > kMaxStages ? kMaxStages : sb;
  if (p.SB < 2) p.SB = 2;
  int cols = 32;
  while (cols < 2 * p.Nblk) cols *= 2;
  p.tmem_cols = cols;
  p.off_b = p.SA * p.a_bytes;
  p.off_bar = (p.off_b + p.SB * p.b_bytes + 127u) & ~127u;
  p.smem_bytes = p.off_bar + bar_bytes;
  if (p.tmem_cols > 256 && p.smem_bytes < 120u * 1024) p.smem_bytes = 120u * 1024;   // one CTA per SM: TMEM has 512 columns
  p.n_tiles = (c.R + kTileM - 1) / kTileM;
  *out = p;
  return VS_OK;
}

}  // namespace

int umma_conv1d(const UmmaConv& c, cudaStream_t st) {
  Params prm;
  prm.c = c;
  VS_TRY(make_plan(c, &prm.p));
  VS_REQUIRE(c.in && c.w && (c.out_raw || c.out_act), "umma_conv1d: null pointer");
  static int n_sm = 0;
  static bool configured = false;
  if (!configured) {
    int dev = 0;
    VS_CUDA_CHECK(cudaGetDevice(&dev));
    VS_CUDA_CHECK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    VS_CUDA_CHECK(cudaFuncSetAttribute(umma_conv1d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  int per_sm = (int)((227u * 1024) / (prm.p.smem_bytes + 1024));
  const int tmem_limit = 512 / prm.p.tmem_cols;
  if (per_sm > tmem_limit) per_sm = tmem_limit;
  if (per_sm > 4) per_sm = 4;
  if (per_sm < 1) per_sm = 1;
  int grid = n_sm * per_sm;
  if (grid > prm.p.n_tiles) grid = prm.p.n_tiles;
  umma_conv1d_kernel<<<grid, kThreads, prm.p.smem_bytes, st>>>(prm);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

}  // namespace vs
