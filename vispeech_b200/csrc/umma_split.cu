// fp32-accurate GEMM-shaped conv1d over ragged rows on tcgen05 kind::f16: every operand is an fp16 hi + lo pair
// (x = hi + lo to 22 bits) and D += a_hi w_hi + a_lo w_hi + a_hi w_lo (the dropped a_lo w_lo term is 2^-22 relative).
//
// This is the latent stages' workhorse (encoder QKV / O / FFN convs of the text encoder, pitch predictor and frame prior
// network, predictor convs, projection): they need fp32-level accuracy (the prior sampling amplifies their error)
// and used to run as 3xTF32 on umma_tf32.cu, whose four loader warps re-staged row-major fp32 rows for every n-block (80 us
// for 17 us of MMAs on the QKV conv).  Here, as in the decoder's conv (umma_conv.cu):
//   * activations are PLANAR fp16 [C/8][R][8] tensors (a hi and a lo one): a plane-slab of a row tile is contiguous in HBM and
//     is one K-chunk column of the UMMA K-major no-swizzle layout, so the A tile arrives by plain TMA bulk copies and a conv
//     tap is the same tile at a row offset;
//   * the three terms accumulate into the SAME accumulator; per 64-channel K-chunk the weight slabs stream as [w_hi, w_lo]
//     (packing.py pack_split16): the w_hi slab multiplies the chunk's hi AND lo planes while it is resident, the w_lo slab its
//     hi planes, so two slabs stream per twelve MMAs; the A tile holds hi and lo once;
//   * half the tensor time of 3xTF32 (kind::f16 runs at twice the kind::tf32 rate) and half the operand bytes.
// K-slices: a conv with Cin = 768 (FFN conv_2) runs as four independent K-slices of 192 channels, each a unit of its own
// writing its own fp32 partial (the consumer - LayerNorm - sums them in a fixed order: deterministic, and four times the
// parallelism where a call has only ~20 row tiles).  With few row tiles the n-blocks are spread over CTAs as well.
// Outputs: fp32 row-major [R][ld] (bias, ReLU, validity mask) and / or planar fp16 hi / lo (the next conv's operand).
#include "umma_conv.cuh"
#include "umma_common.cuh"

namespace vs {
namespace {

using namespace umma;
constexpr int kEpiWarps = 8;                 // two per TMEM lane quarter, alternating 32-column chunks
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kMaxSB = 8;

struct Plan {
  int cs, P, KC, nkc1, n_kc, Nblk, NB, nbo, nb_per_unit, n_tiles, n_units, rows_a, halo_l, SA, SB, NACC, tmem_cols;
  uint32_t a_bytes, b_bytes, smem_bytes, off_b, off_bar, off_bias;
};
struct Params {
  UmmaSplit c;
  Plan p;
};

__global__ void __launch_bounds__(kThreads, 1) umma_split_kernel(const __grid_constant__ Params prm) {
  extern __shared__ __align__(128) uint8_t smem[];
  pdl_trigger();
  const UmmaSplit& c = prm.c;
  const Plan& p = prm.p;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t a_base = smem_base, b_base = smem_base + p.off_b, bar_base = smem_base + p.off_bar;
  // barriers: a_full[2] a_empty[2] b_full[8] b_empty[8] acc_full[4] acc_empty[4]
  auto a_full = [&](int i) { return bar_base + 8u * i; };
  auto a_empty = [&](int i) { return bar_base + 8u * (2 + i); };
  auto b_full = [&](int i) { return bar_base + 8u * (4 + i); };
  auto b_empty = [&](int i) { return bar_base + 8u * (4 + kMaxSB + i); };
  auto acc_full = [&](int i) { return bar_base + 8u * (4 + 2 * kMaxSB + i); };
  auto acc_empty = [&](int i) { return bar_base + 8u * (8 + 2 * kMaxSB + i); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + p.off_bar + 8 * (12 + 2 * kMaxSB));

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(a_full(i), 1); mbar_init(a_empty(i), 1); }
    for (int i = 0; i < kMaxSB; ++i) { mbar_init(b_full(i), 1); mbar_init(b_empty(i), 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(acc_full(i), 1); mbar_init(acc_empty(i), kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  float* bias_s = reinterpret_cast<float*>(smem + p.off_bias);     // [N] (zeros when the conv has no bias)
  for (int i = threadIdx.x; i < c.N; i += kThreads) bias_s[i] = c.bias ? c.bias[i] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();            // everything above (barriers, TMEM, the bias = a weight) overlapped the previous kernel's tail
  const uint32_t tmem_base = *tmem_slot;
  const int slabs_per_nb = c.taps * p.n_kc;
  // unit u -> (row tile, n-block group, K-slice); consecutive units are consecutive tiles
  auto unit_of = [&](int u, int& tile, int& nb0, int& slice) {
    tile = u % p.n_tiles;
    const int rest = u / p.n_tiles;
    nb0 = (rest % p.nbo) * p.nb_per_unit;
    slice = rest / p.nbo;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    uint32_t a_it = 0, b_it = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      int tile, nb0, slice;
      unit_of(u, tile, nb0, slice);
      {
        const int sa = a_it % p.SA;
        const uint32_t ph = (a_it / p.SA) & 1;
        mbar_wait(a_empty(sa), ph ^ 1, 1);
        const int row_lo = tile * kTileM - p.halo_l, row_hi = row_lo + p.rows_a;
        const int c_lo = row_lo < 0 ? 0 : row_lo, c_hi = row_hi > c.R ? c.R : row_hi;
        const uint32_t stage = a_base + sa * p.a_bytes;
        const int n_zero_lo = c_lo - row_lo, n_zero_hi = row_hi - c_hi;
        if (n_zero_lo > 0 || n_zero_hi > 0) {   // rows outside [0,R): zero padding of the conv
          const int per_plane = n_zero_lo + n_zero_hi;
          for (int i = lane; i < 2 * p.P * per_plane; i += 32) {
            const int pl = i / per_plane, j = i % per_plane;
            const int row = j < n_zero_lo ? j : (p.rows_a - n_zero_hi + (j - n_zero_lo));
            asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(stage + (uint32_t)(pl * p.rows_a + row) * 16u), "r"(0)
                         : "memory");
          }
          fence_proxy_async();
        }
        __syncwarp();
        const uint32_t bytes = (uint32_t)(c_hi - c_lo) * 16u;
        if (lane == 0) mbar_arrive_expect_tx(a_full(sa), bytes * 2u * (uint32_t)p.P);
        __syncwarp();
        for (int pl = lane; pl < 2 * p.P; pl += 32) {       // hi planes of this K-slice, then its lo planes
          const __half* src = (pl < p.P ? c.in_hi + ((size_t)(slice * p.P + pl) * c.R + c_lo) * 8
                                        : c.in_lo + ((size_t)(slice * p.P + pl - p.P) * c.R + c_lo) * 8);
          bulk_g2s(stage + (uint32_t)(pl * p.rows_a + n_zero_lo) * 16u, src, bytes, a_full(sa));
        }
        ++a_it;
      }
      if (lane == 0) {
        const uint8_t* w0 = reinterpret_cast<const uint8_t*>(c.w) + (size_t)(slice * p.NB + nb0) * slabs_per_nb * p.b_bytes;
        for (int sl = 0; sl < p.nb_per_unit * slabs_per_nb; ++sl) {
          const int sb = b_it % p.SB;
          const uint32_t ph = (b_it / p.SB) & 1;
          mbar_wait(b_empty(sb), ph ^ 1, 2);
          mbar_arrive_expect_tx(b_full(sb), p.b_bytes);
          bulk_g2s(b_base + sb * p.b_bytes, w0 + (size_t)sl * p.b_bytes, p.b_bytes, b_full(sb));
          ++b_it;
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (warp-uniform loop, elected lane issues)
    const uint32_t idesc = make_idesc(p.Nblk);
    const uint32_t a_lbo = (uint32_t)p.rows_a * 16u, b_lbo = (uint32_t)p.Nblk * 16u;
    const uint32_t a_hi = (uint32_t)(make_desc(0, a_lbo, 128u) >> 32), b_hi = (uint32_t)(make_desc(0, b_lbo, 128u) >> 32);
    const uint32_t a_lo_fixed = (uint32_t)make_desc(0, a_lbo, 128u), b_lo_fixed = (uint32_t)make_desc(0, b_lbo, 128u);
    const uint32_t a_kstep = 2u * (uint32_t)p.rows_a, b_kstep = 2u * (uint32_t)p.Nblk;     // two planes per K = 16 step
    const int taps = c.taps, dil = c.dil, n_kc = p.n_kc, nkc1 = p.nkc1, k16s = p.KC / 16;
    const uint32_t slab16 = p.b_bytes >> 4, b_base16 = b_base >> 4, chunk16 = (uint32_t)(p.KC / 8) * (uint32_t)p.rows_a;
    uint32_t acc_slot = 0, acc_phase = 0, b_slot = 0, b_phase = 0, a_slot = 0, a_phase = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      mbar_wait(a_full(a_slot), a_phase, 3);
      tc_fence_after();
      const uint32_t a_stage16 = (a_base + a_slot * p.a_bytes) >> 4;
      for (int nb = 0; nb < p.nb_per_unit; ++nb) {
        mbar_wait(acc_empty(acc_slot), acc_phase ^ 1, 4);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc_slot * (uint32_t)p.Nblk;
        uint32_t accumulate = 0;
        for (int t = 0; t < taps; ++t) {
          const uint32_t a_tap = a_lo_fixed + a_stage16 + (uint32_t)(t * dil);
          for (int kc = 0; kc < n_kc; ++kc) {
            // slabs alternate w_hi(kk), w_lo(kk) over the K-chunks kk: the w_hi slab serves BOTH a_hi and a_lo while it is
            // resident (two of the three terms), the w_lo slab a_hi - two slabs stream per 12 MMAs instead of three
            const int kk = kc >> 1;
            const bool w_is_lo = kc & 1;
            mbar_wait(b_full(b_slot), b_phase, 5);
            tc_fence_after();
            for (int pass = 0; pass < (w_is_lo ? 1 : 2); ++pass) {               // pass 0: a_hi planes, pass 1: a_lo planes
              uint32_t a_lo = a_tap + (uint32_t)((pass ? nkc1 : 0) + kk) * chunk16;
              uint32_t b_lo = b_lo_fixed + b_base16 + b_slot * slab16;
#pragma unroll 4
              for (int k16 = 0; k16 < k16s; ++k16) {
                tc_mma_f16_lohi(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, accumulate);
                accumulate = 1;
                a_lo += a_kstep;
                b_lo += b_kstep;
              }
            }
            tc_commit(b_empty(b_slot));
            if (++b_slot == (uint32_t)p.SB) { b_slot = 0; b_phase ^= 1; }
          }
        }
        tc_commit(acc_full(acc_slot));
        if (++acc_slot == (uint32_t)p.NACC) { acc_slot = 0; acc_phase ^= 1; }
      }
      tc_commit(a_empty(a_slot));
      if (++a_slot == (uint32_t)p.SA) { a_slot = 0; a_phase ^= 1; }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9)
    const int q = warp & 3, hsel = (warp - 2) >> 2;
    const int n_chunks = p.Nblk / 32;
    const bool relu = c.act == 1;
    uint32_t acc_slot = 0, acc_phase = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      int tile, nb0, slice;
      unit_of(u, tile, nb0, slice);
      const int r = tile * kTileM + q * 32 + lane;
      const bool in_range = r < c.R;
      const bool valid = in_range && (!c.row_utt || c.row_utt[r] >= 0);
      float* o32 = c.out32 ? c.out32 + (size_t)slice * c.out32_slice + (size_t)r * c.out32_ld : nullptr;
      for (int nb = 0; nb < p.nb_per_unit; ++nb) {
        mbar_wait(acc_full(acc_slot), acc_phase, 6);
        tc_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc_slot * (uint32_t)p.Nblk;
        for (int cc = hsel; cc < n_chunks; cc += 2) {
          const int col0 = (nb0 + nb) * p.Nblk + cc * 32;                  // first GEMM column of this chunk
          uint32_t v[32];
          tmem_ld32(t_row + (uint32_t)(cc * 32), v);
          if (!in_range) continue;
          float y[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            float t = __uint_as_float(v[e]) + (slice == 0 ? bias_s[col0 + e] : 0.f);
            if (relu) t = fmaxf(t, 0.f);
            y[e] = valid ? t : 0.f;
          }
          if (o32) {
#pragma unroll
            for (int e = 0; e < 8; ++e)
              *reinterpret_cast<float4*>(o32 + col0 + 4 * e) = make_float4(y[4 * e], y[4 * e + 1], y[4 * e + 2], y[4 * e + 3]);
          }
          if (c.out_hi) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint32_t hw[4], lw[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                hw[e] = pack_f16x2(y[8 * g + 2 * e], y[8 * g + 2 * e + 1]);
                const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hw[e]));
                lw[e] = pack_f16x2(y[8 * g + 2 * e] - hf.x, y[8 * g + 2 * e + 1] - hf.y);
              }
              const size_t o = ((size_t)(col0 / 8 + g) * c.R + r) * 8;
              *reinterpret_cast<uint4*>(c.out_hi + o) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
              *reinterpret_cast<uint4*>(c.out_lo + o) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty(acc_slot));
        if (++acc_slot == (uint32_t)p.NACC) { acc_slot = 0; acc_phase ^= 1; }
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

// fp32 row-major [R][ld] -> planar fp16 hi / lo [C/8][R][8] (x = hi + lo to 22 bits); up to n_sum inputs are summed first
// (K-slice partials), zeros on invalid rows.  One thread per (plane, row): 16-byte stores, coalesced over rows.
__global__ void rows_to_split_kernel(const float* __restrict__ x, int ld, int64_t x_stride, int n_sum, const int32_t* __restrict__ row_utt,
                                     __half* __restrict__ hi, __half* __restrict__ lo, int R, int C) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (C / 8) * R) return;
  const int pl = i / R, r = i % R;
  float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (!row_utt || row_utt[r] >= 0) {
    for (int s = 0; s < n_sum; ++s) {
      const float* src = x + (size_t)s * x_stride + (size_t)r * ld + pl * 8;
      const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
      f[0] += a.x; f[1] += a.y; f[2] += a.z; f[3] += a.w; f[4] += b.x; f[5] += b.y; f[6] += b.z; f[7] += b.w;
    }
  }
  uint32_t hw[4], lw[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    hw[e] = pack_f16x2(f[2 * e], f[2 * e + 1]);
    const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hw[e]));
    lw[e] = pack_f16x2(f[2 * e] - hf.x, f[2 * e + 1] - hf.y);
  }
  *reinterpret_cast<uint4*>(hi + (size_t)i * 8) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
  *reinterpret_cast<uint4*>(lo + (size_t)i * 8) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
}

int make_plan(const UmmaSplit& c, Plan* out) {
  Plan p{};
  VS_REQUIRE(c.k_slices >= 1 && c.Cin % c.k_slices == 0, "umma_split: Cin=%d is not divisible into %d K-slices", c.Cin, c.k_slices);
  p.cs = c.Cin / c.k_slices;
  VS_REQUIRE(p.cs % 32 == 0 && p.cs <= 192, "umma_split: K-slice of %d channels (multiple of 32, <= 192)", p.cs);
  p.KC = p.cs % 64 == 0 ? 64 : 32;
  p.nkc1 = p.cs / p.KC;
  p.n_kc = 2 * p.nkc1;                          // weight slabs per tap: [w_hi(kk), w_lo(kk)] for every K-chunk kk
  p.P = p.cs / 8;
  VS_REQUIRE(c.N % 32 == 0 && c.N >= 32, "umma_split: N=%d must be a multiple of 32", c.N);
  p.Nblk = 0;
  for (int nb = 1; nb <= 16 && !p.Nblk; ++nb)
    if (c.N % nb == 0 && c.N / nb <= 256 && (c.N / nb) % 32 == 0) p.Nblk = c.N / nb;
  VS_REQUIRE(p.Nblk > 0, "umma_split: N=%d has no n-block of <= 256 columns", c.N);
  p.NB = c.N / p.Nblk;
  VS_REQUIRE(c.R > 0 && c.taps >= 1 && c.dil >= 1 && c.pad_l >= 0, "umma_split: bad shape");
  p.halo_l = c.pad_l * c.dil;
  p.rows_a = kTileM + (c.taps - 1) * c.dil;
  p.a_bytes = (uint32_t)(2 * p.P) * (uint32_t)p.rows_a * 16u;
  p.b_bytes = (uint32_t)p.KC * (uint32_t)p.Nblk * 2u;
  p.n_tiles = (c.R + kTileM - 1) / kTileM;
  // few row tiles: spread the n-blocks over CTAs too (each unit then fetches its own copy of the A tile)
  p.nbo = (p.n_tiles * c.k_slices < 74 && p.NB > 1) ? p.NB : 1;
  p.nb_per_unit = p.NB / p.nbo;
  p.n_units = p.n_tiles * p.nbo * c.k_slices;
  const uint32_t bar_bytes = 8u * (12 + 2 * kMaxSB) + 16u;
  const uint32_t fixed = bar_bytes + 128u + (uint32_t)c.N * 4u;
  const uint32_t cap = 220u * 1024;
  p.SA = (2 * p.a_bytes + 4 * p.b_bytes + fixed <= cap) ? 2 : 1;
  VS_REQUIRE(p.SA * p.a_bytes + 2 * p.b_bytes + fixed <= cap, "umma_split: tile does not fit in shared memory");
  int sb = (int)((cap - fixed - p.SA * p.a_bytes) / p.b_bytes);
  p.SB = sb > kMaxSB ? kMaxSB : sb;
  p.off_b = p.SA * p.a_bytes;
  p.off_bar = (p.off_b + p.SB * p.b_bytes + 127u) & ~127u;
  p.off_bias = (p.off_bar + bar_bytes + 15u) & ~15u;
  p.smem_bytes = p.off_bias + (uint32_t)c.N * 4u;
  int nacc = 512 / p.Nblk;
  p.NACC = nacc > 4 ? 4 : nacc;
  int cols = 32;
  while (cols < p.NACC * p.Nblk) cols *= 2;
  p.tmem_cols = cols;
  // one CTA per SM: request enough shared memory that two can never co-reside (their TMEM allocations could not both succeed
  // next to a decoder kernel's on the same SM)
  if (p.smem_bytes < 116u * 1024) p.smem_bytes = 116u * 1024;
  *out = p;
  return VS_OK;
}

}  // namespace

int umma_split(const UmmaSplit& c, cudaStream_t st) {
  Params prm;
  prm.c = c;
  VS_REQUIRE(c.in_hi && c.in_lo && c.w && (c.out32 || (c.out_hi && c.out_lo)), "umma_split: null pointer");
  VS_REQUIRE(c.k_slices == 1 || (c.out32 && !c.out_hi), "umma_split: K-slices write fp32 partials only");
  VS_REQUIRE(!c.out32 || (c.out32_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(c.out32) & 15) == 0), "umma_split: out32 must be 16-byte aligned");
  VS_TRY(make_plan(c, &prm.p));
  int n_sm = 0;
  VS_TRY(device_sm_count(&n_sm));
  VS_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(umma_split_kernel), 227 * 1024));
  const int grid = prm.p.n_units < n_sm ? prm.p.n_units : n_sm;
  VS_CUDA_CHECK(launch_pdl(umma_split_kernel, dim3(grid), dim3(kThreads), prm.p.smem_bytes, st, prm));
  VS_LAUNCH_CHECK();
  return VS_OK;
}

__global__ void sum_partials_kernel(const float* __restrict__ x, int64_t stride, int n, float* __restrict__ out, int64_t total4) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  float4 a = reinterpret_cast<const float4*>(x)[i];
  for (int s = 1; s < n; ++s) {                       // fixed order: deterministic
    const float4 b = reinterpret_cast<const float4*>(x + (size_t)s * stride)[i];
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
  }
  reinterpret_cast<float4*>(out)[i] = a;
}
int sum_partials(const float* x, int64_t stride, int n, float* out, int64_t total, cudaStream_t st) {
  VS_REQUIRE(total % 4 == 0 && stride % 4 == 0, "sum_partials: sizes must be multiples of 4");
  sum_partials_kernel<<<(unsigned)((total / 4 + 255) / 256), 256, 0, st>>>(x, stride, n, out, total / 4);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

int rows_to_split(const float* x, int ld, int64_t x_stride, int n_sum, const int32_t* row_utt, __half* hi, __half* lo, int R, int C,
                  cudaStream_t st) {
  VS_REQUIRE(x && hi && lo && R > 0 && C % 8 == 0 && ld % 4 == 0 && n_sum >= 1, "rows_to_split: bad arguments");
  VS_CUDA_CHECK(launch_pdl(rows_to_split_kernel, dim3(((C / 8) * R + 255) / 256), dim3(256), 0, st, x, ld, x_stride, n_sum, row_utt, hi, lo, R, C));
  VS_LAUNCH_CHECK();
  return VS_OK;
}

}  // namespace vs
