// A whole ResBlock1 (reference modules.py:210-223: three iterations x = x + c2(lrelu(c1(lrelu(x)))), dilations 1, 3, 5) of the
// C = 64 decoder stage for the k = 3 branch as ONE tcgen05 kernel - the design of the last stage's kernel (umma_mrf.cu) applied
// one stage up, where it fits:
//   * the residual stream x lives in TMEM as the fp32 accumulator of c2 itself (the residual add is free and never rounded);
//     only the conv OPERANDS lrelu(x), lrelu(c1) are fp16, in shared memory, written by the epilogues in the UMMA layout;
//   * one tensor is read (a = lrelu(x0), the stage's activated stream) and one written (the ResBlock's output) instead of the
//     three conv-pair launches' six (5.3 GB -> 1.8 GB of DRAM traffic at the C2 size, 1.45 -> ~0.9 ms).
// Only k = 3 fits: a ResBlock's receptive field is its halo (12 rows at k = 3 -> 16 of a 512-row super tile, 6 % recompute; 36
// rows at k = 7 would need super tiles that neither TMEM - x for every tile plus the c1 ring - nor shared memory - both operand
// tiles plus a weight ring of 57 KB convs - can hold).
// Tiling: super tile of S = 4 row tiles (512 rows, 480 written).  TMEM (512 columns): X[t] 64 columns per tile (256) and a ring of
// 4 accumulators for c1 (256).  Roles (608 threads): warp 0 streams the six convs' 24 KB weight sets through a 3-slot ring; warps
// 1, 2 issue c1 / c2; four tile crews of four warps run both epilogues of their tile (see umma_mrf.cu for the protocol).
#include "umma_conv.cuh"
#include "umma_common.cuh"

namespace vs {
namespace {

using namespace umma;

constexpr int kC = 64;
constexpr int kS = 4;                               // row tiles per super tile
constexpr int kRows = kS * kTileM;                  // 512
constexpr int kHalo = 16;                           // >= (1 + 1) + (3 + 1) + (5 + 1) rows of boundary contamination per side
constexpr int kValid = kRows - 2 * kHalo;           // 480 rows written per super tile
constexpr int kPadA = 8, kPadM = 8;                 // zero rows either side of A (>= 5: d = 5) and MID (>= 1)
constexpr int kRowsA = kRows + 2 * kPadA, kRowsM = kRows + 2 * kPadM;
constexpr int kPlanes = kC / 8;
constexpr int kTaps = 3;
constexpr int kRing = (512 - kC * kS) / kC;         // 4 c1 accumulators
constexpr int kWSlots = 3;
constexpr uint32_t kTapBytes = kC * kC * 2;         // 8 KB: one tap's [K = 64][N = 64] slab
constexpr uint32_t kWSlotBytes = kTaps * kTapBytes; // one conv
constexpr int kIssuers = 2;
constexpr int kCrewWarps = 4 * kS;
constexpr int kThreads = (1 + kIssuers + kCrewWarps) * 32;

constexpr uint32_t kOffA = 0;
constexpr uint32_t kOffM = kOffA + kPlanes * kRowsA * 16;
constexpr uint32_t kOffW = kOffM + kPlanes * kRowsM * 16;
constexpr uint32_t kOffBar = kOffW + kWSlots * kWSlotBytes;
constexpr int kNumBars = 2 * kWSlots + kRing + 4 * kS;
constexpr uint32_t kSmemBytes = kOffBar + 8 * kNumBars + 16;
static_assert(kSmemBytes <= 227 * 1024, "umma_resblock: shared memory");
static_assert(kRing == kS && kThreads <= 1024, "umma_resblock: one c1 accumulator per tile (the issuer relies on it)");

struct Params {
  UmmaResBlock c;
  int row_div_shift, n_super;
  float b1[3][kC];             // c1 biases [iteration][channel]
  float bcum[3][kC];           // cumulative c2 biases: x_m = X[t] + bcum[m] after iteration m
};
static_assert(sizeof(Params) <= 4000, "umma_resblock: kernel parameter block");

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_ld32_nw(uint32_t taddr, uint32_t* v) {
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
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1) umma_resblock_kernel(const __grid_constant__ Params prm) {
  extern __shared__ __align__(128) uint8_t smem[];
  pdl_trigger();
  const UmmaResBlock& c = prm.c;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t a_base = smem_base + kOffA, m_base = smem_base + kOffM, w_base = smem_base + kOffW;
  const uint32_t bar = smem_base + kOffBar;
  auto w_full = [&](uint32_t i) { return bar + 8u * i; };
  auto w_empty = [&](uint32_t i) { return bar + 8u * (kWSlots + i); };
  auto acc1_empty = [&](uint32_t i) { return bar + 8u * (2 * kWSlots + i); };
  auto acc1_full = [&](uint32_t t) { return bar + 8u * (2 * kWSlots + kRing + t); };
  auto a_ready = [&](uint32_t t) { return bar + 8u * (2 * kWSlots + kRing + kS + t); };
  auto mid_ready = [&](uint32_t t) { return bar + 8u * (2 * kWSlots + kRing + 2 * kS + t); };
  auto x_full = [&](uint32_t t) { return bar + 8u * (2 * kWSlots + kRing + 3 * kS + t); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + kOffBar + 8 * kNumBars);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kWSlots; ++i) { mbar_init(w_full(i), 1); mbar_init(w_empty(i), 1); }
    for (int i = 0; i < kRing; ++i) mbar_init(acc1_empty(i), 4);
    for (int t = 0; t < kS; ++t) { mbar_init(acc1_full(t), 1); mbar_init(a_ready(t), 4); mbar_init(mid_ready(t), 4); mbar_init(x_full(t), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // the pad rows of A and MID stay zero for the whole kernel: "outside the super tile" is zero padding
  for (int i = threadIdx.x; i < kPlanes * 2 * kPadA; i += kThreads) {
    const int pl = i / (2 * kPadA), r = i % (2 * kPadA);
    const int row = r < kPadA ? r : kRowsA - 2 * kPadA + r;
    sts128(a_base + (uint32_t)(pl * kRowsA + row) * 16u, 0u, 0u, 0u, 0u);
  }
  for (int i = threadIdx.x; i < kPlanes * 2 * kPadM; i += kThreads) {
    const int pl = i / (2 * kPadM), r = i % (2 * kPadM);
    const int row = r < kPadM ? r : kRowsM - 2 * kPadM + r;
    sts128(m_base + (uint32_t)(pl * kRowsM + row) * 16u, 0u, 0u, 0u, 0u);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();                              // the prologue above overlapped the previous kernel's tail
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_x = tmem_base, tm_ring = tmem_base + kC * kS;
  const int R = c.R;

  if (warp == 0) {
    // ------------------------------------------------------------------ weight producer: c1_0, c2_0, c1_1, c2_1, c1_2, c2_2
    if (lane == 0) {
      uint32_t wi = 0;
      for (int u = blockIdx.x; u < prm.n_super; u += gridDim.x)
        for (int q = 0; q < 6; ++q, ++wi) {
          const uint32_t slot = wi % kWSlots, ph = (wi / kWSlots) & 1u;
          mbar_wait(w_empty(slot), ph ^ 1u, 81);
          mbar_arrive_expect_tx(w_full(slot), kWSlotBytes);
          bulk_g2s(w_base + slot * kWSlotBytes, c.w[q >> 1][q & 1], kWSlotBytes, w_full(slot));
        }
    }
  } else if (warp <= kIssuers) {
    // ------------------------------------------------------------------ MMA issuers: warp 1 every c1 (A -> ring slot), warp 2
    // every c2 (MID -> X[t], accumulating onto the residual stream)
    const bool is_c1 = warp == 1;
    const uint32_t idesc = make_idesc(kC);
    constexpr uint32_t b_lbo = (uint32_t)kC * 16u, b_kstep = 2u * kC;
    const uint32_t b_hi = (uint32_t)(make_desc(0, b_lbo, 128u) >> 32), b_lo_fixed = (uint32_t)make_desc(0, b_lbo, 128u);
    constexpr uint32_t a1_lbo = (uint32_t)kRowsA * 16u, a2_lbo = (uint32_t)kRowsM * 16u;
    const uint32_t a1_hi = (uint32_t)(make_desc(0, a1_lbo, 128u) >> 32), a1_lo_fixed = (uint32_t)make_desc(0, a1_lbo, 128u);
    const uint32_t a2_hi = (uint32_t)(make_desc(0, a2_lbo, 128u) >> 32), a2_lo_fixed = (uint32_t)make_desc(0, a2_lbo, 128u);
    constexpr uint32_t a1_kstep = 2u * kRowsA, a2_kstep = 2u * kRowsM;
    uint32_t wi = is_c1 ? 0u : 1u, gen = 0;
    for (int u = blockIdx.x; u < prm.n_super; u += gridDim.x)
      for (int m = 0; m < 3; ++m, wi += 2, ++gen) {
        const int dil = 2 * m + 1;
        const uint32_t ws = wi % kWSlots, wp = (wi / kWSlots) & 1u;
        const uint32_t pg = gen & 1u;
        const uint32_t w_lo = b_lo_fixed + ((w_base + ws * kWSlotBytes) >> 4);
        mbar_wait(w_full(ws), wp, 84);
        if (is_c1) {
          mbar_wait(a_ready(0), pg, 82);
          for (int t = 0; t < kS; ++t) {
            const uint32_t ring_i = gen * kS + (uint32_t)t;
            const uint32_t slot = ring_i % kRing;
            if (t + 1 < kS) mbar_wait(a_ready(t + 1), pg, 83);
            // no wait on acc1_empty: with kRing == kS the slot of tile t is always slot t, and a_ready(t) of this iteration is arrived
            // by crew t AFTER its epilogue 1 of the previous iteration read that slot - one barrier round trip less per tile
            // (a satisfied wait costs the issuer ~250 clk next to a busy pipe, a k = 3 conv of one tile is 12 MMAs = ~600)
            tc_fence_after();
            issue_tile_acc<kC / 16>(tm_ring + slot * kC, a1_lo_fixed + ((a_base + (uint32_t)(kPadA + t * kTileM - dil) * 16u) >> 4), a1_hi,
                                    w_lo, b_hi, idesc, kTaps, (uint32_t)dil, a1_kstep, b_kstep, 0u);
            tc_commit(acc1_full(t));
          }
        } else {
          mbar_wait(mid_ready(0), pg, 86);
          for (int t = 0; t < kS; ++t) {
            if (t + 1 < kS) mbar_wait(mid_ready(t + 1), pg, 87);
            tc_fence_after();
            issue_tile_acc<kC / 16>(tm_x + (uint32_t)t * kC, a2_lo_fixed + ((m_base + (uint32_t)(kPadM + t * kTileM - 1) * 16u) >> 4), a2_hi,
                                    w_lo, b_hi, idesc, kTaps, 1u, a2_kstep, b_kstep, 1u);
            tc_commit(x_full(t));
          }
        }
        tc_commit(w_empty(ws));
      }
  } else {
    // ------------------------------------------------------------------ tile crews
    const int q = warp & 3;                                          // TMEM lane quarter this warp may touch
    const int t = (warp - 1 - kIssuers) >> 2;                        // the tile this crew owns
    const int lrow = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t t_ring = tm_ring + lane_off;
    const uint32_t x_addr = tm_x + lane_off + (uint32_t)t * kC;
    const uint32_t mid_dst = m_base + (uint32_t)(kPadM + t * kTileM + lrow) * 16u;
    const uint32_t a_dst = a_base + (uint32_t)(kPadA + t * kTileM + lrow) * 16u;
    const size_t plane_elems = (size_t)R * 8;
    // X[t] <- x0 = lrelu^-1(a) = min(a, 10 a),  A[t] <- a: the start of the ResBlock
    auto init_tile = [&](int g) {
      uint4 av[kPlanes];
#pragma unroll
      for (int pl = 0; pl < kPlanes; ++pl) av[pl] = make_uint4(0u, 0u, 0u, 0u);
      if (g >= 0 && g < R) {
#pragma unroll
        for (int pl = 0; pl < kPlanes; ++pl) av[pl] = *reinterpret_cast<const uint4*>(c.a + (size_t)pl * plane_elems + (size_t)g * 8);
      }
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t v[32];
#pragma unroll
        for (int p4 = 0; p4 < 4; ++p4) {
          float f[8];
          unpack_f16x8(av[4 * hf + p4], f);
#pragma unroll
          for (int e = 0; e < 8; ++e) v[8 * p4 + e] = __float_as_uint(fminf(f[e], 10.f * f[e]));
        }
        tmem_st32(x_addr + 32u * hf, v);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
#pragma unroll
      for (int pl = 0; pl < kPlanes; ++pl) sts128(a_dst + (uint32_t)(pl * kRowsA) * 16u, av[pl].x, av[pl].y, av[pl].z, av[pl].w);
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_ready(t));
    };
    uint32_t gen = 0;
    bool first = true;
    for (int u = blockIdx.x; u < prm.n_super; u += gridDim.x) {
      const int g0 = u * kValid - kHalo;
      const bool has_next = u + (int)gridDim.x < prm.n_super;
      const int g = g0 + t * kTileM + lrow;
      const int g_next = (u + (int)gridDim.x) * kValid - kHalo + t * kTileM + lrow;
      // rows in gaps / outside the sequence must read as zero padding
      const uint32_t keep = (g >= 0 && g < R && c.row_utt[g >> prm.row_div_shift] >= 0) ? 0xFFFFFFFFu : 0u;
      if (first) {
        init_tile(g);
        first = false;
      }
      for (int m = 0; m < 3; ++m, ++gen) {
        // ---------------- epilogue 1: ring slot -> MID = lrelu(c1 + b1)
        {
          const float* b = prm.b1[m];
          const uint32_t ring_i = gen * kS + (uint32_t)t;
          const uint32_t slot = ring_i % kRing;
          mbar_wait(acc1_full(t), gen & 1u, 88);
          tc_fence_after();
          uint32_t v[kC];
          tmem_ld32_nw(t_ring + slot * kC, v);
          tmem_ld32_nw(t_ring + slot * kC + 32u, v + 32);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc1_empty(slot));   // the accumulator is in registers: hand the slot back
#pragma unroll
          for (int gq = 0; gq < kPlanes; ++gq) {
            float y[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float s = __uint_as_float(v[8 * gq + e]) + b[8 * gq + e];
              y[e] = fmaxf(s, 0.1f * s);
            }
            sts128(mid_dst + (uint32_t)(gq * kRowsM) * 16u, pack_f16x2(y[0], y[1]) & keep, pack_f16x2(y[2], y[3]) & keep,
                   pack_f16x2(y[4], y[5]) & keep, pack_f16x2(y[6], y[7]) & keep);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(mid_ready(t));
        }
        // ---------------- epilogue 2: X[t] (fp32, stays in TMEM) -> A = lrelu(X + bcum), or the ResBlock's output
        const float* b = prm.bcum[m];
        mbar_wait(x_full(t), gen & 1u, 89);
        tc_fence_after();
        uint32_t v[kC];
        tmem_ld32_nw(x_addr, v);
        tmem_ld32_nw(x_addr + 32u, v + 32);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (m < 2) {
#pragma unroll
          for (int gq = 0; gq < kPlanes; ++gq) {
            float y[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float s = __uint_as_float(v[8 * gq + e]) + b[8 * gq + e];
              y[e] = fmaxf(s, 0.1f * s);
            }
            sts128(a_dst + (uint32_t)(gq * kRowsA) * 16u, pack_f16x2(y[0], y[1]) & keep, pack_f16x2(y[2], y[3]) & keep,
                   pack_f16x2(y[4], y[5]) & keep, pack_f16x2(y[6], y[7]) & keep);
          }
          fence_proxy_async();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(a_ready(t));
          continue;
        }
        // y = X + bcum[2]: rows this super tile owns go to HBM (fp16, planar), zeros on gap rows
        const int r_local = t * kTileM + lrow;
        if (r_local >= kHalo && r_local < kRows - kHalo && g < R) {
#pragma unroll
          for (int gq = 0; gq < kPlanes; ++gq) {
            float y[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) y[e] = __uint_as_float(v[8 * gq + e]) + b[8 * gq + e];
            *reinterpret_cast<uint4*>(c.out_raw + (size_t)gq * plane_elems + (size_t)g * 8) =
                make_uint4(pack_f16x2(y[0], y[1]) & keep, pack_f16x2(y[2], y[3]) & keep, pack_f16x2(y[4], y[5]) & keep,
                           pack_f16x2(y[6], y[7]) & keep);
          }
        }
        tc_fence_before();
        if (has_next) init_tile(g_next);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace

bool umma_resblock_supported(int C, int taps) { return C == kC && taps == kTaps; }

int umma_resblock(const UmmaResBlock& c, cudaStream_t st) {
  VS_REQUIRE(c.a && c.out_raw && c.row_utt && c.R > 0, "umma_resblock: null pointer");
  Params prm;
  prm.c = c;
  int s = 0;
  while ((1 << s) < c.row_div) ++s;
  VS_REQUIRE((1 << s) == c.row_div, "umma_resblock: row_div=%d must be a power of two", c.row_div);
  prm.row_div_shift = s;
  prm.n_super = (c.R + kValid - 1) / kValid;
  float cum[kC] = {0.f};
  for (int m = 0; m < 3; ++m) {
    VS_REQUIRE(c.w[m][0] && c.w[m][1] && c.b1_host[m] && c.b2_host[m], "umma_resblock: missing weights");
    for (int e = 0; e < kC; ++e) {
      cum[e] += c.b2_host[m][e];
      prm.b1[m][e] = c.b1_host[m][e];
      prm.bcum[m][e] = cum[e];
    }
  }
  int n_sm = 0;
  VS_TRY(device_sm_count(&n_sm));
  VS_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(umma_resblock_kernel), (int)kSmemBytes));
  const int grid = prm.n_super < n_sm ? prm.n_super : n_sm;
  VS_CUDA_CHECK(launch_pdl<4>(umma_resblock_kernel, dim3(grid), dim3(kThreads), kSmemBytes, st, prm));
  VS_LAUNCH_CHECK();
  return VS_OK;
}

}  // namespace vs
