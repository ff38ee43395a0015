// The whole last MRF stage of the HiFi-GAN decoder (C = 32 channels, the waveform's sample rate) as ONE tcgen05 kernel:
//
//     x0 (ups[3] output)  ->  3 x ResBlock1 (k = 3, 7, 11; dilations 1, 3, 5)  ->  sum / 3  ->  leaky_relu(0.01)
//                         ->  conv_post (32 -> 1, k = 7)  ->  tanh  ->  waveform
//
// (reference models.py:279-288, modules.py:210-223).  Why one kernel:
//   * PRECISION.  The residual stream x, the MRF sum and the conv_post input are the DIRECT signal path to the output:
//     rounding them to 16 bits (once per ResBlock iteration in the unfused chain) is 9/10 of the decoder's log-mel error
//     (precision).  Here x lives in TMEM as the fp32 accumulator of conv2 itself - the MMA of c2 accumulates
//     straight onto it, so the residual add costs nothing and is never rounded - the running MRF sum lives in TMEM too,
//     and conv_post runs on fp32 values.  Only the conv OPERANDS (lrelu(x), lrelu(c1)) are fp16; their rounding is
//     filtered by the next conv's weights and is harmless.
//   * TRAFFIC.  The stage used to move 18 activation-sized tensors through HBM (9 fused conv-pair launches + conv_post);
//     now it reads x0 and writes the waveform.
//
// Tiling.  A CTA owns a SUPER TILE of S = 6 row tiles (768 rows).  Rows outside the super tile count as zero, which
// corrupts at most 60 (k = 11: 10 + 20 + 30) + 3 (conv_post) rows at either end: the super tile advances by 768 - 2*64 =
// 640 rows and only those are written (1.2x recompute; the tensor pipe of this stage is bound by the N = 32 operand-
// fetch floor, not by HBM).  TMEM (512 columns): X[t] 32 columns per tile (192), SUM[t] (192), and a ring of 4
// accumulators for conv1 (128).
//
// Roles (864 threads):
//   warp 0      weight producer: streams the 18 convs' weight slabs (<= 22 KB each) through a 4-slot ring, in the order
//               the MMA warps consume them (out of L2: 252 KB per super tile);
//   warps 1, 2  MMA issuers: warp 1 issues conv1(t) : A -> ring slot, warp 2 conv2(t) : MID -> X[t] (accumulate);
//   warps 3-26  six TILE CREWS of four warps (one warp per TMEM lane quarter); the crew of tile t runs both epilogues of
//               its tile, which alternate in time anyway:
//               epilogue 1: ring slot -> + b1 -> lrelu -> mask -> fp16 -> MID (smem, UMMA layout);
//               epilogue 2: X[t] -> + cumulative b2 -> lrelu -> mask -> fp16 -> A (the next iteration's operand); after a
//               ResBlock's last iteration: SUM[t] (+)= X[t], then X[t] <- x0 and A <- lrelu(x0) for the next ResBlock
//               (x0 = hi + lo read from HBM/L2 as two fp16 planar tensors: fp32-exact to 22 bits); after the last ResBlock:
//               v = lrelu(SUM/3, 0.01); conv_post as 7 per-tap partial dot products per row into a small smem table; once
//               per super tile all crews combine the taps across rows, tanh, store.
// One epilogue is a chain of shared-memory / TMEM round trips (~1.2k clk with the tensor pipe running), i.e. longer than
// the MMAs of a k = 3 conv on all six tiles (1.6k clk): the kernel is bound by how many such chains are in flight.  With
// two crews of 2 x 4 warps (tiles round-robin) it took 7.3 ms at the C2 size, k = 3 and k = 7 iterations waiting for the
// epilogues; one crew per tile keeps six chains in flight.  Every hand-off is an mbarrier, and a conv at tile t (which
// reads rows of tiles t-1 and t+1) waits for all three tiles.
#include "umma_conv.cuh"
#include "umma_common.cuh"

namespace vs {
namespace {

using namespace umma;

constexpr int kC = 32;
constexpr int kS = 6;                               // row tiles per super tile
constexpr int kRows = kS * kTileM;                  // 768
constexpr int kHalo = 64;                           // >= 60 + 3 rows of boundary contamination per side
constexpr int kValid = kRows - 2 * kHalo;           // 640 rows written per super tile
constexpr int kPadA = 32;                           // zero rows either side of A (>= 25: k = 11, d = 5)
constexpr int kPadM = 8;                            // zero rows either side of MID (>= 5)
constexpr int kRowsA = kRows + 2 * kPadA;           // 832
constexpr int kRowsM = kRows + 2 * kPadM;           // 784
constexpr int kPlanes = kC / 8;
constexpr int kRing = (512 - 2 * kC * kS) / kC;     // 4 conv1 accumulators
constexpr int kWSlots = 4;
constexpr uint32_t kTapBytes = kC * kC * 2;         // 2 KB: one tap's [K = 32][N = 32] slab
constexpr uint32_t kWSlotBytes = 11 * kTapBytes;    // one conv, k <= 11
constexpr int kIssuers = 2;                         // MMA-issuing warps: conv1 | conv2
constexpr int kCrewWarps = 4 * kS;                  // one crew of four warps (one per TMEM lane quarter) per tile
constexpr int kThreads = (1 + kIssuers + kCrewWarps) * 32;
static_assert(kThreads <= 1024 && 32 * kCrewWarps >= kValid, "umma_mrf: crew size");
constexpr int kPostTaps = 7;

constexpr uint32_t kOffA = 0;
constexpr uint32_t kOffM = kOffA + kPlanes * kRowsA * 16;
constexpr uint32_t kOffW = kOffM + kPlanes * kRowsM * 16;
constexpr uint32_t kOffP = kOffW + kWSlots * kWSlotBytes;
constexpr uint32_t kOffBar = kOffP + kPostTaps * kRows * 4;
constexpr int kNumBars = 2 * kWSlots + kRing + 4 * kS;
constexpr uint32_t kSmemBytes = kOffBar + 8 * kNumBars + 16;
static_assert(kSmemBytes <= 227 * 1024, "umma_mrf: shared memory");

#ifdef VS_UMMA_TIMING
#define VS_TIMED(var, stmt)                         \
  do {                                              \
    const long long _t0 = dbg ? clock64() : 0;      \
    stmt;                                           \
    if (dbg) var += clock64() - _t0;                \
  } while (0)
#else
#define VS_TIMED(var, stmt) stmt
#endif

struct Params {
  UmmaMrf c;
  long long* dbg;              // wait-clock counters (only with -DVS_UMMA_TIMING; tools/mrf_timing.py)
  int row_div_shift, n_super;
  float b1[3][3][kC];          // c1 biases                              [resblock][iteration][channel]
  float bcum[3][3][kC];        // cumulative c2 biases: x_m = X[t] + bcum[j][m] after iteration m
  float post_w[kPostTaps][kC];
};
static_assert(sizeof(Params) <= 4000, "umma_mrf: kernel parameter block");

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld16h(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1) umma_mrf_kernel(const __grid_constant__ Params prm) {
  extern __shared__ __align__(128) uint8_t smem[];
  pdl_trigger();
  const UmmaMrf& c = prm.c;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
#ifdef VS_UMMA_TIMING
  long long* const dbg = prm.dbg;
  long long tw0 = 0, tw1 = 0, tw2 = 0, tw3 = 0;
  const long long t_start = dbg ? clock64() : 0;
#endif
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t a_base = smem_base + kOffA, m_base = smem_base + kOffM, w_base = smem_base + kOffW;
  const uint32_t bar = smem_base + kOffBar;
  auto w_full = [&](uint32_t i) { return bar + 8u * i; };
  auto w_empty = [&](uint32_t i) { return bar + 8u * (kWSlots + i); };
  // acc1_full is per TILE (phase = ResBlock iteration), not per ring slot: the tile crews wait independently of each other,
  // and a parity wait on a slot shared with another tile could pass on that tile's completion
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
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                 "r"(512u)
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
  const uint32_t tm_x = tmem_base, tm_sum = tmem_base + kC * kS, tm_ring = tmem_base + 2 * kC * kS;
  const int R = c.R;

  if (warp == 0) {
    // ------------------------------------------------------------------ weight producer
    if (lane == 0) {
      uint32_t wi = 0;
      for (int u = blockIdx.x; u < prm.n_super; u += gridDim.x)
        for (int q = 0; q < 18; ++q, ++wi) {
          const uint32_t slot = wi % kWSlots, ph = (wi / kWSlots) & 1u;
          mbar_wait(w_empty(slot), ph ^ 1u, 41);
          const int j = q / 6, m = (q >> 1) % 3, which = q & 1;
          const uint32_t bytes = (uint32_t)(3 + 4 * j) * kTapBytes;
          mbar_arrive_expect_tx(w_full(slot), bytes);
          bulk_g2s(w_base + slot * kWSlotBytes, c.w[j][m][which], bytes, w_full(slot));
        }
    }
  } else if (warp <= kIssuers) {
    // ------------------------------------------------------------------ MMA issuers (warp-uniform, elected lane issues)
    // Two of them, on two schedulers: warp 1 issues every conv1 (A -> ring slot), warp 2 every conv2 (MID -> X[t],
    // accumulate).  The streams are independent: every ordering between them is a data dependency that already goes
    // through an mbarrier (conv2(t) waits for epilogue 1 of tiles t-1..t+1, i.e. for conv1 of those tiles; conv1 of the next
    // iteration waits for epilogue 2 of tiles t-1..t+1, i.e. for conv2 of those tiles), so a tcgen05.commit only ever has to
    // cover the issuing warp's own MMAs.
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
      for (int j = 0; j < 3; ++j)
        for (int m = 0; m < 3; ++m, wi += 2, ++gen) {
          const int taps = 3 + 4 * j, dil = 2 * m + 1;
          const uint32_t ws = wi % kWSlots, wp = (wi / kWSlots) & 1u;      // this conv's weight slot
          const uint32_t pg = gen & 1u;
          const uint32_t w_lo = b_lo_fixed + ((w_base + ws * kWSlotBytes) >> 4);
          VS_TIMED(tw3, mbar_wait(w_full(ws), wp, 44));
          // a conv at tile t reads rows of tiles t-1, t, t+1 (reach <= 25 rows); the crews finish tiles out of order
          // (one crew per tile), so each of the three is waited for - t-1 and t were already seen at the previous tile
          if (is_c1) {
            const int h1 = dil * (taps - 1) / 2;
            VS_TIMED(tw0, mbar_wait(a_ready(0), pg, 41));
            for (int t = 0; t < kS; ++t) {
              const uint32_t ring_i = gen * kS + (uint32_t)t;
              const uint32_t slot = ring_i % kRing, rp = (ring_i / kRing) & 1u;
              if (t + 1 < kS) VS_TIMED(tw0, mbar_wait(a_ready(t + 1), pg, 42));
              VS_TIMED(tw1, mbar_wait(acc1_empty(slot), rp ^ 1u, 43));
              tc_fence_after();
              VS_TIMED(tw2, issue_tile_acc<kC / 16>(tm_ring + slot * kC, a1_lo_fixed + ((a_base + (uint32_t)(kPadA + t * kTileM - h1) * 16u) >> 4),
                                      a1_hi, w_lo, b_hi, idesc, taps, (uint32_t)dil, a1_kstep, b_kstep, 0u));
              tc_commit(acc1_full(t));
            }
          } else {
            const int h2 = (taps - 1) / 2;
            VS_TIMED(tw0, mbar_wait(mid_ready(0), pg, 40));
            for (int t = 0; t < kS; ++t) {
              if (t + 1 < kS) VS_TIMED(tw0, mbar_wait(mid_ready(t + 1), pg, 45));
              tc_fence_after();
              VS_TIMED(tw2, issue_tile_acc<kC / 16>(tm_x + (uint32_t)t * kC, a2_lo_fixed + ((m_base + (uint32_t)(kPadM + t * kTileM - h2) * 16u) >> 4),
                                      a2_hi, w_lo, b_hi, idesc, taps, 1u, a2_kstep, b_kstep, 1u));
              tc_commit(x_full(t));
            }
          }
          tc_commit(w_empty(ws));           // every MMA of this conv has read its weights
        }
  } else {
    // ------------------------------------------------------------------ tile crews: epilogue 1, epilogue 2, hand-over, conv_post
    // Four warps (one per TMEM lane quarter) own ONE tile of the super tile and run both of its epilogues, which alternate
    // in time anyway (conv1 -> epilogue 1 -> conv2 -> epilogue 2 -> next conv1).  Six tiles = six epilogue chains in flight.
    const int q = warp & 3;                                          // TMEM lane quarter this warp may touch
    const int t = (warp - 1 - kIssuers) >> 2;                        // the tile this crew owns
    const int lrow = q * 32 + lane;                                  // row within the tile = TMEM lane
    const int ctid = (warp - 1 - kIssuers) * 32 + lane;              // 0 .. 32 * kCrewWarps - 1 over all crews
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t t_ring = tm_ring + lane_off;
    const uint32_t x_addr = tm_x + lane_off + (uint32_t)t * kC, sum_addr = tm_sum + lane_off + (uint32_t)t * kC;
    const uint32_t mid_dst = m_base + (uint32_t)(kPadM + t * kTileM + lrow) * 16u;
    const uint32_t a_dst = a_base + (uint32_t)(kPadA + t * kTileM + lrow) * 16u;
    float* const P = reinterpret_cast<float*>(smem + kOffP);          // [7][kRows] conv_post per-tap partial sums
    const size_t plane_elems = (size_t)R * 8;
    // X[t] <- x0 = hi + lo,  A[t] <- lrelu(x0): the start of a ResBlock
    auto init_tile = [&](int g) {
      uint4 xh[kPlanes], xl[kPlanes];
#pragma unroll
      for (int pl = 0; pl < kPlanes; ++pl) { xh[pl] = make_uint4(0u, 0u, 0u, 0u); xl[pl] = make_uint4(0u, 0u, 0u, 0u); }
      if (g >= 0 && g < R) {
#pragma unroll
        for (int pl = 0; pl < kPlanes; ++pl) {
          xh[pl] = *reinterpret_cast<const uint4*>(c.x_hi + (size_t)pl * plane_elems + (size_t)g * 8);
          xl[pl] = *reinterpret_cast<const uint4*>(c.x_lo + (size_t)pl * plane_elems + (size_t)g * 8);
        }
      }
      uint32_t v[32];
#pragma unroll
      for (int pl = 0; pl < kPlanes; ++pl) {
        float h[8], l[8], y[8];
        unpack_f16x8(xh[pl], h);
        unpack_f16x8(xl[pl], l);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float x = h[e] + l[e];
          v[8 * pl + e] = __float_as_uint(x);
          y[e] = fmaxf(x, 0.1f * x);
        }
        sts128(a_dst + (uint32_t)(pl * kRowsA) * 16u, pack_f16x2(y[0], y[1]), pack_f16x2(y[2], y[3]), pack_f16x2(y[4], y[5]),
               pack_f16x2(y[6], y[7]));
      }
      tmem_st32(x_addr, v);
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_ready(t));
    };
    // x0 of the NEXT super tile comes from HBM: pull its lines into L2 an iteration ahead of the hand-over
    auto prefetch_x0 = [&](int g) {
      if (g >= 0 && g < R) {
#pragma unroll
        for (int pl = 0; pl < kPlanes; ++pl) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(c.x_hi + (size_t)pl * plane_elems + (size_t)g * 8));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(c.x_lo + (size_t)pl * plane_elems + (size_t)g * 8));
        }
      }
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
      for (int j = 0; j < 3; ++j)
        for (int m = 0; m < 3; ++m, ++gen) {
          // ---------------- epilogue 1: ring slot -> MID = lrelu(c1 + b1)
          {
            const float* b = prm.b1[j][m];
            const uint32_t ring_i = gen * kS + (uint32_t)t;
            const uint32_t slot = ring_i % kRing;
            VS_TIMED(tw0, mbar_wait(acc1_full(t), gen & 1u, 47));
            tc_fence_after();
            uint32_t v[32];
            tmem_ld32(t_ring + slot * kC, v);
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
          if (j == 2 && m == 1 && has_next) prefetch_x0(g_next);
          // ---------------- epilogue 2: X[t] (fp32, stays in TMEM) -> A = lrelu(X + bcum), or the ResBlock hand-over
          const float* b = prm.bcum[j][m];
          VS_TIMED(tw1, mbar_wait(x_full(t), gen & 1u, 48));
          tc_fence_after();
          if (m < 2) {
            uint32_t v[32];
            tmem_ld32(x_addr, v);
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
          // the ResBlock's output y_j = X + bcum[j][2]; SUM (+)= y_j, in two halves of 16 channels (register budget)
          float pacc[kPostTaps];
#pragma unroll
          for (int tp = 0; tp < kPostTaps; ++tp) pacc[tp] = 0.f;
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            uint32_t v[16];
            tmem_ld16h(x_addr + 16u * hf, v);
            if (j > 0) {
              uint32_t s[16];
              tmem_ld16h(sum_addr + 16u * hf, s);
#pragma unroll
              for (int e = 0; e < 16; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) + b[16 * hf + e] + __uint_as_float(s[e]));
            } else {
#pragma unroll
              for (int e = 0; e < 16; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) + b[16 * hf + e]);
            }
            if (j < 2) {
              tmem_st16(sum_addr + 16u * hf, v);
            } else {
              // xs / 3 -> leaky_relu (default slope 0.01, models.py:286) -> conv_post taps as per-row partial sums
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                const float x = __uint_as_float(v[e]) * (1.f / 3.f);
                const float f = keep ? fmaxf(x, 0.01f * x) : 0.f;
#pragma unroll
                for (int tp = 0; tp < kPostTaps; ++tp) pacc[tp] = fmaf(f, prm.post_w[tp][16 * hf + e], pacc[tp]);
              }
            }
          }
          if (j == 2) {
#pragma unroll
            for (int tp = 0; tp < kPostTaps; ++tp) P[tp * kRows + t * kTileM + lrow] = pacc[tp];
          }
          if (j < 2) init_tile(g);
          else if (has_next) init_tile(g_next);
        }
      // conv_post: out[r] = tanh(sum_tap P[tap][r + tap - 3]) for the 640 rows this super tile owns
      VS_TIMED(tw2, asm volatile("bar.sync 1, %0;" ::"n"(32 * kCrewWarps) : "memory"));
      if (ctid < kValid) {
        const int r = kHalo + ctid;
        const int go = g0 + r;
        float acc = 0.f;
#pragma unroll
        for (int tp = 0; tp < kPostTaps; ++tp) acc += P[tp * kRows + r + tp - 3];
        if (go < R) c.wave[go] = c.row_utt[go >> prm.row_div_shift] >= 0 ? tanhf(acc) : 0.f;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kCrewWarps) : "memory");
    }
  }

#ifdef VS_UMMA_TIMING
  if (dbg && lane == 0 && (warp == 1 || warp == 2 || warp == 1 + kIssuers || warp == 1 + kIssuers + 4 * (kS / 2))) {   // [cta][conv1 | conv2 | crew of tile 0 | crew of tile kS/2][total, waits x 4]
    long long* o = dbg + ((size_t)blockIdx.x * 4 + (warp == 1 ? 0 : warp == 2 ? 1 : warp == 1 + kIssuers ? 2 : 3)) * 5;
    o[0] = clock64() - t_start; o[1] = tw0; o[2] = tw1; o[3] = tw2; o[4] = tw3;
  }
#endif
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace

int umma_mrf(const UmmaMrf& c, cudaStream_t st) {
  VS_REQUIRE(c.x_hi && c.x_lo && c.wave && c.row_utt && c.post_w_host && c.R > 0, "umma_mrf: null pointer");
  Params prm;
  prm.c = c;
  prm.dbg = static_cast<long long*>(umma_conv_timing_buffer());
  int s = 0;
  while ((1 << s) < c.row_div) ++s;
  VS_REQUIRE((1 << s) == c.row_div, "umma_mrf: row_div=%d must be a power of two", c.row_div);
  prm.row_div_shift = s;
  prm.n_super = (c.R + kValid - 1) / kValid;
  for (int j = 0; j < 3; ++j) {
    float cum[kC] = {0.f};
    for (int m = 0; m < 3; ++m) {
      VS_REQUIRE(c.w[j][m][0] && c.w[j][m][1] && c.b1_host[j][m] && c.b2_host[j][m], "umma_mrf: missing weights");
      for (int e = 0; e < kC; ++e) {
        cum[e] += c.b2_host[j][m][e];
        prm.b1[j][m][e] = c.b1_host[j][m][e];
        prm.bcum[j][m][e] = cum[e];
      }
    }
  }
  for (int i = 0; i < kPostTaps * kC; ++i) prm.post_w[i / kC][i % kC] = c.post_w_host[i];
  int n_sm = 0;
  VS_TRY(device_sm_count(&n_sm));
  VS_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(umma_mrf_kernel), (int)kSmemBytes));
  const int grid = prm.n_super < n_sm ? prm.n_super : n_sm;
  VS_CUDA_CHECK(launch_pdl<4>(umma_mrf_kernel, dim3(grid), dim3(kThreads), kSmemBytes, st, prm));
  VS_LAUNCH_CHECK();
  return VS_OK;
}

int umma_mrf_rows_per_cta() { return kValid; }

}  // namespace vs
