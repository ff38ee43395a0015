// One ResBlock1 iteration (reference modules.py:211-220)
//     y = c2( lrelu( c1( lrelu(x) ) + b1 ) ) + b2 + x        (c1: k taps, dilation d;  c2: k taps, dilation 1)
// as ONE tcgen05 kernel on a CTA PAIR (cta_group::2, M = 256 rows per MMA): the fusion of umma_respair.cu (one tensor read, one
// written, the intermediate never leaves the SM) with the operand economy of umma_pair.cu.
//
// Why a pair: at N = 64 a single-CTA MMA re-reads 4 KB of A and 2 KB of B from shared memory for 32 clk of tensor time - the
// 48 clk operand-fetch floor that bounds the C = 64 stage.  In a pair each SM reads its 128 rows of A and only HALF of B (the
// 32 output channels whose weights it holds): 5 KB = 40 clk, both convs' weights are resident at every k (2 x 11 x 64 x 32
// halves = 88 KB per CTA), and at C = 128, k = 3 (96 KB per CTA) the HBM-bound conv pair of that stage becomes one pass.
//
// Per CTA, tile i covers 128 conv1 rows and V = 128 - (k - 1) output rows; a unit is the pair's two consecutive tiles.
//   warp 0       producer    A[slot] <- rows [s0 - h2 - h2 d, + 128 + (k-1) d) of a = lrelu(x), 64-channel chunks (zero fill outside)
//   warp 1       MMA issuer  (leader CTA only)  conv1(u): A -> acc1   then   conv2(u-1): MID -> acc2
//   warp 2       relay       "my operands have landed" -> the leader's barrier
//   warps 3-6    epilogue 1  acc1 -> + b1 -> lrelu -> mask -> f16 -> MID (shared memory, UMMA K-major layout)
//   warps 7-14   epilogue 2  acc2 + b2 + lrelu^-1(a) [+ MRF sum] -> [scale, lrelu] -> HBM
// Barriers the issuer waits on live in the leader CTA (the peer's warps arrive remotely).  What the issuer signals is ONE
// tcgen05.commit per unit, multicast to the same barrier R[n % 4] of both CTAs: measured (tools/pair_timing.py), every
// cta_group::2 commit costs the tensor pipe ~400 clk - more than the 12 MMAs of a k = 3 conv at N = 64 - so round n issues
// conv1(n) and conv2(n - LAG) and then commits once, and everybody else derives what they need from "round n is complete":
// the producer may refill the A slots of unit n, epilogue 1 may read acc1(n) and overwrite MID[n % LAG] (its last reader
// conv2(n - LAG) is part of the same round), epilogue 2 may read acc2(n - LAG).  LAG = 2 at C = 64 (a conv is shorter than
// epilogue 1's latency, so conv2 trails by two units and MID is double-buffered), 1 at C = 128.
#include "umma_conv.cuh"
#include "umma_common.cuh"

namespace vs {
namespace {

using namespace umma;
constexpr int kEpi1Warps = 4, kEpi2Warps = 8;
constexpr int kThreads = (3 + kEpi1Warps + kEpi2Warps) * 32;     // 480
constexpr int kMaxSlots = 4;
constexpr int kRounds = 8;                         // rings of the two per-round barriers (> LAG + accumulators per conv: epilogue 2 arrives that far ahead)
constexpr int kKCH = 64;
constexpr int kChunkPlanes = kKCH / 8;
constexpr int kPairM = 2 * kTileM;
enum { M_ACT = 0, M_RAW = 1, M_RAW_RES2 = 2, M_ACT_RES2_SCALE = 3, M_GENERIC = 4 };

struct Plan {
  int V, h2, rows_a, rows_m, halo_a, n_chunks, planes, nhalf, nslot, units_a, n_tiles, n_units, row_div_shift;
  uint32_t slot_bytes, mid_bytes, w_bytes, off_mid, off_w, off_bar, smem_bytes;
};
struct Params {
  UmmaPair c;
  Plan p;
  float bias[2][128];     // b1, b2 in the kernel's constant bank (an LDS costs ~200 clk while the tensor pipe owns shared memory)
  long long* dbg;
};
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

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {    // default .release.cta semantics, see umma_pair.cu
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\t.reg .b16 m;\n\t"
      "mov.b16 m, 3;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}" ::"r"(bar)
      : "memory");
}
__device__ __forceinline__ void tc_mma_pair(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, pe;\n\t.reg .b64 da, db;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// the same two for a caller that is already ONE thread (no election, no warp-collective bookkeeping around every instruction)
__device__ __forceinline__ void tc_commit_pair_1t(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .b16 m;\n\t"
      "mov.b16 m, 3;\n\t"
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}" ::"r"(bar)
      : "memory");
}
__device__ __forceinline__ void tc_mma_pair_1t(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

template <int C, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) umma_pairfused_kernel(const __grid_constant__ Params prm) {
  extern __shared__ __align__(128) uint8_t smem[];
  pdl_trigger();
  const UmmaPair& c = prm.c;
  const Plan& p = prm.p;
  constexpr int N = C, NK = C / 16;                       // K = 16 steps per tap over all channels
  constexpr int NACC = 256 / N;                           // accumulators per conv: 2 x NACC x N = 512 TMEM columns
  constexpr int LAG = C == 64 ? 2 : 1;                    // conv2 trails conv1 by LAG units; MID has LAG buffers
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t rank = cluster_rank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
#ifdef VS_UMMA_TIMING
  long long* const dbg = prm.dbg;
  long long tw0 = 0, tw1 = 0, tw2 = 0;
  const long long t_start = dbg ? clock64() : 0;
#endif

  const uint32_t smem_base = smem_u32(smem);
  const uint32_t a_base = smem_base, mid_base = smem_base + p.off_mid, w_base = smem_base + p.off_w, bar_base = smem_base + p.off_bar;
  const uint32_t w2_base = w_base + p.w_bytes;
  // barriers (8 B each; the same offsets in both CTAs)
  auto a_land = [&](int i) { return bar_base + 8u * i; };                        // local: this CTA's chunk has landed
  auto go_a = [&](int i) { return bar_base + 8u * (8 + i); };                    // leader's: A of unit n has landed in both CTAs (2 relays)
  auto a_done = [&](int i) { return bar_base + 8u * (35 + i); };                 // local, commit multicast after conv1(n): its A slots are free
  auto go_m = [&](int i) { return bar_base + 8u * (24 + i); };                   // leader's: conv2 of round n may run (8 + 16 epilogue warps)
  auto round_done = [&](int i) { return bar_base + 8u * (16 + i); };             // local, THE commit multicast of round n % 8
  const uint32_t w_land = bar_base + 8u * 32, w_full = bar_base + 8u * 33;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + p.off_bar + 8 * 34);

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.nslot; ++i) mbar_init(a_land(i), 1);
    for (int i = 0; i < kRounds; ++i) { mbar_init(round_done(i), 1); mbar_init(go_a(i), 2); mbar_init(a_done(i), 1); mbar_init(go_m(i), 2 * kEpi1Warps + 2 * kEpi2Warps); }
    mbar_init(w_land, 1);
    mbar_init(w_full, 2);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  // the rows of MID past the 128 that epilogue 1 writes are only ever read for output rows that are discarded; zero them once
  for (int i = threadIdx.x; i < LAG * p.planes * (p.rows_m - kTileM); i += kThreads) {
    const int b = i / (p.planes * (p.rows_m - kTileM)), rem = i % (p.planes * (p.rows_m - kTileM));
    const int pl = rem / (p.rows_m - kTileM), row = kTileM + rem % (p.rows_m - kTileM);
    sts128(mid_base + (uint32_t)b * p.mid_bytes + (uint32_t)(pl * p.rows_m + row) * 16u, 0u, 0u, 0u, 0u);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  pdl_wait();                              // the prologue above overlapped the previous kernel's tail
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_acc1 = tmem_base, tm_acc2 = tmem_base + 256u;
  const int R = c.R, taps = c.taps, dil = c.dil;

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    {
      const uint32_t plane_bytes = (uint32_t)p.nhalf * 16u;
      const int n_wslabs = taps * p.planes;
      if (lane == 0) mbar_arrive_expect_tx(w_land, 2u * p.w_bytes);
      __syncwarp();
      for (int sl = lane; sl < 2 * n_wslabs; sl += 32) {
        const int which = sl / n_wslabs, s1 = sl % n_wslabs;
        const __half* src = (which ? c.w2 : c.w1) + ((size_t)s1 * N + (size_t)rank * p.nhalf) * 8;
        bulk_g2s(w_base + (uint32_t)which * p.w_bytes + (uint32_t)s1 * plane_bytes, src, plane_bytes, w_land);
      }
    }
    uint32_t slot = 0;
    int n = 0;
    for (int u = pair; u < p.n_units; u += n_pairs, ++n) {
      // the A slots of unit n were last read by conv1(n - units_a): complete when that round is
      if (n >= p.units_a) VS_TIMED(tw0, mbar_wait(a_done((n - p.units_a) % kRounds), (uint32_t)((n - p.units_a) / kRounds) & 1u, 1));
      const int tile = 2 * u + (int)rank;
      const int row_lo = tile * p.V - p.h2 - p.halo_a, row_hi = row_lo + p.rows_a;
      const int c_lo = row_lo < 0 ? 0 : row_lo, c_hi = row_hi > R ? R : row_hi;
      const int n_zero_lo = c_lo - row_lo;
      const int n_rows = c_hi > c_lo ? c_hi - c_lo : 0;
      const int n_zero_hi = p.rows_a - n_zero_lo - n_rows;
      for (int ch = 0; ch < p.n_chunks; ++ch) {
        const uint32_t stage = a_base + slot * p.slot_bytes;
        if (n_rows < p.rows_a) {
          const int per_plane = p.rows_a - n_rows;
          for (int i = lane; i < kChunkPlanes * per_plane; i += 32) {
            const int pl = i / per_plane, j = i % per_plane;
            const int row = j < n_zero_lo ? j : (p.rows_a - n_zero_hi + (j - n_zero_lo));
            sts128(stage + (uint32_t)(pl * p.rows_a + row) * 16u, 0u, 0u, 0u, 0u);
          }
          fence_proxy_async();
        }
        __syncwarp();
        const uint32_t bytes = (uint32_t)n_rows * 16u;
        if (lane == 0) {
          if (bytes) mbar_arrive_expect_tx(a_land(slot), bytes * kChunkPlanes);
          else mbar_arrive(a_land(slot));
        }
        __syncwarp();
        if (bytes && lane < kChunkPlanes)
          bulk_g2s(stage + (uint32_t)(lane * p.rows_a + n_zero_lo) * 16u, c.x + ((size_t)(ch * kChunkPlanes + lane) * R + c_lo) * 8, bytes,
                   a_land(slot));
        if (++slot == (uint32_t)p.nslot) slot = 0;
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ relay: "unit n of my A has landed" -> go_a(n) of the leader
    if (lane == 0) {
      mbar_wait(w_land, 0, 2);
      mbar_arrive_remote(map_to_cta(w_full, 0));
      uint32_t slot = 0, ph = 0;
      int n = 0;
      for (int u = pair; u < p.n_units; u += n_pairs, ++n) {
        for (int ch = 0; ch < p.n_chunks; ++ch) {
          mbar_wait(a_land(slot), ph, 3);
          if (++slot == (uint32_t)p.nslot) { slot = 0; ph ^= 1u; }
        }
        mbar_arrive_remote(map_to_cta(go_a(n % kRounds), 0));
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (rank == 0) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kPairM >> 4) << 24);
      const uint32_t a_lbo = (uint32_t)p.rows_a * 16u, m_lbo = (uint32_t)p.rows_m * 16u, b_lbo = (uint32_t)p.nhalf * 16u;
      const uint32_t a_hi = (uint32_t)(make_desc(0, a_lbo, 128u) >> 32), m_hi = (uint32_t)(make_desc(0, m_lbo, 128u) >> 32),
                     b_hi = (uint32_t)(make_desc(0, b_lbo, 128u) >> 32);
      const uint32_t a_fixed = (uint32_t)make_desc(0, a_lbo, 128u), m_fixed = (uint32_t)make_desc(0, m_lbo, 128u),
                     b_fixed = (uint32_t)make_desc(0, b_lbo, 128u);
      const uint32_t a_kstep = 2u * (uint32_t)p.rows_a, m_kstep = 2u * (uint32_t)p.rows_m, b_kstep = 2u * (uint32_t)p.nhalf;
      const uint32_t b_tapstep = (uint32_t)p.planes * (uint32_t)p.nhalf, b_chunkstep = (uint32_t)kChunkPlanes * (uint32_t)p.nhalf;
      const uint32_t w1_16 = b_fixed + (w_base >> 4), w2_16 = b_fixed + (w2_base >> 4);
      const uint32_t nslot = (uint32_t)p.nslot;
      const int n_chunks = p.n_chunks;
      uint32_t slot = 0, i1 = 0, i2 = 0, mb = 0;
      mbar_wait(w_full, 0, 4);
      tc_fence_after();
      // Round n = conv1(n), conv2(n - LAG), one commit.  A satisfied mbarrier probe costs the issuer ~250 clk while the tensor
      // pipe runs (the warp stalls on the probe's predicate, so an "early" probe hides nothing) and a k = 3 conv is 12 MMAs of
      // ~40 clk: the issuer therefore waits on TWO barriers per round only - go_a(n), on which the two relays arrive (A of unit n
      // has landed; acc1's slot is free because MID(n - NACC) was written rounds ago), before conv1, and go_m(n), on which the
      // epilogue-1 warps of unit n - LAG (MID written) and the epilogue-2 warps of unit n - LAG - NACC (the acc2 slot that
      // conv2(n - LAG) overwrites is drained) arrive, before conv2.
      auto conv2 = [&]() {
        const uint32_t d2 = tm_acc2 + i2 * (uint32_t)N;
        uint32_t a_tap = m_fixed + ((mid_base + mb * p.mid_bytes) >> 4), b_tap = w2_16;
        uint32_t accumulate = 0;
#pragma unroll 1
        for (int t = 0; t < taps; ++t, a_tap += 1u, b_tap += b_tapstep) {
#pragma unroll
          for (int k = 0; k < NK; ++k) {
            tc_mma_pair(d2, a_tap + (uint32_t)k * m_kstep, m_hi, b_tap + (uint32_t)k * b_kstep, b_hi, idesc, accumulate);
            accumulate = 1;
          }
        }
        if (++i2 == (uint32_t)NACC) i2 = 0;
        if (++mb == (uint32_t)LAG) mb = 0;
      };
      int n = 0;
      for (int u = pair; u < p.n_units; u += n_pairs, ++n) {
        VS_TIMED(tw0, mbar_wait(go_a(n % kRounds), (uint32_t)(n / kRounds) & 1u, 5));
        tc_fence_after();
        const uint32_t d1 = tm_acc1 + i1 * (uint32_t)N;
        uint32_t accumulate = 0;
        uint32_t b_chunk = w1_16;
        for (int ch = 0; ch < n_chunks; ++ch, b_chunk += b_chunkstep) {
          uint32_t a_tap = a_fixed + ((a_base + slot * p.slot_bytes) >> 4), b_tap = b_chunk;
#pragma unroll 1
          for (int t = 0; t < taps; ++t, a_tap += (uint32_t)dil, b_tap += b_tapstep) {
#pragma unroll
            for (int k = 0; k < kKCH / 16; ++k) {
              tc_mma_pair(d1, a_tap + (uint32_t)k * a_kstep, a_hi, b_tap + (uint32_t)k * b_kstep, b_hi, idesc, accumulate);
              accumulate = 1;
            }
          }
          if (++slot == nslot) slot = 0;
        }
        if (++i1 == (uint32_t)NACC) i1 = 0;
        tc_commit_pair(a_done(n % kRounds));          // a commit costs the pipe nothing (tools/pair_microbench.cu)
        if (n >= LAG) {
          VS_TIMED(tw1, mbar_wait(go_m(n % kRounds), (uint32_t)(n / kRounds) & 1u, 7));
          tc_fence_after();
          conv2();
        }
        VS_TIMED(tw2, tc_commit_pair(round_done(n % kRounds)));
      }
      for (int k = 0; k < LAG; ++k) {      // trailing rounds: the conv2s still owed
        const int r = n + k;
        if (r >= LAG) {
          VS_TIMED(tw1, mbar_wait(go_m(r % kRounds), (uint32_t)(r / kRounds) & 1u, 6));
          tc_fence_after();
          conv2();
        }
        tc_commit_pair(round_done(r % kRounds));
      }
    }
    __syncwarp();
  } else if (warp < 3 + kEpi1Warps) {
    // ------------------------------------------------------------------ epilogue 1: acc1 -> MID = lrelu(c1 + b1), masked, f16
    const int q = warp & 3;
    const int j = q * 32 + lane;                                      // conv1 row of the tile
    const float slope = c.in_slope;
    uint32_t i1 = 0, mb = 0;
    if (lane == 0)
      for (int r = 0; r < LAG; ++r) mbar_arrive_remote(map_to_cta(go_m(r), 0));     // rounds 0 .. LAG - 1 have no conv2
    int n = 0;
    for (int u = pair; u < p.n_units; u += n_pairs, ++n) {
      const int tile = 2 * u + (int)rank;
      const int g = tile * p.V - p.h2 + j;                            // global row of MID row j
      const uint32_t keep = (g >= 0 && g < R && (!c.row_utt || c.row_utt[g >> p.row_div_shift] >= 0)) ? 0xFFFFFFFFu : 0u;
      // round n complete: acc1(n) is full AND conv2(n - LAG), the last reader of MID[n % LAG], has finished
      VS_TIMED(tw0, mbar_wait(round_done(n % kRounds), (uint32_t)(n / kRounds) & 1u, 10));
      tc_fence_after();
      const uint32_t t_row = tm_acc1 + ((uint32_t)(q * 32) << 16) + i1 * (uint32_t)N;
      const uint32_t dst = mid_base + mb * p.mid_bytes + (uint32_t)j * 16u;
#pragma unroll
      for (int cc = 0; cc < N / 32; ++cc) {
        uint32_t v[32];
        tmem_ld32(t_row + (uint32_t)(cc * 32), v);
        if (cc == N / 32 - 1) tc_fence_before();                      // the accumulator is in registers
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) {
          float y[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float s = __uint_as_float(v[8 * gq + e]) + prm.bias[0][cc * 32 + gq * 8 + e];
            y[e] = fmaxf(s, slope * s);
          }
          sts128(dst + (uint32_t)((cc * 4 + gq) * p.rows_m) * 16u, pack_f16x2(y[0], y[1]) & keep, pack_f16x2(y[2], y[3]) & keep,
                 pack_f16x2(y[4], y[5]) & keep, pack_f16x2(y[6], y[7]) & keep);
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(map_to_cta(go_m((n + LAG) % kRounds), 0));   // MID(n) written, acc1(n) read
      if (++i1 == (uint32_t)NACC) i1 = 0;
      if (++mb == (uint32_t)LAG) mb = 0;
    }
  } else {
    // ------------------------------------------------------------------ epilogue 2: acc2 + b2 + x [+ res2] -> HBM
    constexpr bool kGeneric = MODE == M_GENERIC;
    const bool has_res2 = kGeneric ? (c.res2 != nullptr) : (MODE == M_RAW_RES2 || MODE == M_ACT_RES2_SCALE);
    const bool has_raw = kGeneric ? (c.out_raw != nullptr) : (MODE == M_RAW || MODE == M_RAW_RES2);
    const bool has_act = kGeneric ? (c.out_act != nullptr) : (MODE == M_ACT || MODE == M_ACT_RES2_SCALE);
    const bool has_scale = kGeneric ? (c.act_scale != 1.f) : (MODE == M_ACT_RES2_SCALE);
    const float rinv = 1.f / c.in_slope, slope = c.act_slope, scale = c.act_scale;
    const int e2 = warp - 3 - kEpi1Warps;
    const int q = warp & 3, hsel = e2 >> 2;
    constexpr int kHalf = N / 2, kCC = kHalf / 32 > 0 ? kHalf / 32 : 1, kCW = kHalf >= 32 ? 32 : kHalf;   // N = 64: one 32-column chunk per warp
    const uint32_t g8_0 = (uint32_t)(hsel * kHalf) >> 3;
    const size_t plane_stride = (size_t)R * 8;
    const int j = q * 32 + lane;
    uint32_t i2 = 0;
    if (lane == 0)
      for (int r = 0; r < LAG + NACC; ++r) mbar_arrive_remote(map_to_cta(go_m(r), 0));   // the first NACC conv2s find their acc2 slots free
    int n = 0;
    for (int u = pair; u < p.n_units; u += n_pairs, ++n) {
      const int tile = 2 * u + (int)rank;
      const int r = tile * p.V + j;
      const bool in_range = j < p.V && r < R;
      bool valid = in_range;
      if (in_range && c.row_utt) valid = c.row_utt[r >> p.row_div_shift] >= 0;
      const size_t row_elem = (size_t)r * 8;
      uint4 rv[kCC * 4], rv2[kCC * 4];
      if (valid) {
#pragma unroll
        for (int g = 0; g < kCC * (kCW / 8); ++g) {
          const size_t o = (size_t)(g8_0 + g) * plane_stride + row_elem;
          rv[g] = *reinterpret_cast<const uint4*>(c.x + o);
          if (has_res2) rv2[g] = *reinterpret_cast<const uint4*>(c.res2 + o);
        }
      }
      VS_TIMED(tw0, mbar_wait(round_done((n + LAG) % kRounds), (uint32_t)((n + LAG) / kRounds) & 1u, 12));   // conv2(n) is part of round n + LAG
      tc_fence_after();
      const uint32_t t_row = tm_acc2 + ((uint32_t)(q * 32) << 16) + i2 * (uint32_t)N + (uint32_t)(hsel * kHalf);
#pragma unroll
      for (int cc = 0; cc < kCC; ++cc) {
        uint32_t v[32];
        tmem_ld32(t_row + (uint32_t)(cc * 32), v);
        if (cc == kCC - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(map_to_cta(go_m((n + LAG + NACC) % kRounds), 0));   // acc2(n) drained: conv2(n + NACC) may overwrite it
        }
        if (in_range) {
#pragma unroll
          for (int g = 0; g < kCW / 8; ++g) {
            const int gg = cc * 4 + g;
            const size_t o = (size_t)(g8_0 + gg) * plane_stride + row_elem;
            uint4 raw = make_uint4(0, 0, 0, 0), act = make_uint4(0, 0, 0, 0);
            if (valid) {
              float y[8], f[8];
              unpack_f16x8(rv[gg], f);
#pragma unroll
              for (int e = 0; e < 8; ++e)
                y[e] = __uint_as_float(v[8 * g + e]) + prm.bias[1][hsel * kHalf + cc * 32 + g * 8 + e] + fminf(f[e], f[e] * rinv);
              if (has_res2) {
                unpack_f16x8(rv2[gg], f);
#pragma unroll
                for (int e = 0; e < 8; ++e) y[e] += f[e];
              }
              if (has_raw) raw = make_uint4(pack_f16x2(y[0], y[1]), pack_f16x2(y[2], y[3]), pack_f16x2(y[4], y[5]), pack_f16x2(y[6], y[7]));
              if (has_act) {
                float z[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  const float t = has_scale ? y[e] * scale : y[e];
                  z[e] = fmaxf(t, t * slope);
                }
                act = make_uint4(pack_f16x2(z[0], z[1]), pack_f16x2(z[2], z[3]), pack_f16x2(z[4], z[5]), pack_f16x2(z[6], z[7]));
              }
            }
            if (has_raw) *reinterpret_cast<uint4*>(c.out_raw + o) = raw;
            if (has_act) *reinterpret_cast<uint4*>(c.out_act + o) = act;
          }
        }
      }
      if (++i2 == (uint32_t)NACC) i2 = 0;
    }
  }

#ifdef VS_UMMA_TIMING
  if (dbg && lane == 0 && (warp < 4 || warp == 7)) {   // [cta][producer | MMA | relay | epilogue 1 | epilogue 2][total, wait0, wait1, wait2]
    long long* o = dbg + ((size_t)blockIdx.x * 5 + (warp == 7 ? 4 : warp)) * 4;
    o[0] = clock64() - t_start; o[1] = tw0; o[2] = tw1; o[3] = tw2;
  }
#endif
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

int make_plan(const UmmaPair& c, Plan* out) {
  Plan p{};
  VS_REQUIRE((c.C == 64 || c.C == 128) && (c.taps & 1) && c.taps >= 3 && c.dil >= 1, "umma_pairfused: unsupported shape C=%d taps=%d", c.C, c.taps);
  p.h2 = (c.taps - 1) / 2;
  p.V = kTileM - (c.taps - 1);
  p.halo_a = p.h2 * c.dil;
  p.rows_a = kTileM + (c.taps - 1) * c.dil;
  p.rows_m = kTileM + (c.taps - 1);
  p.planes = c.C / 8;
  p.n_chunks = c.C / kKCH;
  p.nhalf = c.C / 2;
  p.slot_bytes = (uint32_t)kChunkPlanes * (uint32_t)p.rows_a * 16u;
  p.mid_bytes = (uint32_t)p.planes * (uint32_t)p.rows_m * 16u;
  p.w_bytes = (uint32_t)c.taps * (uint32_t)c.C * (uint32_t)p.nhalf * 2u;
  int s = 0;
  while ((1 << s) < c.row_div) ++s;
  VS_REQUIRE((1 << s) == c.row_div, "umma_pairfused: row_div=%d must be a power of two", c.row_div);
  p.row_div_shift = s;
  const uint32_t bar_bytes = 8u * 44 + 16u;     // a_land[8] go_a[8] round_done[8] go_m[8] w_land w_full, the TMEM slot, a_done[8]
  const uint32_t cap = 227u * 1024;
  const uint32_t n_mid = c.C == 64 ? 2u : 1u;                          // = LAG of the kernel
  const uint32_t fixed = 2u * p.w_bytes + n_mid * p.mid_bytes + bar_bytes + 256u;
  if (fixed + 2u * (uint32_t)p.n_chunks * p.slot_bytes > cap) return VS_ERR_INVALID;   // does not fit: the caller falls back (no error message)
  int units_a = (int)((cap - fixed) / (p.slot_bytes * (uint32_t)p.n_chunks));          // whole units of A in flight (>= 2)
  if (units_a * p.n_chunks > kMaxSlots) units_a = kMaxSlots / p.n_chunks;
  p.units_a = units_a;
  p.nslot = units_a * p.n_chunks;
  p.off_mid = (uint32_t)p.nslot * p.slot_bytes;
  p.off_w = p.off_mid + n_mid * p.mid_bytes;
  p.off_bar = (p.off_w + 2u * p.w_bytes + 127u) & ~127u;
  p.smem_bytes = p.off_bar + bar_bytes;
  if (p.smem_bytes < 120u * 1024) p.smem_bytes = 120u * 1024;       // one CTA per SM: it owns all 512 TMEM columns
  p.n_tiles = (c.R + p.V - 1) / p.V;
  p.n_units = (p.n_tiles + 1) / 2;
  *out = p;
  return VS_OK;
}

template <int C>
int launch(const Params& prm, int grid, cudaStream_t st) {
  const UmmaPair& c = prm.c;
  int mode = M_GENERIC;
  if (c.act_slope > 0.f) {
    if (c.out_act && !c.out_raw && !c.res2 && c.act_scale == 1.f) mode = M_ACT;
    else if (c.out_raw && !c.out_act && !c.res2) mode = M_RAW;
    else if (c.out_raw && !c.out_act && c.res2) mode = M_RAW_RES2;
    else if (c.out_act && !c.out_raw && c.res2) mode = M_ACT_RES2_SCALE;
  }
#define VS_PF_CASE(MD)                                                                                         \
  case MD: {                                                                                                   \
    VS_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(umma_pairfused_kernel<C, MD>), 227 * 1024));      \
    VS_CUDA_CHECK(launch_pdl<4>(umma_pairfused_kernel<C, MD>, dim3(grid), dim3(kThreads), prm.p.smem_bytes, st, prm));                               \
    break;                                                                                                     \
  }
  switch (mode) {
    VS_PF_CASE(M_ACT)
    VS_PF_CASE(M_RAW)
    VS_PF_CASE(M_RAW_RES2)
    VS_PF_CASE(M_ACT_RES2_SCALE)
    default:
      VS_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(umma_pairfused_kernel<C, M_GENERIC>), 227 * 1024));
      VS_CUDA_CHECK(launch_pdl<4>(umma_pairfused_kernel<C, M_GENERIC>, dim3(grid), dim3(kThreads), prm.p.smem_bytes, st, prm));
  }
#undef VS_PF_CASE
  VS_LAUNCH_CHECK();
  return VS_OK;
}

}  // namespace

bool umma_pairfused_supported(int C, int taps, int dil) {
  if (!(C == 64 || C == 128)) return false;
  UmmaPair c;
  c.C = C; c.taps = taps; c.dil = dil; c.R = 1024; c.row_div = 1;
  Plan p;
  if (!((taps & 1) && taps >= 3 && dil >= 1)) return false;
  return make_plan(c, &p) == VS_OK;
}

int umma_pairfused(const UmmaPair& c, cudaStream_t st) {
  VS_REQUIRE(c.x && c.w1 && c.w2 && (c.out_raw || c.out_act) && c.R > 0, "umma_pairfused: null pointer");
  VS_REQUIRE(c.in_slope > 0.f && c.in_slope <= 1.f && c.act_slope > 0.f && c.act_slope <= 1.f, "umma_pairfused: slopes must be in (0, 1]");
  Params prm;
  prm.c = c;
  prm.dbg = static_cast<long long*>(umma_conv_timing_buffer());
  VS_REQUIRE(make_plan(c, &prm.p) == VS_OK, "umma_pairfused: C=%d taps=%d dil=%d does not fit", c.C, c.taps, c.dil);
  for (int e = 0; e < 128; ++e) prm.bias[0][e] = prm.bias[1][e] = 0.f;
  if (c.b1_host && c.b2_host) {
    for (int e = 0; e < c.C; ++e) { prm.bias[0][e] = c.b1_host[e]; prm.bias[1][e] = c.b2_host[e]; }
  } else {   // op-level API (tests, tools): fetch the biases; the decoder passes host copies made at model finalize
    VS_REQUIRE(c.b1 && c.b2, "umma_pairfused: no biases");
    VS_CUDA_CHECK(cudaMemcpyAsync(prm.bias[0], c.b1, c.C * sizeof(float), cudaMemcpyDeviceToHost, st));
    VS_CUDA_CHECK(cudaMemcpyAsync(prm.bias[1], c.b2, c.C * sizeof(float), cudaMemcpyDeviceToHost, st));
    VS_CUDA_CHECK(cudaStreamSynchronize(st));
  }
  int n_sm = 0;
  VS_TRY(device_sm_count(&n_sm));
  int n_pairs = n_sm / 2;
  if (n_pairs > prm.p.n_units) n_pairs = prm.p.n_units;
  return c.C == 64 ? launch<64>(prm, 2 * n_pairs, st) : launch<128>(prm, 2 * n_pairs, st);
}

}  // namespace vs
