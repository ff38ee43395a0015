// The decoder's wide convs (Cin = Cout = 128, and 256 at k = 3) as a CTA-PAIR implicit GEMM: tcgen05.mma.cta_group::2,
// M = 256 rows per MMA (128 per CTA), N = Cout split between the two CTAs' shared memories.
//
// Why a pair (reference modules.py:210-223 convs1 / convs2 of the C = 128 stage: 18 launches, 7 of the decoder's 28 ms):
// as a single-CTA kernel these convs cannot keep their weights on chip (k = 7: 229 KB, k = 11: 360 KB), so every 128-row tile
// re-streams them out of L2 (64 B/clk/SM asked of L2 at k = 7) and the MMA issuer waits on weight slabs and on its only A stage
// (120 clk per MMA against 64).  In a pair each CTA holds the weights of ITS 64 output channels only - all of k = 11 is 176 KB -
// so the weights are RESIDENT for the whole kernel, each MMA reads 4 KB (A) + 2 KB (B half) of shared memory per SM (48 clk
// against 64 clk of tensor time), and what is left of shared memory is a ring of A K-chunks:
//   * A (planar f16 [C/8][R][8]) arrives as 32-channel chunks (4 plane slabs of 128 + halo rows, 4 bulk copies); the MMA loop is
//     chunk-major (chunk -> tap -> 2 x K16), so a chunk's slot is handed back after taps x 2 MMAs and the ring (3-8 slots) runs
//     about one tile ahead without ever holding a whole tile twice;
//   * only the leader CTA issues MMAs; the peer's operand arrivals reach the leader through a relay thread (wait on the local
//     TMA barrier, arrive on the leader's), slot releases and accumulator hand-offs come back by tcgen05.commit multicast, the
//     peer's epilogue warps release accumulators with remote arrives;
//   * TMEM: 4 accumulators of 128 columns per CTA (2 of 256): epilogues of unit i overlap the MMAs of units i+1 .. i+3.
// Epilogue: the decoder's c1 / c2 forms (bias, residual recovered from the activated stream, MRF running sum, leaky-ReLU, scale,
// validity mask), specialised at compile time like umma_conv.cu's.
#include "umma_conv.cuh"
#include "umma_common.cuh"

namespace vs {
namespace {

using namespace umma;
constexpr int kEpiWarps = 8;
constexpr int kThreads = (4 + kEpiWarps) * 32;     // producer | MMA issuer (leader) | relay | 8 epilogue warps | second MMA issuer
constexpr int kMaxSlots = 8;
constexpr int kMaxAcc = 4;
// channels per A chunk: 64 at C = 128 (fewer chunk hand-offs per unit); 32 at C = 256, whose 192 KB of resident k = 3 weights leave room for
// three small slots only
__host__ __device__ constexpr int kch_for(int c) { return c == 128 ? 64 : 32; }
constexpr int kPairM = 2 * kTileM;


struct Plan {
  int rows_a, halo_l, n_chunks, planes, nhalf, nslot, nacc, n_units, row_div_shift, n_issuers;
  uint32_t slot_bytes, w_bytes, off_w, off_bar, off_bias, smem_bytes;
};
struct Params {
  UmmaConv c;
  Plan p;
  long long* dbg;         // optional per-CTA wait-time counters (option "umma_timing_buffer"; builds with -DVS_UMMA_TIMING only)
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

constexpr int F_RES = 1, F_RES2 = 2, F_RAW = 8, F_ACT = 16, F_SCALE = 32, F_RESINV = 256;

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t addr, uint32_t rank) {     // shared::cta address -> shared::cluster address in CTA `rank`
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  // default semantics (.release.cta), as CUTLASS's ClusterBarrier::arrive(cta_id): a cluster-scope release costs ~1.6k clk per arrive
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
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
// wait on a barrier that the OTHER CTA of the pair arrives on (cluster-scope acquire); bounded like mbar_wait
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity, int tag) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  const long long t0 = clock64();
  for (;;) {
#pragma unroll 1
    for (int it = 0; it < 256; ++it)
      if (mbar_try_wait_cluster(bar, parity)) return;
    if (clock64() - t0 > 4000000000LL) {
      printf("umma_pair: mbarrier timeout tag=%d block=%d thread=%d parity=%u\n", tag, blockIdx.x, threadIdx.x, parity);
      __trap();
    }
  }
}
// one elected lane commits; the arrival is multicast to the same barrier of both CTAs of the pair
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

template <int F, int C>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) umma_pair_kernel(const __grid_constant__ Params prm) {
  constexpr int kKCH = kch_for(C), kChunkPlanes = kKCH / 8;
  constexpr int kMaxCC = C / 64;                      // 32-column chunks per epilogue warp (each warp owns N / 2 columns of its lane quarter)
  constexpr bool kPrefetchRes2 = C == 128;            // at C = 256 the MRF-sum rows are fetched chunk by chunk (registers)
  extern __shared__ __align__(128) uint8_t smem[];
  pdl_trigger();
  const UmmaConv& c = prm.c;
  const Plan& p = prm.p;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t rank = cluster_rank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
#ifdef VS_UMMA_TIMING
  long long* const dbg = prm.dbg;
  long long tw0 = 0, tw1 = 0, tw2 = 0;
  const long long t_start = dbg ? clock64() : 0;
#endif

  const uint32_t smem_base = smem_u32(smem);
  const uint32_t a_base = smem_base, w_base = smem_base + p.off_w, bar_base = smem_base + p.off_bar;
  // barriers (8 B each; the same offsets in both CTAs): a_land[8] a_full[8] a_empty[8] acc_full[4] acc_empty[4] w_land w_full
  auto a_land = [&](int i) { return bar_base + 8u * i; };                       // local: this CTA's chunk has landed
  auto a_full = [&](int i) { return bar_base + 8u * (kMaxSlots + i); };         // leader's: both CTAs' chunks have landed
  auto a_empty = [&](int i) { return bar_base + 8u * (2 * kMaxSlots + i); };    // local (commit multicast): the MMAs have read it
  auto acc_full = [&](int i) { return bar_base + 8u * (3 * kMaxSlots + i); };   // local (commit multicast)
  auto acc_empty = [&](int i) { return bar_base + 8u * (3 * kMaxSlots + kMaxAcc + i); };   // leader's: 16 epilogue warps
  const uint32_t w_land = bar_base + 8u * (3 * kMaxSlots + 2 * kMaxAcc), w_full = w_land + 8u;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + p.off_bar + 8 * (3 * kMaxSlots + 2 * kMaxAcc + 2));

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.nslot; ++i) { mbar_init(a_land(i), 1); mbar_init(a_full(i), 2); mbar_init(a_empty(i), 1); }
    for (int i = 0; i < p.nacc; ++i) { mbar_init(acc_full(i), 1); mbar_init(acc_empty(i), 2 * kEpiWarps); }
    mbar_init(w_land, 1);
    mbar_init(w_full, 2);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {      // the same warp of both CTAs allocates the pair's TMEM columns
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  float* bias_s = reinterpret_cast<float*>(smem + p.off_bias);
  for (int i = threadIdx.x; i < c.N; i += kThreads) bias_s[i] = c.bias ? c.bias[i] : 0.f;
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // the peer's barriers are initialised before anything arrives on them
  tc_fence_after();
  pdl_wait();                              // the prologue above overlapped the previous kernel's tail
  const uint32_t tmem_base = *tmem_slot;
  const int R = c.R, taps = c.taps, dil = c.dil;

  if (warp == 0) {
    // ------------------------------------------------------------------ producer: resident weights once, then the A chunk ring
    {
      const uint32_t plane_bytes = (uint32_t)p.nhalf * 16u;
      const int n_wslabs = taps * p.planes;
      if (lane == 0) mbar_arrive_expect_tx(w_land, p.w_bytes);
      __syncwarp();
      for (int sl = lane; sl < n_wslabs; sl += 32)
        bulk_g2s(w_base + (uint32_t)sl * plane_bytes, c.w + ((size_t)sl * c.N + (size_t)rank * p.nhalf) * 8, plane_bytes, w_land);
    }
    uint32_t slot = 0, ph = 0;
    for (int u = pair; u < p.n_units; u += n_pairs) {
      const int row_lo = u * kPairM + (int)rank * kTileM - p.halo_l, row_hi = row_lo + p.rows_a;
      const int c_lo = row_lo < 0 ? 0 : row_lo, c_hi = row_hi > R ? R : row_hi;
      const int n_zero_lo = c_lo - row_lo;
      const int n_rows = c_hi > c_lo ? c_hi - c_lo : 0;
      const int n_zero_hi = p.rows_a - n_zero_lo - n_rows;
      for (int ch = 0; ch < p.n_chunks; ++ch) {
        VS_TIMED(tw0, mbar_wait(a_empty(slot), ph ^ 1u, 1));
        const uint32_t stage = a_base + slot * p.slot_bytes;
        if (n_rows < p.rows_a) {       // rows outside [0, R): the conv's zero padding
          const int per_plane = p.rows_a - n_rows;
          for (int i = lane; i < kChunkPlanes * per_plane; i += 32) {
            const int pl = i / per_plane, j = i % per_plane;
            const int row = j < n_zero_lo ? j : (p.rows_a - n_zero_hi + (j - n_zero_lo));
            asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(stage + (uint32_t)(pl * p.rows_a + row) * 16u), "r"(0) : "memory");
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
          bulk_g2s(stage + (uint32_t)(lane * p.rows_a + n_zero_lo) * 16u, c.in + ((size_t)(ch * kChunkPlanes + lane) * R + c_lo) * 8, bytes,
                   a_land(slot));
        if (++slot == (uint32_t)p.nslot) { slot = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ relay: "my half has landed" -> the leader's barrier
    if (lane == 0) {
      mbar_wait(w_land, 0, 2);
      mbar_arrive_remote(map_to_cta(w_full, 0));
      uint32_t slot = 0, ph = 0;
      for (int u = pair; u < p.n_units; u += n_pairs)
        for (int ch = 0; ch < p.n_chunks; ++ch) {
          VS_TIMED(tw0, mbar_wait(a_land(slot), ph, 3));
          mbar_arrive_remote(map_to_cta(a_full(slot), 0));
          if (++slot == (uint32_t)p.nslot) { slot = 0; ph ^= 1u; }
        }
    }
  } else if (warp == 1 || warp == 3 + kEpiWarps) {
    // ------------------------------------------------------------------ MMA issuers (leader CTA only)
    if (rank == 0) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(kPairM >> 4) << 24);
      const uint32_t a_lbo = (uint32_t)p.rows_a * 16u, b_lbo = (uint32_t)p.nhalf * 16u;
      const uint32_t a_hi = (uint32_t)(make_desc(0, a_lbo, 128u) >> 32), b_hi = (uint32_t)(make_desc(0, b_lbo, 128u) >> 32);
      const uint32_t a_lo_fixed = (uint32_t)make_desc(0, a_lbo, 128u), b_lo_fixed = (uint32_t)make_desc(0, b_lbo, 128u);
      const uint32_t a_kstep = 2u * (uint32_t)p.rows_a, b_kstep = 2u * (uint32_t)p.nhalf;     // 16-byte units per K = 16 step
      const uint32_t b_tapstep = (uint32_t)p.planes * (uint32_t)p.nhalf;                     // 16-byte units per tap
      const uint32_t b_chunkstep = (uint32_t)kChunkPlanes * (uint32_t)p.nhalf;
      const uint32_t w16 = b_lo_fixed + (w_base >> 4);
      const uint32_t nslot = (uint32_t)p.nslot, nacc = (uint32_t)p.nacc, ncols = (uint32_t)c.N;
      const int n_chunks = p.n_chunks;
      mbar_wait_cluster(w_full, 0, 4);
      tc_fence_after();
      auto chunk_mmas = [&](uint32_t d_tmem, uint32_t slot, uint32_t b_chunk, uint32_t accumulate) {
        uint32_t a_tap = a_lo_fixed + ((a_base + slot * p.slot_bytes) >> 4), b_tap = b_chunk;
#pragma unroll 1
        for (int t = 0; t < taps; ++t, a_tap += (uint32_t)dil, b_tap += b_tapstep) {
          tc_mma_pair(d_tmem, a_tap, a_hi, b_tap, b_hi, idesc, accumulate);
          accumulate = 1;
#pragma unroll
          for (int k16 = 1; k16 < kKCH / 16; ++k16)
            tc_mma_pair(d_tmem, a_tap + (uint32_t)k16 * a_kstep, a_hi, b_tap + (uint32_t)k16 * b_kstep, b_hi, idesc, 1u);
        }
      };
      // A satisfied mbarrier probe still costs the issuer a few hundred clocks while the tensor pipe is busy, so the next chunk's
      // barrier is PROBED before the current chunk's MMAs are issued and the blocking wait is taken only if that probe failed.
      if (p.n_issuers == 2) {
        // Two issuing warps take alternate units (issuer iw: local units iw, iw + 2, ...): one warp's barrier round trips and
        // descriptor arithmetic overlap the other's MMAs (c2 forms of the C = 128 stage: -4 %).  Slots and accumulators are
        // functions of the unit index; the plan makes the ring a whole, even number of units, so consecutive users of one barrier
        // are always the same warp and its phases are consumed in order.
        const int iw = warp == 1 ? 0 : 1;
        const uint32_t units_in_ring = nslot / (uint32_t)n_chunks;
        for (int n = iw, u = pair + iw * n_pairs; u < p.n_units; n += 2, u += 2 * n_pairs) {
          const uint32_t acc_slot = (uint32_t)n % nacc, acc_phase = ((uint32_t)n / nacc) & 1u;
          uint32_t slot = ((uint32_t)n % units_in_ring) * (uint32_t)n_chunks;
          const uint32_t ph = ((uint32_t)n / units_in_ring) & 1u;
          VS_TIMED(tw1, mbar_wait_cluster(acc_empty(acc_slot), acc_phase ^ 1u, 5));
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc_slot * ncols;
          uint32_t b_chunk = w16;
          bool a_ready = false;
          for (int ch = 0; ch < n_chunks; ++ch, b_chunk += b_chunkstep, ++slot) {
            if (!a_ready) VS_TIMED(tw0, mbar_wait_cluster(a_full(slot), ph, 6));
            tc_fence_after();
            a_ready = ch + 1 < n_chunks ? mbar_test_wait(a_full(slot + 1), ph) : false;
            chunk_mmas(d_tmem, slot, b_chunk, ch ? 1u : 0u);
            VS_TIMED(tw2, tc_commit_pair(a_empty(slot)));        // both CTAs' producers may refill the slot once these MMAs have read it
          }
          VS_TIMED(tw2, tc_commit_pair(acc_full(acc_slot)));     // accumulator complete in both CTAs' TMEM -> both epilogues
        }
      } else if (warp == 1) {
        // one issuer, chunk-granular ring (C = 256: three small slots next to 192 KB of weights)
        uint32_t slot = 0, ph = 0, acc_slot = 0, acc_phase = 0;
        bool a_ready = false, acc_ready = false;
        for (int u = pair; u < p.n_units; u += n_pairs) {
          if (!acc_ready) VS_TIMED(tw1, mbar_wait_cluster(acc_empty(acc_slot), acc_phase ^ 1u, 5));
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc_slot * ncols;
          uint32_t next_acc = acc_slot + 1, next_acc_phase = acc_phase;
          if (next_acc == nacc) { next_acc = 0; next_acc_phase ^= 1u; }
          uint32_t b_chunk = w16;
          for (int ch = 0; ch < n_chunks; ++ch, b_chunk += b_chunkstep) {
            if (!a_ready) VS_TIMED(tw0, mbar_wait_cluster(a_full(slot), ph, 6));
            tc_fence_after();
            uint32_t nslot_i = slot + 1, nph = ph;
            if (nslot_i == nslot) { nslot_i = 0; nph ^= 1u; }
            a_ready = mbar_test_wait(a_full(nslot_i), nph);                     // consumed at the top of the next chunk
            if (ch == n_chunks - 1) acc_ready = mbar_test_wait(acc_empty(next_acc), next_acc_phase ^ 1u);
            chunk_mmas(d_tmem, slot, b_chunk, ch ? 1u : 0u);
            VS_TIMED(tw2, tc_commit_pair(a_empty(slot)));
            slot = nslot_i; ph = nph;
          }
          VS_TIMED(tw2, tc_commit_pair(acc_full(acc_slot)));
          acc_slot = next_acc; acc_phase = next_acc_phase;
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (warps 3..10): this CTA's 128 rows x N columns
    constexpr bool kGeneric = F < 0;
    const bool has_res = kGeneric ? (c.res != nullptr) : ((F & F_RES) != 0);
    const bool has_res2 = kGeneric ? (c.res2 != nullptr) : ((F & F_RES2) != 0);
    const bool has_raw = kGeneric ? (c.out_raw != nullptr) : ((F & F_RAW) != 0);
    const bool has_act = kGeneric ? (c.out_act != nullptr) : ((F & F_ACT) != 0);
    const bool has_scale = kGeneric ? (c.act_scale != 1.f) : ((F & F_SCALE) != 0);
    const bool res_inv = kGeneric ? (c.res_inv_slope != 0.f) : ((F & F_RESINV) != 0);
    const float rinv = c.res_inv_slope, slope = c.act_slope, scale = c.act_scale;
    const int q = warp & 3, hsel = (warp - 3) >> 2;
    const int half_cols = c.N / 2, n_cc = half_cols / 32;
    const uint32_t g8_0 = (uint32_t)(hsel * half_cols) >> 3;
    const size_t plane_stride = (size_t)R * 8;
    uint32_t acc_slot = 0, acc_phase = 0;
    for (int u = pair; u < p.n_units; u += n_pairs) {
      const int r = u * kPairM + (int)rank * kTileM + q * 32 + lane;
      const bool in_range = r < R;
      bool valid = in_range;
      if (in_range && c.row_utt) valid = c.row_utt[r >> p.row_div_shift] >= 0;
      const size_t row_elem = (size_t)r * 8;
      // the residual reads do not depend on the accumulator: they are in flight while this unit's MMAs still run
      uint4 rv[kMaxCC * 4], rv2[kPrefetchRes2 ? kMaxCC * 4 : 4];
      if (valid && (has_res || has_res2)) {
#pragma unroll
        for (int g = 0; g < kMaxCC * 4; ++g)
          if (g < n_cc * 4) {
            const size_t o = (size_t)(g8_0 + g) * plane_stride + row_elem;
            if (has_res) rv[g] = *reinterpret_cast<const uint4*>(c.res + o);
            if (has_res2 && kPrefetchRes2) rv2[g] = *reinterpret_cast<const uint4*>(c.res2 + o);
          }
      }
      VS_TIMED(tw0, mbar_wait(acc_full(acc_slot), acc_phase, 7));
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc_slot * (uint32_t)c.N + (uint32_t)(hsel * half_cols);
#pragma unroll
      for (int cc = 0; cc < kMaxCC; ++cc) {
        if (cc >= n_cc) break;
        const uint32_t g8 = g8_0 + (uint32_t)(cc * 4);
        if (has_res2 && !kPrefetchRes2 && valid) {
#pragma unroll
          for (int g = 0; g < 4; ++g) rv2[g] = *reinterpret_cast<const uint4*>(c.res2 + (size_t)(g8 + g) * plane_stride + row_elem);
        }
        uint32_t v[32];
        VS_TIMED(tw1, tmem_ld32(t_row + (uint32_t)(cc * 32), v));
        if (in_range) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const size_t o = (size_t)(g8 + g) * plane_stride + row_elem;
            uint4 raw = make_uint4(0, 0, 0, 0), act = make_uint4(0, 0, 0, 0);
            if (valid) {
              const uint32_t co0 = (g8 + (uint32_t)g) << 3;
              float y[8];
              const float4 b0 = *reinterpret_cast<const float4*>(bias_s + co0);
              const float4 b1 = *reinterpret_cast<const float4*>(bias_s + co0 + 4);
              y[0] = __uint_as_float(v[8 * g + 0]) + b0.x; y[1] = __uint_as_float(v[8 * g + 1]) + b0.y;
              y[2] = __uint_as_float(v[8 * g + 2]) + b0.z; y[3] = __uint_as_float(v[8 * g + 3]) + b0.w;
              y[4] = __uint_as_float(v[8 * g + 4]) + b1.x; y[5] = __uint_as_float(v[8 * g + 5]) + b1.y;
              y[6] = __uint_as_float(v[8 * g + 6]) + b1.z; y[7] = __uint_as_float(v[8 * g + 7]) + b1.w;
              if (has_res) {
                float f[8];
                unpack_f16x8(rv[cc * 4 + g], f);
#pragma unroll
                for (int e = 0; e < 8; ++e) y[e] += res_inv ? fminf(f[e], f[e] * rinv) : f[e];
              }
              if (has_res2) {
                float f[8];
                unpack_f16x8(rv2[kPrefetchRes2 ? cc * 4 + g : g], f);
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
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(map_to_cta(acc_empty(acc_slot), 0));
      if (++acc_slot == (uint32_t)p.nacc) { acc_slot = 0; acc_phase ^= 1u; }
    }
  }

#ifdef VS_UMMA_TIMING
  if (dbg && lane == 0 && warp < 4) {   // [cta][producer | MMA | relay | first epilogue warp][total, wait0, wait1, wait2]
    long long* o = dbg + ((size_t)blockIdx.x * 4 + warp) * 4;
    o[0] = clock64() - t_start; o[1] = tw0; o[2] = tw1; o[3] = tw2;
  }
#endif
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // nobody leaves while the other CTA may still arrive on its barriers or read its memory
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

int make_plan(const UmmaConv& c, Plan* out) {
  Plan p{};
  p.planes = c.Cin / 8;
  const int kch = kch_for(c.Cin);
  p.n_chunks = c.Cin / kch;
  p.nhalf = c.N / 2;
  p.halo_l = c.pad_l * c.dil;
  p.rows_a = kTileM + (c.taps - 1) * c.dil;
  p.slot_bytes = (uint32_t)(kch / 8) * (uint32_t)p.rows_a * 16u;
  p.w_bytes = (uint32_t)c.taps * (uint32_t)c.Cin * (uint32_t)p.nhalf * 2u;
  int s = 0;
  while ((1 << s) < c.row_div) ++s;
  VS_REQUIRE((1 << s) == c.row_div, "umma_pair: row_div=%d must be a power of two", c.row_div);
  p.row_div_shift = s;
  const uint32_t bar_bytes = 8u * (3 * kMaxSlots + 2 * kMaxAcc + 2) + 16u;
  const uint32_t fixed = bar_bytes + 256u + (uint32_t)c.N * 4u;
  const uint32_t cap = 227u * 1024;
  VS_REQUIRE(p.w_bytes + fixed + 2 * p.slot_bytes <= cap, "umma_pair: weights + 2 A chunks do not fit in shared memory");
  const int nslot = (int)((cap - fixed - p.w_bytes) / p.slot_bytes);
  // two issuing warps need a ring of a whole, even number of units (see the kernel); else one issuer and a chunk-granular ring
  const int units = nslot / p.n_chunks;
  p.n_issuers = units >= 2 ? 2 : 1;
  p.nslot = units >= 2 ? (units >= 4 ? 4 : 2) * p.n_chunks : (nslot > kMaxSlots ? kMaxSlots : nslot);
  VS_REQUIRE(p.nslot <= kMaxSlots, "umma_pair: A ring deeper than its barrier table");
  p.nacc = 512 / c.N > kMaxAcc ? kMaxAcc : 512 / c.N;
  p.off_w = (uint32_t)p.nslot * p.slot_bytes;
  p.off_bar = (p.off_w + p.w_bytes + 127u) & ~127u;
  p.off_bias = (p.off_bar + bar_bytes + 15u) & ~15u;
  p.smem_bytes = p.off_bias + (uint32_t)c.N * 4u;
  if (p.smem_bytes < 120u * 1024) p.smem_bytes = 120u * 1024;       // one CTA per SM: it owns all 512 TMEM columns
  p.n_units = (c.R + kPairM - 1) / kPairM;
  *out = p;
  return VS_OK;
}

}  // namespace

bool umma_pair_supported(const UmmaConv& c) {
  if (!(c.Cin == c.N && (c.Cin == 128 || (c.Cin == 256 && c.taps <= 3)))) return false;
  if (c.up != 1 || c.ubias || c.out_lo || c.taps < 1 || c.dil < 1) return false;
  const uint32_t w_bytes = (uint32_t)c.taps * c.Cin * (c.N / 2) * 2u;
  const uint32_t slot = (uint32_t)(kch_for(c.Cin) / 8) * (uint32_t)(kTileM + (c.taps - 1) * c.dil) * 16u;
  return w_bytes + 2 * slot + 2048u + (uint32_t)c.N * 4u <= 227u * 1024;
}

int umma_pair_conv(const UmmaConv& c, cudaStream_t st) {
  VS_REQUIRE(umma_pair_supported(c), "umma_pair: unsupported shape Cin=%d N=%d taps=%d", c.Cin, c.N, c.taps);
  VS_REQUIRE(c.in && c.w && (c.out_raw || c.out_act) && c.R > 0, "umma_pair: null pointer");
  VS_REQUIRE(c.act_slope > 0.f && c.act_slope <= 1.f, "umma_pair: act_slope must be in (0, 1]");
  Params prm;
  prm.c = c;
  prm.dbg = static_cast<long long*>(umma_conv_timing_buffer());
  VS_TRY(make_plan(c, &prm.p));
  int n_sm = 0;
  VS_TRY(device_sm_count(&n_sm));
  int n_pairs = n_sm / 2;
  {  // a persistent kernel must not ask for more clusters than can be co-resident (a GPC with an odd SM count strands one SM)
    static int cached[16] = {0};
    int dev = 0;
    VS_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 16 && cached[dev] == 0) {
      VS_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(umma_pair_kernel<-1, 128>), 227 * 1024));
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2 * n_pairs); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = 227 * 1024 - 1024;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, umma_pair_kernel<-1, 128>, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = n_pairs; }
      cached[dev] = n;
    }
    if (dev >= 0 && dev < 16 && cached[dev] < n_pairs) n_pairs = cached[dev];
  }
  if (n_pairs > prm.p.n_units) n_pairs = prm.p.n_units;
  const int flags = (c.res ? F_RES : 0) | (c.res2 ? F_RES2 : 0) | (c.out_raw ? F_RAW : 0) | (c.out_act ? F_ACT : 0) |
                    (c.act_scale != 1.f ? F_SCALE : 0) | ((c.res && c.res_inv_slope != 0.f) ? F_RESINV : 0);
#define VS_PAIR_CASE(FL)                                                                                  \
  case FL: {                                                                                              \
    if (c.Cin == 128) {                                                                                   \
      VS_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(umma_pair_kernel<FL, 128>), 227 * 1024));  \
      VS_CUDA_CHECK(launch_pdl<4>(umma_pair_kernel<FL, 128>, dim3(2 * n_pairs), dim3(kThreads), prm.p.smem_bytes, st, prm)); \
    } else {                                                                                              \
      VS_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(umma_pair_kernel<FL, 256>), 227 * 1024));  \
      VS_CUDA_CHECK(launch_pdl<4>(umma_pair_kernel<FL, 256>, dim3(2 * n_pairs), dim3(kThreads), prm.p.smem_bytes, st, prm)); \
    }                                                                                                     \
    break;                                                                                                \
  }
  switch (flags) {
    VS_PAIR_CASE(F_ACT)
    VS_PAIR_CASE(F_RES | F_RESINV | F_ACT)
    VS_PAIR_CASE(F_RES | F_RESINV | F_RAW)
    VS_PAIR_CASE(F_RES | F_RESINV | F_RES2 | F_RAW)
    VS_PAIR_CASE(F_RES | F_RESINV | F_RES2 | F_ACT | F_SCALE)
    default: {
      if (c.Cin == 128) {
        VS_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(umma_pair_kernel<-1, 128>), 227 * 1024));
        VS_CUDA_CHECK(launch_pdl<4>(umma_pair_kernel<-1, 128>, dim3(2 * n_pairs), dim3(kThreads), prm.p.smem_bytes, st, prm));
      } else {
        VS_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(umma_pair_kernel<-1, 256>), 227 * 1024));
        VS_CUDA_CHECK(launch_pdl<4>(umma_pair_kernel<-1, 256>, dim3(2 * n_pairs), dim3(kThreads), prm.p.smem_bytes, st, prm));
      }
    }
  }
#undef VS_PAIR_CASE
  VS_LAUNCH_CHECK();
  return VS_OK;
}

}  // namespace vs
