// conv1d as an implicit GEMM on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM), sm_100a.
//
//   D[128 rows x Nblk] (fp32, TMEM) = sum_{tap t} sum_{ci}  A_t[128 x Cin] (f16, smem) * W_t[Cin x Nblk] (f16, smem)
//
// Mapping (time on M, output channels on N):
//   * Activations live in HBM as planar f16 [C/8][R][8].  One plane-slab of a row tile (128 + halo rows x 16 B)
//     is contiguous in HBM *and* is exactly one K-chunk column of the UMMA "K-major, no swizzle" canonical
//     layout (8-row core matrices of 128 contiguous bytes, SBO = 128 B, LBO = slab pitch).  So the A tile is
//     fetched by Cin/8 plain TMA bulk copies (cp.async.bulk, completion on an mbarrier) with no tensor map, and
//     every conv tap is the SAME smem tile addressed with a start offset of tap*dil rows (16 B each):
//     the halo is loaded once, taps cost no extra traffic, and no im2col is ever materialised.
//   * Weights are pre-packed (packing.py) so that one (n-block, tap, 64-channel chunk) slab is one bulk copy that
//     lands in the same canonical layout; they stream through an SB-deep mbarrier ring out of L2.
//   * Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane) + TMEM owner,
//     warps 2..5 = epilogue (tcgen05.ld -> bias/residual/leaky-relu -> f16 -> coalesced 16 B planar stores).
//     TMEM holds two accumulator buffers so the epilogue of unit i overlaps the MMAs of unit i+1.
//   * Persistent CTAs stride over row tiles; for N > 256 (conv_pre, ConvTranspose phases) the n-blocks loop
//     inside the CTA so the A tile is fetched once.
//   * ConvTranspose1d runs as a polyphase conv: GEMM column gn = phase*Cout + co, taps = input offsets
//     {-1,0,+1}; the epilogue scatters row r, phase ph to output row up*r + ph.
#include "umma_conv.cuh"
#include "umma_common.cuh"

namespace vs {
namespace {

using namespace umma;
constexpr int kEpiWarps = 8;                 // two per TMEM lane quarter, splitting the column chunks
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kMaxStages = 8;
constexpr int kMaxAcc = 8;                   // TMEM accumulator ring depth (small N: many tiles in flight)

struct Plan {
  int rows_a, halo_l, planes, KC, n_kc, Nblk, NB, SA, SB, tmem_cols, n_tiles, Cout;
  int up_shift, row_div_shift, ctas_per_sm;
  int NACC;               // accumulator buffers in TMEM (ring between the MMA issuer and the epilogue)
  int MT;                 // row tiles (128 rows each) per A stage
  int resident_b;         // 1: every weight slab stays in smem for the whole kernel; 0: slabs stream through a ring
  int n_super;            // number of A stages' worth of work = ceil(n_tiles / MT)
  // Small calls (option "conv_spread"): a streamed-weight conv with fewer row tiles than SMs is cut along N as well - the n-block
  // shrinks from the packed width Nw to Nblk (a column range of each packed slab, fetched plane by plane) and one work item is
  // (A stage, NB / NS consecutive n-blocks), so a 3-tile ConvTranspose with N = 2048 runs on 96 CTAs instead of 3.
  int Nw, n_sub;          // packed n-block width (packing.py: min(N, 256)) and Nw / Nblk
  int NS, nb_per, n_items;   // n-block groups per A stage, n-blocks per group, work items = n_super * NS
  uint32_t w_bytes;       // all weight slabs
  uint32_t a_bytes, b_bytes, smem_bytes;
  uint32_t off_b, off_bar, off_bias;
};

struct Params {
  UmmaConv c;
  Plan p;
  long long* dbg;         // optional per-CTA wait-time counters (vs_set_option("umma_timing_buffer", device pointer))
};
// Wait-time counters are compiled in only with -DVS_UMMA_TIMING (VS_UMMA_TIMING=1 python vispeech_b200/build.py): even
// the dormant checks cost the streaming convs 5-8 %.
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

// Epilogue feature flags (template parameter F of the kernel; F < 0 = all decided at run time)
constexpr int F_RES = 1, F_RES2 = 2, F_UBIAS = 4, F_RAW = 8, F_ACT = 16, F_SCALE = 32, F_UP = 64, F_CW16 = 128, F_RESINV = 256, F_LO = 512;

template <int F>
__global__ void __launch_bounds__(kThreads, 2) umma_conv1d_kernel(const __grid_constant__ Params prm) {
  extern __shared__ __align__(128) uint8_t smem[];
  pdl_trigger();
  const UmmaConv& c = prm.c;
  const Plan& p = prm.p;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
#ifdef VS_UMMA_TIMING
  long long* const dbg = prm.dbg;
  long long tw0 = 0, tw1 = 0, tw2 = 0;
  const long long t_start = dbg ? clock64() : 0;
#endif

  const uint32_t smem_base = smem_u32(smem);
  const uint32_t a_base = smem_base;
  const uint32_t b_base = smem_base + p.off_b;
  const uint32_t bar_base = smem_base + p.off_bar;
  // barrier table (8 B each): a_full[8] a_empty[8] b_full[8] b_empty[8] acc_full[8] acc_empty[8]
  auto a_full = [&](int i) { return bar_base + 8u * i; };
  auto a_empty = [&](int i) { return bar_base + 8u * (kMaxStages + i); };
  auto b_full = [&](int i) { return bar_base + 8u * (2 * kMaxStages + i); };
  auto b_empty = [&](int i) { return bar_base + 8u * (3 * kMaxStages + i); };
  auto acc_full = [&](int i) { return bar_base + 8u * (4 * kMaxStages + i); };
  auto acc_empty = [&](int i) { return bar_base + 8u * (4 * kMaxStages + kMaxAcc + i); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + p.off_bar + 8 * (4 * kMaxStages + 2 * kMaxAcc));

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.SA; ++i) { mbar_init(a_full(i), 1); mbar_init(a_empty(i), 1); }
    for (int i = 0; i < (p.SB > 0 ? p.SB : 1); ++i) { mbar_init(b_full(i), 1); mbar_init(b_empty(i), 1); }
    for (int i = 0; i < p.NACC; ++i) { mbar_init(acc_full(i), 1); mbar_init(acc_empty(i), kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  float* bias_s = reinterpret_cast<float*>(smem + p.off_bias);     // [Cout] (zeros when the conv has no bias)
  for (int i = threadIdx.x; i < p.Cout; i += kThreads) bias_s[i] = c.bias ? c.bias[i] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                              // the prologue above overlapped the previous kernel's tail

  const int stages_per_unit = c.taps * p.n_kc;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    uint32_t a_it = 0, b_it = 0;
    const int n_slabs = p.NB * stages_per_unit;
    if (p.resident_b) {              // small convs: all weights fetched once, never streamed again
      if (lane == 0) mbar_arrive_expect_tx(b_full(0), p.w_bytes);
      __syncwarp();
      for (int sl = lane; sl < n_slabs; sl += 32)
        bulk_g2s(b_base + sl * p.b_bytes, reinterpret_cast<const uint8_t*>(c.w) + (size_t)sl * p.b_bytes, p.b_bytes,
                 b_full(0));
    }
    auto issue_a = [&](int item) {
      const int super = item / p.NS;
      const int sa = a_it % p.SA;
      const uint32_t ph = (a_it / p.SA) & 1;
      VS_TIMED(tw0, mbar_wait(a_empty(sa), ph ^ 1, 1));
      const int row_lo = super * p.MT * kTileM - p.halo_l, row_hi = row_lo + p.rows_a;
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
    for (int i = 0; i < look && next_a < p.n_items; ++i, next_a += gridDim.x) issue_a(next_a);
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      if (next_a < p.n_items) { issue_a(next_a); next_a += gridDim.x; }
      if (!p.resident_b && p.n_sub > 1) {
        // narrowed n-blocks: columns [sub * Nblk, +Nblk) of packed slab (nb / n_sub, stage), one copy per 8-channel plane; the planes go
        // out from different lanes (one thread issuing 8 small copies per slab was the whole kernel's pace at k = 11)
        const int super = item / p.NS, grp = item - super * p.NS;
        const int tiles_here = min(p.MT, p.n_tiles - super * p.MT);
        const int sl_lo = grp * p.nb_per * stages_per_unit, sl_hi = sl_lo + p.nb_per * stages_per_unit;
        const uint32_t plane_bytes = (uint32_t)p.Nblk * 16u;
        for (int m = 0; m < tiles_here; ++m)
          for (int sl = sl_lo; sl < sl_hi; ++sl) {
            const int sb = b_it % p.SB;
            const uint32_t ph = (b_it / p.SB) & 1;
            if (lane == 0) {
              VS_TIMED(tw1, mbar_wait(b_empty(sb), ph ^ 1, 2));
              mbar_arrive_expect_tx(b_full(sb), p.b_bytes);
            }
            __syncwarp();
            // packed [NB][taps][Cin / 8 planes][Nw][8]: the plane pitch is Nw * 16 bytes across the whole of Cin, so a stage may be
            // any run of planes (KC here is not packing.py's 64)
            const int nb = sl / stages_per_unit, stg = sl - nb * stages_per_unit;
            const int nb_w = nb / p.n_sub, sub = nb - nb_w * p.n_sub;
            const int t = stg / p.n_kc, kc = stg - t * p.n_kc;
            const uint8_t* src = reinterpret_cast<const uint8_t*>(c.w) +
                                 ((size_t)(nb_w * c.taps + t) * p.planes + (size_t)kc * (p.KC / 8) + lane) * ((size_t)p.Nw * 16u) +
                                 (size_t)sub * plane_bytes;
            if (lane < p.KC / 8) bulk_g2s(b_base + sb * p.b_bytes + lane * plane_bytes, src, plane_bytes, b_full(sb));
            ++b_it;
          }
      } else if (!p.resident_b && lane == 0) {
        const int super = item / p.NS, grp = item - super * p.NS;
        const int tiles_here = min(p.MT, p.n_tiles - super * p.MT);
        const int sl_lo = grp * p.nb_per * stages_per_unit, sl_hi = sl_lo + p.nb_per * stages_per_unit;
        for (int m = 0; m < tiles_here; ++m)
          for (int sl = sl_lo; sl < sl_hi; ++sl) {
            const int sb = b_it % p.SB;
            const uint32_t ph = (b_it / p.SB) & 1;
            VS_TIMED(tw1, mbar_wait(b_empty(sb), ph ^ 1, 2));
            mbar_arrive_expect_tx(b_full(sb), p.b_bytes);
            bulk_g2s(b_base + sb * p.b_bytes, reinterpret_cast<const uint8_t*>(c.w) + (size_t)sl * p.b_bytes, p.b_bytes,
                     b_full(sb));
            ++b_it;
          }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (warp-uniform loop, elected lane issues)
    {
      const uint32_t idesc = make_idesc(p.Nblk);
      const uint32_t a_lbo = (uint32_t)p.rows_a * 16u, b_lbo = (uint32_t)p.Nblk * 16u;
      // loop-invariant descriptor pieces (all in 16-byte units): see make_desc()
      const uint32_t a_hi = (uint32_t)(make_desc(0, a_lbo, 128u) >> 32), b_hi = (uint32_t)(make_desc(0, b_lbo, 128u) >> 32);
      const uint32_t a_lo_fixed = (uint32_t)make_desc(0, a_lbo, 128u), b_lo_fixed = (uint32_t)make_desc(0, b_lbo, 128u);
      const uint32_t a_kstep = 2u * (uint32_t)p.rows_a;          // two planes per K=16 step
      const uint32_t b_kstep = 2u * (uint32_t)p.Nblk;
      const int taps = c.taps, dil = c.dil, n_kc = p.n_kc, k16s = p.KC / 16, MT = p.MT, NACC = p.NACC, SB = p.SB, SA = p.SA;
      const int n_tiles = p.n_tiles, n_items = p.n_items, NS = p.NS, nb_per = p.nb_per;
      const bool resident = p.resident_b != 0;
      const int nks = n_kc * k16s;
      const bool lean = resident && p.NB == 1 && (nks == 2 || nks == 4 || nks == 8);
      const uint32_t slab16 = p.b_bytes >> 4, a_bytes = p.a_bytes, nblk = (uint32_t)p.Nblk;
      const uint32_t b_base16 = b_base >> 4;
      uint32_t acc_slot = 0, acc_phase = 0, b_slot = 0, b_phase = 0, a_slot = 0, a_phase = 0;
      if (resident) { mbar_wait(b_full(0), 0, 7); tc_fence_after(); }
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int super = item / NS, nb_lo = (item - super * NS) * nb_per, nb_hi = nb_lo + nb_per;
        VS_TIMED(tw0, mbar_wait(a_full(a_slot), a_phase, 3));
        tc_fence_after();
        const uint32_t a_stage16 = (a_base + a_slot * a_bytes) >> 4;
        const int tiles_here = min(MT, n_tiles - super * MT);
        if (lean) {
          // Resident weights, one n-block: the issue loop itself is the bottleneck for the small-C convs (one warp
          // issues every MMA; tools/mma_microbench.cu: nested runtime loops cost 55-120 clk per MMA against a
          // 40-48 clk operand-fetch floor).  So the K-chunk walk of a tap is fully unrolled (NK = Cin/16 as a template
          // parameter) and there is nothing in the loop but descriptor adds and the MMA.
          for (int m = 0; m < tiles_here; ++m) {
            VS_TIMED(tw1, mbar_wait(acc_empty(acc_slot), acc_phase ^ 1, 4));
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc_slot * nblk;
            const uint32_t a_tile = a_lo_fixed + a_stage16 + (uint32_t)(m * kTileM), b0 = b_lo_fixed + b_base16;
            if (nks == 2) issue_tile<2>(d_tmem, a_tile, a_hi, b0, b_hi, idesc, taps, (uint32_t)dil, a_kstep, b_kstep);
            else if (nks == 4) issue_tile<4>(d_tmem, a_tile, a_hi, b0, b_hi, idesc, taps, (uint32_t)dil, a_kstep, b_kstep);
            else issue_tile<8>(d_tmem, a_tile, a_hi, b0, b_hi, idesc, taps, (uint32_t)dil, a_kstep, b_kstep);
            tc_commit(acc_full(acc_slot));
            if (++acc_slot == (uint32_t)NACC) { acc_slot = 0; acc_phase ^= 1; }
          }
        } else {
          for (int m = 0; m < tiles_here; ++m)
            for (int nb = nb_lo; nb < nb_hi; ++nb) {
              VS_TIMED(tw1, mbar_wait(acc_empty(acc_slot), acc_phase ^ 1, 4));
              tc_fence_after();
              const uint32_t d_tmem = tmem_base + acc_slot * nblk;
              uint32_t accumulate = 0;
              uint32_t b_lo = b_lo_fixed + b_base16 + (uint32_t)(nb * taps * n_kc) * slab16;   // resident: walks all slabs
              uint32_t a_tap = a_lo_fixed + a_stage16 + (uint32_t)(m * kTileM);
              for (int t = 0; t < taps; ++t, a_tap += (uint32_t)dil) {
                uint32_t a_lo = a_tap;
                for (int kc = 0; kc < n_kc; ++kc) {
                  if (!resident) {
                    VS_TIMED(tw2, mbar_wait(b_full(b_slot), b_phase, 5));
                    tc_fence_after();
                    b_lo = b_lo_fixed + b_base16 + b_slot * slab16;
                  }
#pragma unroll 4
                  for (int k16 = 0; k16 < k16s; ++k16) {
                    tc_mma_f16_lohi(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, accumulate);
                    accumulate = 1;
                    a_lo += a_kstep;
                    b_lo += b_kstep;
                  }
                  if (!resident) {
                    tc_commit(b_empty(b_slot));    // frees the weight slot once these MMAs have read it
                    if (++b_slot == (uint32_t)SB) { b_slot = 0; b_phase ^= 1; }
                  }
                }
              }
              tc_commit(acc_full(acc_slot));       // accumulator complete -> epilogue
              if (++acc_slot == (uint32_t)NACC) { acc_slot = 0; acc_phase ^= 1; }
            }
        }
        tc_commit(a_empty(a_slot));              // every MMA of the stage has read the A tile
        if (++a_slot == (uint32_t)SA) { a_slot = 0; a_phase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9)
    // Warp w may touch TMEM lanes [32*(w%4), +32).  Two warps share each lane quarter and take alternate
    // column chunks, so 8 warps drain one accumulator.  The small-C convs are bound by the SM's instruction issue
    // rate in this loop (ncu: 2.5 of 4 IPC, 60 % of it here), so the feature set F is a template parameter and dead
    // paths vanish: per 8-channel group the specialised code is ~2 LDS (bias) + 8 FADD + 16 FMUL/FMNMX + 4 F2FP
    // + 1-2 STG.128 (+ LDG.128 and 16 unpack/add per residual).
    constexpr bool kGeneric = F < 0;
    const bool has_res = kGeneric ? (c.res != nullptr) : ((F & F_RES) != 0);
    const bool has_res2 = kGeneric ? (c.res2 != nullptr) : ((F & F_RES2) != 0);
    const bool has_ub = kGeneric ? (c.ubias != nullptr) : ((F & F_UBIAS) != 0);
    const bool has_raw = kGeneric ? (c.out_raw != nullptr) : ((F & F_RAW) != 0);
    const bool has_act = kGeneric ? (c.out_act != nullptr) : ((F & F_ACT) != 0);
    const bool has_scale = kGeneric ? (c.act_scale != 1.f) : ((F & F_SCALE) != 0);
    const bool has_up = kGeneric ? (c.up != 1) : ((F & F_UP) != 0);
    const bool res_inv = kGeneric ? (c.res_inv_slope != 0.f) : ((F & F_RESINV) != 0);
    const bool has_lo = kGeneric ? (c.out_lo != nullptr) : ((F & F_LO) != 0);
    const float rinv = c.res_inv_slope;
    const int q = warp & 3;
    const int hsel = (warp - 2) >> 2;
    const int CW = (kGeneric ? (p.Nblk >= 64) : ((F & F_CW16) == 0)) ? 32 : 16;   // columns per tcgen05.ld
    const int n_chunks = p.Nblk / CW;
    uint32_t acc_slot = 0, acc_phase = 0;
    const int R_out = c.R * c.up;
    const uint32_t up_mask = (uint32_t)c.up - 1u;
    const size_t plane_stride = (size_t)R_out * 8;          // elements between consecutive 8-channel planes
    const float slope = c.act_slope, scale = c.act_scale;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x)
    for (int super = item / p.NS, nb_lo = (item - super * p.NS) * p.nb_per, tile = super * p.MT; tile < min(p.n_tiles, (super + 1) * p.MT); ++tile) {
      const int r = tile * kTileM + q * 32 + lane;
      const bool in_range = r < c.R;
      int utt = -1;
      if (in_range) utt = c.row_utt ? c.row_utt[(c.up * r) >> p.row_div_shift] : 0;
      const bool valid = utt >= 0;
      const float* ub = nullptr;
      if (has_ub && valid) ub = c.ubias + (size_t)(c.ubias_idx ? c.ubias_idx[utt] : utt) * c.N;
      const size_t row_elem = (size_t)c.up * r * 8;         // element offset of this thread's (first) output row in a plane
      for (int nb = nb_lo; nb < nb_lo + p.nb_per; ++nb) {
        const int ab = (int)acc_slot;
        VS_TIMED(tw0, mbar_wait(acc_full(ab), acc_phase, 6));
        tc_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * p.Nblk);
        for (int cc = hsel; cc < n_chunks; cc += 2) {
          const int col0 = cc * CW;
          const uint32_t g8 = (uint32_t)(nb * p.Nblk + col0) >> 3;        // index of the chunk's first 8-column group
          // element offset of group g: plane * plane_stride + (up*r + phase) * 8
          auto group_off = [&](int g) -> size_t {
            const uint32_t gg = g8 + (uint32_t)g;
            if (has_up) return (size_t)(gg >> p.up_shift) * plane_stride + row_elem + (size_t)(gg & up_mask) * 8;
            return (size_t)gg * plane_stride + row_elem;
          };
          uint4 rv[4], rv2[4];
          if (valid && (has_res || has_res2)) {
#pragma unroll
            for (int g = 0; g < 4; ++g)
              if (g * 8 < CW) {
                const size_t o = group_off(g);
                if (has_res) rv[g] = *reinterpret_cast<const uint4*>(c.res + o);
                if (has_res2) rv2[g] = *reinterpret_cast<const uint4*>(c.res2 + o);
              }
          }
          uint32_t v[32];
          if (CW == 32) tmem_ld32(t_row + (uint32_t)col0, v);
          else tmem_ld16(t_row + (uint32_t)col0, v);
          if (in_range) {
#pragma unroll
            for (int g = 0; g < 4; ++g)
              if (g * 8 < CW) {
                const size_t o = group_off(g);
                uint4 raw = make_uint4(0, 0, 0, 0), act = make_uint4(0, 0, 0, 0), low = make_uint4(0, 0, 0, 0);
                if (valid) {
                  const uint32_t co0 = ((g8 + (uint32_t)g) >> p.up_shift) << 3;
                  float y[8];
                  const float4 b0 = *reinterpret_cast<const float4*>(bias_s + co0);
                  const float4 b1 = *reinterpret_cast<const float4*>(bias_s + co0 + 4);
                  y[0] = __uint_as_float(v[8 * g + 0]) + b0.x; y[1] = __uint_as_float(v[8 * g + 1]) + b0.y;
                  y[2] = __uint_as_float(v[8 * g + 2]) + b0.z; y[3] = __uint_as_float(v[8 * g + 3]) + b0.w;
                  y[4] = __uint_as_float(v[8 * g + 4]) + b1.x; y[5] = __uint_as_float(v[8 * g + 5]) + b1.y;
                  y[6] = __uint_as_float(v[8 * g + 6]) + b1.z; y[7] = __uint_as_float(v[8 * g + 7]) + b1.w;
                  if (has_ub) {
                    const float* u8 = ub + (size_t)(g8 + (uint32_t)g) * 8;
#pragma unroll
                    for (int e = 0; e < 8; ++e) y[e] += __ldg(u8 + e);
                  }
                  if (has_res) {
                    float f[8];
                    unpack_f16x8(rv[g], f);
#pragma unroll
                    for (int e = 0; e < 8; ++e) y[e] += res_inv ? fminf(f[e], f[e] * rinv) : f[e];
                  }
                  if (has_res2) {
                    float f[8];
                    unpack_f16x8(rv2[g], f);
#pragma unroll
                    for (int e = 0; e < 8; ++e) y[e] += f[e];
                  }
                  if (has_raw)
                    raw = make_uint4(pack_f16x2(y[0], y[1]), pack_f16x2(y[2], y[3]), pack_f16x2(y[4], y[5]),
                                     pack_f16x2(y[6], y[7]));
                  if (has_lo) {                            // y = fp16(y) + fp16(y - fp16(y)): 22 bits of y in two planar tensors
                    float h[8];
                    unpack_f16x8(raw, h);
                    low = make_uint4(pack_f16x2(y[0] - h[0], y[1] - h[1]), pack_f16x2(y[2] - h[2], y[3] - h[3]),
                                     pack_f16x2(y[4] - h[4], y[5] - h[5]), pack_f16x2(y[6] - h[6], y[7] - h[7]));
                  }
                  if (has_act) {
                    float z[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {          // leaky_relu(v) = max(v, slope*v) for 0 < slope <= 1
                      const float t = has_scale ? y[e] * scale : y[e];
                      z[e] = fmaxf(t, t * slope);
                    }
                    act = make_uint4(pack_f16x2(z[0], z[1]), pack_f16x2(z[2], z[3]), pack_f16x2(z[4], z[5]),
                                     pack_f16x2(z[6], z[7]));
                  }
                }
                if (has_raw) *reinterpret_cast<uint4*>(c.out_raw + o) = raw;
                if (has_lo) *reinterpret_cast<uint4*>(c.out_lo + o) = low;
                if (has_act) *reinterpret_cast<uint4*>(c.out_act + o) = act;
              }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty(ab));
        if (++acc_slot == (uint32_t)p.NACC) { acc_slot = 0; acc_phase ^= 1; }
      }
    }
  }

#ifdef VS_UMMA_TIMING
  if (dbg && lane == 0 && warp < 3) {   // [cta][warp 0 producer | 1 MMA | 2 first epilogue warp][total, wait0, wait1, wait2]
    long long* o = dbg + ((size_t)blockIdx.x * 3 + warp) * 4;
    o[0] = clock64() - t_start; o[1] = tw0; o[2] = tw1; o[3] = tw2;
  }
#endif
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

int make_plan(const UmmaConv& c, int n_sm, Plan* out) {
  Plan p{};
  VS_REQUIRE(c.Cin % 16 == 0 && c.Cin >= 16, "umma_conv1d: Cin=%d must be a multiple of 16", c.Cin);
  VS_REQUIRE(c.N % 32 == 0, "umma_conv1d: N=%d must be a multiple of 32", c.N);
  VS_REQUIRE(c.up >= 1 && c.N % c.up == 0 && (c.N / c.up) % 8 == 0, "umma_conv1d: bad upsample factor");
  VS_REQUIRE(c.R > 0 && c.taps >= 1 && c.dil >= 1 && c.pad_l >= 0, "umma_conv1d: bad shape");
  p.Cout = c.N / c.up;
  p.Nw = c.N < 256 ? c.N : 256;
  VS_REQUIRE(c.N % p.Nw == 0, "umma_conv1d: N=%d not a multiple of the 256-column block", c.N);
  p.Nblk = p.Nw;
  p.n_tiles = (c.R + kTileM - 1) / kTileM;
  // small call with streamed weights: narrow the n-block until the (row tile, n-block) units fill the SMs (never below 64 columns:
  // an N = 32 MMA costs what an N = 64 one does)
  const bool streamed = (uint64_t)c.Cin * c.N * c.taps * 2u > 96u * 1024;
  const bool spread = streamed && opts().v[OPT_CONV_SPREAD] && p.n_tiles * (c.N / p.Nw) * 2 <= n_sm;
  if (spread)
    for (int cand = 64; cand < p.Nw; cand *= 2)
      if (p.Nw % cand == 0 && p.n_tiles * (c.N / cand) <= n_sm) { p.Nblk = cand; break; }
  p.n_sub = p.Nw / p.Nblk;
  p.NB = c.N / p.Nblk;
  p.KC = c.Cin < 64 ? c.Cin : 64;
  if (p.n_sub > 1) {        // narrowed slabs are fetched plane by plane anyway: make them deep instead (<= 32 planes = one per lane, <= 32 KB),
    int kc = 32768 / (p.Nblk * 2);   // so that a k = 11 item is 11 fetches in flight at once and not 44 small ones behind a 4-slot ring
    if (kc > c.Cin) kc = c.Cin;
    if (kc > 256) kc = 256;
    while (c.Cin % kc) kc -= 16;
    p.KC = kc;
  }
  VS_REQUIRE(c.Cin % p.KC == 0, "umma_conv1d: Cin=%d not a multiple of the %d-channel chunk", c.Cin, p.KC);
  p.n_kc = c.Cin / p.KC;
  p.planes = c.Cin / 8;
  p.halo_l = c.pad_l * c.dil;
  const int halo = (c.taps - 1) * c.dil;
  p.b_bytes = (uint32_t)p.KC * p.Nblk * 2u;
  p.w_bytes = p.b_bytes * (uint32_t)(p.NB * c.taps * p.n_kc);
  auto log2_exact = [](int v) { int s = 0; while ((1 << s) < v) ++s; return (1 << s) == v ? s : -1; };
  p.up_shift = log2_exact(c.up);
  p.row_div_shift = log2_exact(c.row_div);
  VS_REQUIRE(p.up_shift >= 0 && p.row_div_shift >= 0, "umma_conv1d: up=%d and row_div=%d must be powers of two", c.up,
             c.row_div);
  const uint32_t bar_bytes = 8u * (4 * kMaxStages + 2 * kMaxAcc) + 16u;
  const uint32_t fixed = bar_bytes + 128u + (uint32_t)p.Cout * 4u;
  // Shared-memory policy.
  //  * Small convs (all weight slabs <= 96 KB): weights stay resident, the A stage spans MT row tiles so that
  //    activation fetches are few and large, two A stages.  Preferred footprint <= half an SM so two CTAs co-reside.
  //  * Large convs: one row tile per A stage, weights stream through a ring of >= 4 slabs out of L2.
  const uint32_t half_sm = 110u * 1024, full_sm = 220u * 1024;
  auto a_bytes_for = [&](int mt) { return (uint32_t)p.planes * (uint32_t)(kTileM * mt + halo) * 16u; };
  p.resident_b = p.w_bytes <= 96u * 1024 ? 1 : 0;
  uint32_t cap = full_sm;
  if (p.resident_b) {
    p.SB = 0;
    p.SA = 2;
    p.MT = 0;
    const int mts[3] = {4, 2, 1};
    for (int pass = 0; pass < 2 && !p.MT; ++pass)
      for (int i = 0; i < 3 && !p.MT; ++i)
        if (2 * a_bytes_for(mts[i]) + p.w_bytes + fixed <= (pass == 0 ? half_sm : full_sm)) p.MT = mts[i];
    if (!p.MT) { p.MT = 1; p.SA = 1; }
    VS_REQUIRE(p.SA * a_bytes_for(p.MT) + p.w_bytes + fixed <= full_sm, "umma_conv1d: tile does not fit in shared memory");
  } else {
    p.MT = 1;
    const uint32_t a1 = a_bytes_for(1);
    auto fits = [&](int sa, int sb, uint32_t lim) { return sa * a1 + sb * p.b_bytes + fixed <= lim; };
    if (p.n_sub > 1) { p.SA = 1; cap = full_sm; }          // one item per CTA: the whole SM for the weight ring
    else if (fits(2, 4, half_sm)) { p.SA = 2; cap = half_sm; }
    else if (fits(1, 4, half_sm)) { p.SA = 1; cap = half_sm; }
    else if (fits(2, 3, full_sm)) { p.SA = 2; cap = full_sm; }
    else { p.SA = 1; cap = full_sm; }
    VS_REQUIRE(fits(p.SA, 2, cap), "umma_conv1d: tile does not fit in shared memory");
    const int sb = (int)((cap - fixed - p.SA * a1) / p.b_bytes);
    const int sb_max = cap == half_sm ? 6 : kMaxStages;
    p.SB = sb > sb_max ? sb_max : sb;
  }
  p.rows_a = kTileM * p.MT + halo;
  p.a_bytes = a_bytes_for(p.MT);
  VS_REQUIRE(p.rows_a * 16 < (1 << 18), "umma_conv1d: halo too large for the descriptor pitch");
  p.off_b = p.SA * p.a_bytes;
  p.off_bar = (p.off_b + (p.resident_b ? p.w_bytes : p.SB * p.b_bytes) + 127u) & ~127u;
  p.off_bias = (p.off_bar + bar_bytes + 15u) & ~15u;
  p.smem_bytes = p.off_bias + (uint32_t)p.Cout * 4u;
  int per_sm = (int)((227u * 1024) / (p.smem_bytes + 1024));
  if (per_sm > 2) per_sm = 2;
  if (per_sm < 1) per_sm = 1;
  // TMEM: 512 columns per SM shared by the co-resident CTAs; as many accumulator buffers as fit (>= 2, <= 8)
  int nacc = (512 / per_sm) / p.Nblk;
  if (nacc < 2) { per_sm = 1; nacc = 512 / p.Nblk; }
  p.NACC = nacc > kMaxAcc ? kMaxAcc : nacc;
  int cols = 32;
  while (cols < p.NACC * p.Nblk) cols *= 2;
  p.tmem_cols = cols;
  // request enough shared memory that no more than per_sm CTAs can ever share an SM (their TMEM would not fit)
  const uint32_t min_smem = (227u * 1024) / (uint32_t)(per_sm + 1) + 1024u;
  if (p.smem_bytes < min_smem) p.smem_bytes = min_smem;
  p.ctas_per_sm = per_sm;
  p.n_super = (p.n_tiles + p.MT - 1) / p.MT;
  p.NS = 1;
  if (spread)                                   // as many n-block groups per A stage as keep the items within one wave
    for (int ns = p.NB; ns >= 1; --ns)
      if (p.NB % ns == 0 && p.n_super * ns <= n_sm) { p.NS = ns; break; }
  p.nb_per = p.NB / p.NS;
  p.n_items = p.n_super * p.NS;
  VS_REQUIRE(p.n_sub == 1 || !p.resident_b, "umma_conv1d: internal: narrowed n-block with resident weights");
  *out = p;
  return VS_OK;
}

}  // namespace

void* umma_conv_timing_buffer() { return reinterpret_cast<void*>(opts().v[OPT_TIMING_BUFFER]); }

int umma_conv1d(const UmmaConv& c, cudaStream_t st) {
  if (opts().v[OPT_PAIR_CONV] && umma_pair_supported(c) && (opts().v[OPT_PAIR_CONV] > 1 || c.Cin == 128)) return umma_pair_conv(c, st);   // 1: C = 128 only
  Params prm;
  prm.c = c;
  prm.dbg = static_cast<long long*>(umma_conv_timing_buffer());
  int n_sm = 0;
  VS_TRY(device_sm_count(&n_sm));
  VS_TRY(make_plan(c, n_sm, &prm.p));
  VS_REQUIRE(c.in && c.w && (c.out_raw || c.out_act), "umma_conv1d: null pointer");
  VS_REQUIRE(c.act_slope > 0.f && c.act_slope <= 1.f, "umma_conv1d: act_slope must be in (0, 1]");
  const int per_sm = prm.p.ctas_per_sm;
  int grid = n_sm * per_sm;
  if (grid > prm.p.n_items) grid = prm.p.n_items;
  int flags = (c.res ? F_RES : 0) | (c.res2 ? F_RES2 : 0) | (c.ubias ? F_UBIAS : 0) | (c.out_raw ? F_RAW : 0) |
              (c.out_act ? F_ACT : 0) | (c.act_scale != 1.f ? F_SCALE : 0) | (c.up != 1 ? F_UP : 0) |
              (prm.p.Nblk < 64 ? F_CW16 : 0) | ((c.res && c.res_inv_slope != 0.f) ? F_RESINV : 0) | (c.out_lo ? F_LO : 0);
  VS_REQUIRE(!c.out_lo || c.out_raw, "umma_conv1d: out_lo needs out_raw");
#define VS_UMMA_CASE(FL)                                                                                              \
  case FL: {                                                                                                          \
    VS_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(umma_conv1d_kernel<FL>), 227 * 1024));                   \
    VS_CUDA_CHECK(launch_pdl<4>(umma_conv1d_kernel<FL>, dim3(grid), dim3(kThreads), prm.p.smem_bytes, st, prm));                                            \
    break;                                                                                                            \
  }
  switch (flags) {
    // the decoder's epilogues: c1 | c2 mid | c2 last (j<2, j=2) | conv_pre | ups (raw+act, raw) ; each also at N = 32
    VS_UMMA_CASE(F_ACT)
    VS_UMMA_CASE(F_ACT | F_CW16)
    VS_UMMA_CASE(F_RES | F_RAW | F_ACT)
    VS_UMMA_CASE(F_RES | F_RAW | F_ACT | F_CW16)
    VS_UMMA_CASE(F_RES | F_RAW)
    VS_UMMA_CASE(F_RES | F_RAW | F_CW16)
    VS_UMMA_CASE(F_RES | F_RES2 | F_RAW)
    VS_UMMA_CASE(F_RES | F_RES2 | F_RAW | F_CW16)
    VS_UMMA_CASE(F_RES | F_RES2 | F_ACT | F_SCALE)
    VS_UMMA_CASE(F_RES | F_RES2 | F_ACT | F_SCALE | F_CW16)
    VS_UMMA_CASE(F_RES | F_RESINV | F_ACT)
    VS_UMMA_CASE(F_RES | F_RESINV | F_ACT | F_CW16)
    VS_UMMA_CASE(F_RES | F_RESINV | F_RAW)
    VS_UMMA_CASE(F_RES | F_RESINV | F_RAW | F_CW16)
    VS_UMMA_CASE(F_RES | F_RESINV | F_RES2 | F_RAW)
    VS_UMMA_CASE(F_RES | F_RESINV | F_RES2 | F_RAW | F_CW16)
    VS_UMMA_CASE(F_RES | F_RESINV | F_RES2 | F_ACT | F_SCALE)
    VS_UMMA_CASE(F_RES | F_RESINV | F_RES2 | F_ACT | F_SCALE | F_CW16)
    VS_UMMA_CASE(F_ACT | F_UP)
    VS_UMMA_CASE(F_UBIAS | F_ACT)
    VS_UMMA_CASE(F_RAW | F_ACT | F_UP)
    VS_UMMA_CASE(F_RAW | F_UP)
    VS_UMMA_CASE(F_RAW | F_UP | F_LO)
    default: {
      VS_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(umma_conv1d_kernel<-1>), 227 * 1024));
      VS_CUDA_CHECK(launch_pdl<4>(umma_conv1d_kernel<-1>, dim3(grid), dim3(kThreads), prm.p.smem_bytes, st, prm));
    }
  }
#undef VS_UMMA_CASE
  VS_LAUNCH_CHECK();
  return VS_OK;
}

}  // namespace vs
