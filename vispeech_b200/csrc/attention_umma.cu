// Windowed relative-position self-attention (reference attentions.py:148-179) on tcgen05 / TMEM, frame level.
//
//   scores[i,j] = (q_i / sqrt(d)) . k_j + (q_i / sqrt(d)) . Ek[j-i+4]   (second term only for |j-i| <= 4)
//   out_i       = sum_j p[i,j] v_j + sum_{|w|<=4} p[i,i+w] Ev[w+4],      p = softmax_j(scores)
//
// Split into the dense part and the 9-wide band:
//   1. attention_umma_kernel: plain flash attention over ALL keys of the utterance without the relative terms, on the 5th-gen
//      tensor cores.  Per CTA one (utterance, head, 128-query tile); per 128-key tile
//          S = Q K^T       18 x tcgen05.mma M128 N128 K16   (fp16 hi/lo operands, three terms: hi.hi + lo.hi + hi.lo)
//          P = exp2(S - m) by eight softmax warps (two threads per query row, 64 keys each; S read from TMEM, P written
//                          hi/lo into shared memory in the UMMA K-major layout)
//          O += P V        24 x tcgen05.mma M128 N96 K16    (V is the MN-major B operand: its planar tile is used as is)
//      with S double-buffered in TMEM so that S of tile j+1 is issued while the softmax of tile j runs.  The running max is
//      only raised (and O rescaled in TMEM) when a tile's max exceeds it by more than 8 (factor 2^8 in the exp2 domain: P
//      stays far inside fp16's range), which for these scores practically never happens after the first tile.
//      Outputs the UNNORMALISED O, the reference max m (log2 domain) and the row sum l.
//   2. rel_band_fixup_kernel (CUDA cores, one CTA per 64 rows and head): recomputes the <= 9 band scores in fp32, swaps their
//      plain weights exp(S - m) for the true ones exp(S + B - m'), adds the relative-value term and normalises:
//          out = [O a + sum_w (e1_w - e0_w) v_{i+w} + e1_w Ev[w]] / [l a + sum_w (e1_w - e0_w)],  a = 2^(m - m').
// Operands are fp16 hi + lo pairs (x = hi + lo to 22 bits): the three-term product is fp32-accurate (the prior sampling
// amplifies attention errors: plain TF32 / one-term fp16 here breaks the 1e-2 latent bar) at the kind::f16
// tensor rate, and every operand tile - Q, K, V planes of the planar [C/8][R][8] layout - arrives by plain TMA bulk copies.
#include "umma_common.cuh"

namespace vs {
namespace {

using namespace umma;

constexpr int D = kHeadDim;                 // 96
constexpr int TQ = 128, TK = 64;            // 64-key tiles: K and V are double-buffered so that no load sits on the critical path
constexpr int kPlanesD = D / 8;             // 12 planes per head and operand half
constexpr int kQkvPlanes = 3 * kHidden / 8; // 72 planes of the planar qkv tensor (q | k | v, 2 heads each)
constexpr uint32_t kPitchQ = TQ * 16;       // bytes between planes of the 128-row Q / P tiles
constexpr uint32_t kPitchK = TK * 16;       // ... of the 64-row K / V tiles
constexpr uint32_t kHalfQ = kPlanesD * kPitchQ;         // 24576: hi (or lo) half of the Q tile
constexpr uint32_t kHalfK = kPlanesD * kPitchK;         // 12288: hi (or lo) half of a K / V tile
constexpr uint32_t kHalfP = (TK / 8) * kPitchQ;         // 16384: hi (or lo) half of the P tile [128 queries][64 keys]
constexpr uint32_t kOffQ = 0, kOffK = kOffQ + 2 * kHalfQ, kOffV = kOffK + 4 * kHalfK, kOffP = kOffV + 4 * kHalfK;
constexpr uint32_t kOffX = kOffP + 4 * kHalfP;          // two P tiles (one per softmax group), then the (m, l) exchange
constexpr uint32_t kOffBar = kOffX + 2 * TQ * 8;
constexpr int kNumBars = 18;
constexpr uint32_t kSmemBytes = kOffBar + 8 * kNumBars + 16;
static_assert(kSmemBytes <= 227 * 1024, "attention_umma: shared memory");
constexpr int kThreads = 64 + 256;          // TMA producer, MMA issuer, two softmax groups of four warps
constexpr uint32_t kTmemCols = 512;         // S0 | S1 (64 columns each) | O0 | O1 (96 each)
constexpr float kRescaleGap = 8.f;          // raise the reference max only when a tile's max exceeds it by this much (log2 units)

struct Params {
  VsRows rows;
  const __half* q_tiles;   // [q tile][head][hi 12 planes x 128 rows x 8 | lo ...]: one TMA bulk copy per tile (qkv_to_tiles_kernel)
  const __half* k_tiles;   // [k tile][head][hi 12 planes x 64 rows x 8 | lo ...]
  const __half* v_tiles;
  float* o_main;           // [R][192] unnormalised O
  float* m_out;            // [R][2]  reference max, log2 domain, of the scaled scores
  float* l_out;            // [R][2]  row sum of exp2(s - m)
  int R;
  long long* dbg;          // -DVS_UMMA_TIMING only: per CTA 16 clock stamps (tools/attention_timing.py)
};
#ifdef VS_UMMA_TIMING
#define VS_STAMP(slot) do { if (prm.dbg && lane == 0) prm.dbg[((size_t)blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * (size_t)blockIdx.z)) * 16 + (slot)] = clock64() - t_begin; } while (0)
#else
#define VS_STAMP(slot) do { } while (0)
#endif

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t* v) {
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
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16_nowait(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st32_nowait(uint32_t taddr, const uint32_t* v) {
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
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Instruction descriptor: kind::f16, fp16 inputs, fp32 accumulate, M = 128, N = n; b_mn = 1: B is MN-major (see make_idesc)
__device__ __forceinline__ uint32_t idesc_f16(int n, int b_mn) {
  return (1u << 4) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
}

__global__ void __launch_bounds__(kThreads, 1) attention_umma_kernel(const __grid_constant__ Params prm) {
  extern __shared__ __align__(128) uint8_t smem[];
  pdl_trigger();
  const int b = blockIdx.z, h = blockIdx.y;
  const int T = prm.rows.utt_len[b], start = prm.rows.utt_start[b];
  const int q0 = blockIdx.x * TQ;
  if (q0 >= T) return;
  const int nk = (T + TK - 1) / TK;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
#ifdef VS_UMMA_TIMING
  const long long t_begin = clock64();
#endif
  const uint32_t sb = smem_u32(smem);
  const uint32_t q_s = sb + kOffQ, k_s = sb + kOffK, v_s = sb + kOffV, p_s = sb + kOffP, bar = sb + kOffBar;
  // barriers (two of each, index = tile parity, unless noted)
  enum { Q_FULL = 0, K_FULL = 1, K_EMPTY = 3, V_FULL = 5, V_EMPTY = 7, S_FULL = 9, S_EMPTY = 11, P_READY = 13, O_DONE = 15 };
  auto B = [&](int i) { return bar + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + kOffBar + 8 * kNumBars);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kNumBars; ++i) mbar_init(B(i), (i >= S_EMPTY && i < O_DONE) ? 4 : 1);   // S_EMPTY[2], P_READY[2]: one arrive per warp
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();                              // the prologue above overlapped the previous kernel's tail
  const uint32_t tm_s = tmem, tm_o = tmem + 2 * TK;            // O of group g at tm_o + 96 g

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer: ONE bulk copy per operand tile
    // (the tiles are laid out in HBM exactly as they sit in shared memory; as 24 x 1 KB plane slabs per tile the kernel was
    // bound by the number of bulk copies in flight: 360 per CTA, 90 us per launch at the C2 size)
    if (lane == 0) {
      const size_t tq = (size_t)(start / TQ + b) + blockIdx.x, tk0 = (size_t)(start / TK + b);
      mbar_arrive_expect_tx(B(Q_FULL), 2u * kHalfQ);
      bulk_g2s(q_s, prm.q_tiles + (tq * kHeads + h) * (size_t)(kHalfQ), 2u * kHalfQ, B(Q_FULL));
      for (int j = 0; j < nk; ++j) {
        const int st = j & 1, n = j >> 1;
        const size_t off = ((tk0 + j) * kHeads + h) * (size_t)(kHalfK);           // in halves: a tile is 2 * kHalfK bytes
        if (n > 0) mbar_wait(B(K_EMPTY + st), (uint32_t)(n - 1) & 1u, 60);
        mbar_arrive_expect_tx(B(K_FULL + st), 2u * kHalfK);
        bulk_g2s(k_s + (uint32_t)st * 2u * kHalfK, prm.k_tiles + off, 2u * kHalfK, B(K_FULL + st));
        if (n > 0) mbar_wait(B(V_EMPTY + st), (uint32_t)(n - 1) & 1u, 61);
        mbar_arrive_expect_tx(B(V_FULL + st), 2u * kHalfK);
        bulk_g2s(v_s + (uint32_t)st * 2u * kHalfK, prm.v_tiles + off, 2u * kHalfK, B(V_FULL + st));
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (warp-uniform, elected lane issues)
    const uint32_t id_s = idesc_f16(TK, 0), id_o = idesc_f16(D, 1);
    const uint32_t q_hi = (uint32_t)(make_desc(0, kPitchQ, 128u) >> 32), q_lo = (uint32_t)make_desc(0, kPitchQ, 128u);   // K-major, 128-row tiles
    const uint32_t k_hi = (uint32_t)(make_desc(0, kPitchK, 128u) >> 32), k_lo = (uint32_t)make_desc(0, kPitchK, 128u);   // K-major, 64-row tiles
    const uint32_t v_hi = (uint32_t)(make_desc(0, 128u, kPitchK) >> 32), v_lo = (uint32_t)make_desc(0, 128u, kPitchK);   // MN-major V: LBO = 8 keys, SBO = plane pitch
    auto issue_s = [&](int j) {                                                    // S[j & 1] = Q K_j^T, three terms
      const uint32_t d = tm_s + (uint32_t)(j & 1) * TK;
      const uint32_t kb = k_s + (uint32_t)(j & 1) * 2u * kHalfK;
      uint32_t acc = 0;
#pragma unroll
      for (int s = 0; s < D / 16; ++s) {
        const uint32_t a_h = q_lo + ((q_s + (uint32_t)(2 * s) * kPitchQ) >> 4), a_l = q_lo + ((q_s + kHalfQ + (uint32_t)(2 * s) * kPitchQ) >> 4);
        const uint32_t b_h = k_lo + ((kb + (uint32_t)(2 * s) * kPitchK) >> 4), b_l = k_lo + ((kb + kHalfK + (uint32_t)(2 * s) * kPitchK) >> 4);
        tc_mma_f16_lohi(d, a_h, q_hi, b_h, k_hi, id_s, acc);
        tc_mma_f16_lohi(d, a_l, q_hi, b_h, k_hi, id_s, 1u);
        tc_mma_f16_lohi(d, a_h, q_hi, b_l, k_hi, id_s, 1u);
        acc = 1u;
      }
    };
    VS_STAMP(0);                                   // set-up done (barriers, TMEM)
    mbar_wait(B(Q_FULL), 0, 62);
    mbar_wait(B(K_FULL), 0, 63);
    VS_STAMP(1);                                   // Q and K_0 have landed
    tc_fence_after();
    issue_s(0);
    tc_commit(B(S_FULL));
    tc_commit(B(K_EMPTY));
    for (int j = 0; j < nk; ++j) {
      if (j + 1 < nk) {
        const int jn = j + 1, st = jn & 1, n = jn >> 1;
        mbar_wait(B(K_FULL + st), (uint32_t)n & 1u, 64);
        if (n > 0) mbar_wait(B(S_EMPTY + st), (uint32_t)(n - 1) & 1u, 65);
        tc_fence_after();
        issue_s(jn);
        tc_commit(B(S_FULL + st));
        tc_commit(B(K_EMPTY + st));
      }
      const int st = j & 1;                                                        // = the softmax group of this tile
      mbar_wait(B(P_READY + st), (uint32_t)(j >> 1) & 1u, 66);
      mbar_wait(B(V_FULL + st), (uint32_t)(j >> 1) & 1u, 67);
      tc_fence_after();
      const uint32_t vb = v_s + (uint32_t)st * 2u * kHalfK, pb = p_s + (uint32_t)st * 2u * kHalfP, od = tm_o + (uint32_t)st * D;
      uint32_t acc = j > 1 ? 1u : 0u;
#pragma unroll
      for (int s = 0; s < TK / 16; ++s) {                                          // O[group] += P V_j, three terms
        const uint32_t a_h = q_lo + ((pb + (uint32_t)(2 * s) * kPitchQ) >> 4), a_l = q_lo + ((pb + kHalfP + (uint32_t)(2 * s) * kPitchQ) >> 4);
        const uint32_t b_h = v_lo + ((vb + (uint32_t)s * 256u) >> 4), b_l = v_lo + ((vb + kHalfK + (uint32_t)s * 256u) >> 4);
        tc_mma_f16_lohi(od, a_h, q_hi, b_h, v_hi, id_o, acc);
        tc_mma_f16_lohi(od, a_l, q_hi, b_h, v_hi, id_o, 1u);
        tc_mma_f16_lohi(od, a_h, q_hi, b_l, v_hi, id_o, 1u);
        acc = 1u;
      }
      tc_commit(B(O_DONE + st));
      tc_commit(B(V_EMPTY + st));
      if (j < 6) VS_STAMP(2 + j);                  // PV_j issued
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ softmax warps (thread = query row = TMEM lane)
    // Two groups of four warps; group g owns the key tiles j = g (mod 2) with its own S buffer, P tile, running (m, l) and O
    // accumulator - a split-key flash attention inside the CTA - so the two groups' latency chains (barrier -> tcgen05.ld ->
    // exp2 -> smem store -> barrier) overlap.  The halves are merged at the end: O = O0 2^(m0 - m) + O1 2^(m1 - m).
    const int q = warp & 3, grp = (warp - 2) >> 2;
    const int lrow = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const float c2 = rsqrtf((float)D) * 1.4426950408889634f;                       // scores -> log2 domain
    const uint32_t to = tm_o + lane_off + (uint32_t)grp * D;
    const uint32_t prow = p_s + (uint32_t)grp * 2u * kHalfP + (uint32_t)lrow * 16u;
    float m_ref = -INFINITY, l = 0.f;
    int n = 0;                                                                     // tiles this group has processed
    for (int j = grp; j < nk; j += 2, ++n) {
      mbar_wait(B(S_FULL + grp), (uint32_t)n & 1u, 68);
      tc_fence_after();
      uint32_t v[TK];
      const uint32_t ts = tm_s + lane_off + (uint32_t)grp * TK;
#pragma unroll
      for (int cch = 0; cch < TK / 32; ++cch) tmem_ld32_nowait(ts + 32u * cch, v + 32 * cch);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(B(S_EMPTY + grp));                                // S is in registers: the buffer may be overwritten
      const int n_valid = T - j * TK;                                              // keys of this tile inside the utterance
      float mx = -INFINITY;
#pragma unroll
      for (int e = 0; e < TK; ++e) {
        const float s = e < n_valid ? __uint_as_float(v[e]) * c2 : -INFINITY;
        v[e] = __float_as_uint(s);
        mx = fmaxf(mx, s);
      }
      float factor = 1.f;
      if (mx > m_ref + kRescaleGap) {                                              // first tile: m_ref = -inf -> factor 0 (O, l are 0)
        factor = ex2(m_ref - mx);
        m_ref = mx;
        l *= factor;
      }
      // exp2 / hi-lo split while the group's previous PV may still be running; only the store has to wait for it
      uint32_t hi[TK / 2], lo[TK / 2];
      float l0 = 0.f, l1 = 0.f;
#pragma unroll
      for (int e = 0; e < TK / 2; ++e) {
        const float p0 = ex2(__uint_as_float(v[2 * e]) - m_ref), p1 = ex2(__uint_as_float(v[2 * e + 1]) - m_ref);
        l0 += p0;
        l1 += p1;
        hi[e] = pack_f16x2(p0, p1);
        const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi[e]));
        lo[e] = pack_f16x2(p0 - hf.x, p1 - hf.y);
      }
      l += l0 + l1;
      if (n > 0) {
        mbar_wait(B(O_DONE + grp), (uint32_t)(n - 1) & 1u, 69);                    // the group's previous PV is done: P may be overwritten, O is current
        tc_fence_after();
        if (__any_sync(0xffffffffu, factor != 1.f)) {                              // rare: rescale this warp's 32 rows of O in TMEM
          uint32_t o[D];
#pragma unroll
          for (int cch = 0; cch < D / 32; ++cch) tmem_ld32_nowait(to + 32u * cch, o + 32 * cch);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < D; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * factor);
#pragma unroll
          for (int cch = 0; cch < D / 32; ++cch) tmem_st32_nowait(to + 32u * cch, o + 32 * cch);
          tmem_st_wait();
        }
      }
#pragma unroll
      for (int g = 0; g < TK / 8; ++g) {
        sts128(prow + (uint32_t)g * kPitchQ, hi[4 * g], hi[4 * g + 1], hi[4 * g + 2], hi[4 * g + 3]);
        sts128(prow + kHalfP + (uint32_t)g * kPitchQ, lo[4 * g], lo[4 * g + 1], lo[4 * g + 2], lo[4 * g + 3]);
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(B(P_READY + grp));
      if (warp == 2 && n < 3) VS_STAMP(8 + n);     // group 0, lane quarter 2: P of its n-th tile stored
    }
    // merge the two groups: each publishes (m, l), then takes half of the 96 output columns of both accumulators
    float2* const xch = reinterpret_cast<float2*>(smem + kOffX);                   // [group][row]
    xch[grp * TQ + lrow] = make_float2(m_ref, l);
    if (n > 0) {
      mbar_wait(B(O_DONE + grp), (uint32_t)(n - 1) & 1u, 70);                      // this group's last PV
      tc_fence_after();
    }
    if (warp == 2) VS_STAMP(11);                   // group 0 saw its last O_DONE
    tc_fence_before();
    asm volatile("bar.sync 1, 256;" ::: "memory");                                 // both groups' accumulators are final
    tc_fence_after();
    const float2 ml0 = xch[lrow], ml1 = xch[TQ + lrow];
    const bool has1 = nk > 1;                                                      // group 1 saw no tile: its O was never written
    const float m = has1 ? fmaxf(ml0.x, ml1.x) : ml0.x;
    const float a0 = ex2(ml0.x - m), a1 = has1 ? ex2(ml1.x - m) : 0.f;
    uint32_t o0[48], o1[48];
    const uint32_t tc0 = tm_o + lane_off + (uint32_t)grp * 48u;                    // this thread's 48 columns of O0; O1 is D columns on
    tmem_ld32_nowait(tc0, o0);
    tmem_ld16_nowait(tc0 + 32u, o0 + 32);
    if (has1) {
      tmem_ld32_nowait(tc0 + D, o1);
      tmem_ld16_nowait(tc0 + D + 32u, o1 + 32);
    }
    tmem_ld_wait();
    if (q0 + lrow < T) {
      const size_t r = (size_t)start + q0 + lrow;
      float4* dst = reinterpret_cast<float4*>(prm.o_main + r * kHidden + h * D + grp * 48);
#pragma unroll
      for (int e = 0; e < 12; ++e) {
        float f[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          f[c] = __uint_as_float(o0[4 * e + c]) * a0;
          if (has1) f[c] = fmaf(__uint_as_float(o1[4 * e + c]), a1, f[c]);
        }
        dst[e] = make_float4(f[0], f[1], f[2], f[3]);
      }
      if (grp == 0) {
        prm.m_out[r * kHeads + h] = m;
        prm.l_out[r * kHeads + h] = ml0.y * a0 + ml1.y * a1;
      }
    }
    if (warp == 2) VS_STAMP(12);                   // output stored
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
  }
}

// fp32 row-major qkv [R][576] -> fp16 hi / lo operand TILES (x = hi + lo to 22 bits), laid out as they sit in shared memory:
//   K, V: tile (b, j) = rows [64 j, 64 j + 64) of utterance b at index start_b / 64 + b + j (a bound on the tiles of the
//         utterances before it), per head [hi: 12 planes x 64 rows x 8 ch][lo: same];  Q: 128-row tiles at start_b / 128 + b + j.
// One CTA per (64-row tile, utterance) writes every row of its tile (zeros past the utterance's end), so the attention
// kernel never reads a byte nobody wrote and needs no bounds handling.
__global__ void __launch_bounds__(256) qkv_to_tiles_kernel(VsRows rows, const float* __restrict__ x, __half* __restrict__ qt,
                                                           __half* __restrict__ kt, __half* __restrict__ vt) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y, j = blockIdx.x;
  const int T = rows.utt_len[b], start = rows.utt_start[b];
  if (j * TK >= T) return;
  const size_t tk = (size_t)(start / TK + b) + j, tq = (size_t)(start / TQ + b) + (j >> 1);
  for (int idx = threadIdx.x; idx < TK * kQkvPlanes; idx += 256) {
    const int row = idx % TK, p = idx / TK;                 // consecutive threads: consecutive rows of one plane (16 B apart)
    const int r = j * TK + row;
    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (r < T) {
      const float4 a = *reinterpret_cast<const float4*>(x + (size_t)(start + r) * 3 * kHidden + p * 8);
      const float4 c = *reinterpret_cast<const float4*>(x + (size_t)(start + r) * 3 * kHidden + p * 8 + 4);
      f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = c.x; f[5] = c.y; f[6] = c.z; f[7] = c.w;
    }
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      hw[e] = pack_f16x2(f[2 * e], f[2 * e + 1]);
      const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hw[e]));
      lw[e] = pack_f16x2(f[2 * e] - hf.x, f[2 * e + 1] - hf.y);
    }
    const int op = p / (2 * kPlanesD), hh = (p / kPlanesD) & 1, pl = p % kPlanesD;
    uint8_t* dst;
    uint32_t half;
    if (op == 0) {
      dst = reinterpret_cast<uint8_t*>(qt) + (tq * kHeads + hh) * (size_t)(2 * kHalfQ) + (size_t)pl * kPitchQ + (size_t)((j & 1) * TK + row) * 16;
      half = kHalfQ;
    } else {
      dst = reinterpret_cast<uint8_t*>(op == 1 ? kt : vt) + (tk * kHeads + hh) * (size_t)(2 * kHalfK) + (size_t)pl * kPitchK + (size_t)row * 16;
      half = kHalfK;
    }
    *reinterpret_cast<uint4*>(dst) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    *reinterpret_cast<uint4*>(dst + half) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
  }
}

// The 9-wide band: relative-key scores, relative-value term, normalisation.  One CTA per (64 consecutive rows, head): the
// q / k / v rows it touches (64 + 8 halo rows) are staged in shared memory once, then
//   phase 2: the 9 plain and 9 relative-key scores of every row,
//   phase 3: per row the new max, the rescale factor and the weights (e1 - e0)_w, e1_w,
//   phase 4: out = [O a + sum_w (e1 - e0)_w v_{i+w} + e1_w Ev[w]] / den.
// Phases 2 and 4 use register tiles of 4 rows x 9 offsets per thread over a 6-channel slice (a thread per (row, w) re-read
// both operands of every FMA from shared memory: 1.3 MB of smem reads per CTA, 80 us per launch; the tiles cut that 9x).
__device__ __forceinline__ void fx_cp_async16(void* dst, const void* src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  const int n = valid ? 16 : 0;                      // src-size 0 -> the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}
constexpr int FX_ROWS = 64, FX_HALO = kWindow, FX_PITCH = 100;   // pitch 100 floats: rows 8 apart share banks, 4 apart do not
struct FxSmem {
  float q[FX_ROWS][FX_PITCH];
  float k[FX_ROWS + 2 * FX_HALO][FX_PITCH];
  float v[FX_ROWS + 2 * FX_HALO][FX_PITCH];
  float ek[kRel][D], ev[kRel][D];
  float s2[FX_ROWS][kRel], b2[FX_ROWS][kRel], c1[FX_ROWS][kRel], c2[FX_ROWS][kRel];
  float a[FX_ROWS], inv[FX_ROWS];
};

__global__ void __launch_bounds__(256, 2) rel_band_fixup_kernel(VsRows rows, const float* __restrict__ qkv, const float* __restrict__ ek,
                                                             const float* __restrict__ ev, const float* __restrict__ o_main,
                                                             const float* __restrict__ m_in, const float* __restrict__ l_in,
                                                             float* __restrict__ out, __half* __restrict__ out_hi,
                                                             __half* __restrict__ out_lo, int R) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) unsigned char fx_raw[];
  FxSmem& S = *reinterpret_cast<FxSmem*>(fx_raw);
  const int tid = threadIdx.x, h = blockIdx.y, r0 = blockIdx.x * FX_ROWS;
  const int ld = 3 * kHidden;
  const float scale = rsqrtf((float)D), log2e = 1.4426950408889634f;
  // all tile loads in flight at once (cp.async, zero-filled outside [0, R)): as load -> store pairs in a loop every one of
  // the ~20 iterations paid a full L2 round trip (ncu: 80 % of the stall samples on the first STS / FMUL)
  for (int i = tid; i < (FX_ROWS + 2 * FX_HALO) * (D / 4); i += 256) {
    const int rr = i / (D / 4), c4 = i % (D / 4);
    const int g = r0 - FX_HALO + rr;
    const bool ok = g >= 0 && g < R;
    const float* src = qkv + (size_t)(ok ? g : 0) * ld + kHidden + h * D + 4 * c4;
    fx_cp_async16(&S.k[rr][4 * c4], src, ok);
    fx_cp_async16(&S.v[rr][4 * c4], src + kHidden, ok);
  }
  for (int i = tid; i < FX_ROWS * (D / 4); i += 256) {
    const int rr = i / (D / 4), c4 = i % (D / 4);
    const int g = r0 + rr;
    const bool ok = g < R;
    fx_cp_async16(&S.q[rr][4 * c4], qkv + (size_t)(ok ? g : 0) * ld + h * D + 4 * c4, ok);   // unscaled: 1/sqrt(d) goes on the scores
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int i = tid; i < kRel * D; i += 256) { S.ek[i / D][i % D] = ek[i]; S.ev[i / D][i % D] = ev[i]; }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  // register tiles: thread = (group of 4 consecutive rows, 6-channel slice); the 16 threads of a row group are a half warp
  const int rg = tid >> 4, cs = tid & 15;
  const int R0 = 4 * rg, C0 = 6 * cs;
  // phase 2: band scores (log2 domain): acc[row][w] over this thread's 6 channels, then a half-warp butterfly
  {
    float as[4][kRel], ab[4][kRel];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int w = 0; w < kRel; ++w) { as[r][w] = 0.f; ab[r][w] = 0.f; }
#pragma unroll
    for (int st = 0; st < 3; ++st) {
      const int c = C0 + 2 * st;
      float2 qq[4], kk[4 + kRel - 1], ee[kRel];
#pragma unroll
      for (int r = 0; r < 4; ++r) qq[r] = *reinterpret_cast<const float2*>(&S.q[R0 + r][c]);
#pragma unroll
      for (int r = 0; r < 4 + kRel - 1; ++r) kk[r] = *reinterpret_cast<const float2*>(&S.k[R0 + r][c]);
#pragma unroll
      for (int w = 0; w < kRel; ++w) ee[w] = *reinterpret_cast<const float2*>(&S.ek[w][c]);
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int w = 0; w < kRel; ++w) {
          as[r][w] = fmaf(qq[r].x, kk[r + w].x, fmaf(qq[r].y, kk[r + w].y, as[r][w]));
          ab[r][w] = fmaf(qq[r].x, ee[w].x, fmaf(qq[r].y, ee[w].y, ab[r][w]));
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int w = 0; w < kRel; ++w) {
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
          as[r][w] += __shfl_xor_sync(0xffffffffu, as[r][w], o);
          ab[r][w] += __shfl_xor_sync(0xffffffffu, ab[r][w], o);
        }
        if (cs == w) {                                           // lane w of the half warp stores score w
          S.s2[R0 + r][w] = as[r][w] * (scale * log2e);
          S.b2[R0 + r][w] = ab[r][w] * (scale * log2e);
        }
      }
  }
  __syncthreads();
  // phase 3: weights
  if (tid < FX_ROWS) {
    const int rw = tid, gg = r0 + rw;
    const int uu = gg < R ? rows.row_utt[gg] : -1;
    float a = 0.f, inv = 0.f;
#pragma unroll
    for (int w = 0; w < kRel; ++w) { S.c1[rw][w] = 0.f; S.c2[rw][w] = 0.f; }
    if (uu >= 0) {
      const int ii = gg - rows.utt_start[uu], TT = rows.utt_len[uu];
      const float m = m_in[(size_t)gg * kHeads + h], l = l_in[(size_t)gg * kHeads + h];
      float m2 = m;
#pragma unroll
      for (int w = 0; w < kRel; ++w) {
        const int j = ii + w - kWindow;
        if (j >= 0 && j < TT) m2 = fmaxf(m2, S.s2[rw][w] + S.b2[rw][w]);
      }
      a = exp2f(m - m2);
      float den = l * a;
#pragma unroll
      for (int w = 0; w < kRel; ++w) {
        const int j = ii + w - kWindow;
        if (j >= 0 && j < TT) {
          const float e1 = exp2f(S.s2[rw][w] + S.b2[rw][w] - m2);
          const float d1 = e1 - exp2f(S.s2[rw][w] - m2);
          S.c1[rw][w] = d1;
          S.c2[rw][w] = e1;
          den += d1;
        }
      }
      inv = 1.f / den;
    }
    S.a[rw] = a;
    S.inv[rw] = inv;                                             // 0 on gap rows: their output is zero
  }
  __syncthreads();
  // phase 4: out = (O a + sum_w c1_w v_{i+w} + c2_w Ev[w]) / den for 4 rows x 6 channels
  {
    float2 acc[4][3];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int g = r0 + R0 + r;
      const float a = S.a[R0 + r];
#pragma unroll
      for (int st = 0; st < 3; ++st) {
        float2 o2 = make_float2(0.f, 0.f);
        if (g < R && a != 0.f) o2 = *reinterpret_cast<const float2*>(o_main + (size_t)g * kHidden + h * D + C0 + 2 * st);
        acc[r][st] = make_float2(o2.x * a, o2.y * a);
      }
    }
#pragma unroll
    for (int st = 0; st < 3; ++st) {
      const int c = C0 + 2 * st;
      float2 vv[4 + kRel - 1], ee[kRel];
#pragma unroll
      for (int r = 0; r < 4 + kRel - 1; ++r) vv[r] = *reinterpret_cast<const float2*>(&S.v[R0 + r][c]);
#pragma unroll
      for (int w = 0; w < kRel; ++w) ee[w] = *reinterpret_cast<const float2*>(&S.ev[w][c]);
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int w = 0; w < kRel; ++w) {
          const float d1 = S.c1[R0 + r][w], e1 = S.c2[R0 + r][w];
          acc[r][st].x = fmaf(d1, vv[r + w].x, fmaf(e1, ee[w].x, acc[r][st].x));
          acc[r][st].y = fmaf(d1, vv[r + w].y, fmaf(e1, ee[w].y, acc[r][st].y));
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int g = r0 + R0 + r;
      if (g >= R) continue;
      const float inv = S.inv[R0 + r];
#pragma unroll
      for (int st = 0; st < 3; ++st) {
        const float2 y = make_float2(acc[r][st].x * inv, acc[r][st].y * inv);
        const int ch = h * D + C0 + 2 * st;                                     // even: the pair sits in one 8-channel plane
        if (out_hi) {                                                            // the O conv's operand: planar fp16 hi / lo
          const size_t o = ((size_t)(ch >> 3) * R + g) * 8 + (ch & 7);
          const __half2 hh = __floats2half2_rn(y.x, y.y);
          const float2 hf = __half22float2(hh);
          *reinterpret_cast<__half2*>(out_hi + o) = hh;
          *reinterpret_cast<__half2*>(out_lo + o) = __floats2half2_rn(y.x - hf.x, y.y - hf.y);
        } else {
          *reinterpret_cast<float2*>(out + (size_t)g * kHidden + ch) = y;
        }
      }
    }
  }
}

}  // namespace

// tiles: K / V <= R / 64 + n_utt + 1, Q <= R / 128 + n_utt + 1 (every utterance rounds its last tile up)
static int64_t umma_ws_bytes(int R, int n_utt) {
  const int64_t tk = R / TK + n_utt + 1, tq = R / TQ + n_utt + 1;
  return 2 * tk * kHeads * 2 * kHalfK + tq * kHeads * 2 * kHalfQ + (int64_t)R * (kHidden + 2 * kHeads) * 4 + 5 * 256;
}
// sized for batches whose utterances average >= 64 rows; a batch of many tiny utterances next to a long one needs more and
// takes the mma.sync kernel instead (rel_attention_umma_fits)
int64_t attention_umma_ws_floats(int R) { return (umma_ws_bytes(R, R / 64 + 1) + 3) / 4; }
bool rel_attention_umma_fits(const VsRows& rows, const Workspace& ws) {
  return ws.size - ws.off >= umma_ws_bytes(rows.n_rows, rows.n_utt);
}

int rel_attention_umma(const VsRows& rows, const float* qkv, const float* ek, const float* ev, float* out, Workspace& ws,
                       cudaStream_t st, __half* out_hi, __half* out_lo) {
  const int R = rows.n_rows;
  VS_REQUIRE(rows.max_len > 0 && rows.max_len <= R, "rel_attention_umma: bad max_len %d", rows.max_len);
  const int64_t tk = R / TK + rows.n_utt + 1, tq = R / TQ + rows.n_utt + 1;
  __half* kt = ws.take<__half>(tk * kHeads * kHalfK);
  __half* vt = ws.take<__half>(tk * kHeads * kHalfK);
  __half* qt = ws.take<__half>(tq * kHeads * kHalfQ);
  float* o_main = ws.take<float>((int64_t)R * kHidden);
  float* m = ws.take<float>((int64_t)R * kHeads);
  float* l = ws.take<float>((int64_t)R * kHeads);
  if (!ws.ok) { set_error("rel_attention_umma: workspace too small"); return VS_ERR_WORKSPACE; }
  VS_CUDA_CHECK(launch_pdl<64>(qkv_to_tiles_kernel, dim3(dim3((rows.max_len + TK - 1) / TK, rows.n_utt)), dim3(256), 0, st, rows, qkv, qt, kt, vt));
  VS_LAUNCH_CHECK();
  VS_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(attention_umma_kernel), (int)kSmemBytes));
  Params prm;
  prm.rows = rows; prm.q_tiles = qt; prm.k_tiles = kt; prm.v_tiles = vt; prm.o_main = o_main; prm.m_out = m; prm.l_out = l; prm.R = R;
  prm.dbg = reinterpret_cast<long long*>(opts().v[OPT_TIMING_BUFFER]);
  dim3 grid((rows.max_len + TQ - 1) / TQ, kHeads, rows.n_utt);
  VS_CUDA_CHECK(launch_pdl<2>(attention_umma_kernel, dim3(grid), dim3(kThreads), kSmemBytes, st, prm));
  VS_LAUNCH_CHECK();
  VS_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(rel_band_fixup_kernel), (int)sizeof(FxSmem)));
  VS_CUDA_CHECK(launch_pdl<128>(rel_band_fixup_kernel, dim3(dim3((R + FX_ROWS - 1) / FX_ROWS, kHeads)), dim3(256), sizeof(FxSmem), st, rows, qkv, ek, ev, o_main, m, l, out, out_hi, out_lo, R));
  VS_LAUNCH_CHECK();
  return VS_OK;
}

}  // namespace vs
