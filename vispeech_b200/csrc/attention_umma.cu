// Windowed relative-position self-attention (reference attentions.py:148-179) on tcgen05 / TMEM, frame level.
//
//   scores[i,j] = (q_i / sqrt(d)) . k_j + (q_i / sqrt(d)) . Ek[j-i+4]   (second term only for |j-i| <= 4)
//   out_i       = sum_j p[i,j] v_j + sum_{|w|<=4} p[i,i+w] Ev[w+4],      p = softmax_j(scores)
//
// Split into the dense part and the 9-wide band:
//   1. attention_umma_kernel: plain flash attention over ALL keys of the utterance without the relative terms, on the 5th-gen
//      tensor cores.  Per CTA one (utterance, head, 128-query tile); per 128-key tile
//          S = Q K^T       18 x tcgen05.mma M128 N128 K16   (fp16 hi/lo operands, three terms: hi.hi + lo.hi + hi.lo)
//          P = exp2(S - m) by four softmax warps (thread = query row; S read from TMEM, P written hi/lo into shared memory
//                          in the UMMA K-major layout)
//          O += P V        24 x tcgen05.mma M128 N96 K16    (V is the MN-major B operand: its planar tile is used as is)
//      with S double-buffered in TMEM so that S of tile j+1 is issued while the softmax of tile j runs.  The running max is
//      only raised (and O rescaled in TMEM) when a tile's max exceeds it by more than 8 (factor 2^8 in the exp2 domain: P
//      stays far inside fp16's range), which for these scores practically never happens after the first tile.
//      Outputs the UNNORMALISED O, the reference max m (log2 domain) and the row sum l.
//   2. rel_band_fixup_kernel (CUDA cores, one warp per (row, head)): recomputes the <= 9 band scores in fp32, swaps their
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
constexpr int TQ = 128, TK = 128;
constexpr int kPlanesD = D / 8;             // 12 planes per head and operand half
constexpr int kQkvPlanes = 3 * kHidden / 8; // 72 planes of the planar qkv tensor (q | k | v, 2 heads each)
constexpr uint32_t kPitch = TQ * 16;        // bytes between planes of a 128-row tile
constexpr uint32_t kHalfQ = kPlanesD * kPitch;          // 24576: hi (or lo) half of a Q / K / V tile
constexpr uint32_t kHalfP = (TK / 8) * kPitch;          // 32768: hi (or lo) half of the P tile
constexpr uint32_t kOffQ = 0, kOffK = kOffQ + 2 * kHalfQ, kOffV = kOffK + 2 * kHalfQ, kOffP = kOffV + 2 * kHalfQ;
constexpr uint32_t kOffBar = kOffP + 2 * kHalfP;
constexpr int kNumBars = 12;
constexpr uint32_t kSmemBytes = kOffBar + 8 * kNumBars + 16;
static_assert(kSmemBytes <= 227 * 1024, "attention_umma: shared memory");
constexpr int kThreads = 64 + 128;          // TMA producer, MMA issuer, four softmax warps
constexpr float kRescaleGap = 8.f;          // raise the reference max only when a tile's max exceeds it by this much (log2 units)

struct Params {
  VsRows rows;
  const __half* hi;        // planar [72][R][8]: q (2 x 12 planes) | k | v
  const __half* lo;
  float* o_main;           // [R][192] unnormalised O
  float* m_out;            // [R][2]  reference max, log2 domain, of the scaled scores
  float* l_out;            // [R][2]  row sum of exp2(s - m)
  int R;
};

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
  const int b = blockIdx.z, h = blockIdx.y;
  const int T = prm.rows.utt_len[b], start = prm.rows.utt_start[b];
  const int q0 = blockIdx.x * TQ;
  if (q0 >= T) return;
  const int nk = (T + TK - 1) / TK;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t sb = smem_u32(smem);
  const uint32_t q_s = sb + kOffQ, k_s = sb + kOffK, v_s = sb + kOffV, p_s = sb + kOffP, bar = sb + kOffBar;
  // barriers: 0 q_full | 1 k_full | 2 k_empty | 3 v_full | 4 v_empty | 5,6 s_full[2] | 7,8 s_empty[2] | 9 p_ready | 10 o_done
  auto B = [&](int i) { return bar + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + kOffBar + 8 * kNumBars);

  if (threadIdx.x == 0) {
    for (int i = 0; i < 5; ++i) mbar_init(B(i), 1);
    mbar_init(B(5), 1); mbar_init(B(6), 1);
    mbar_init(B(7), 4); mbar_init(B(8), 4);
    mbar_init(B(9), 4);
    mbar_init(B(10), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // Q, K and V tiles start as zeros: rows past the end of the buffer are never loaded and whatever sits there meets P = 0 or a
  // masked score, which must stay finite
  for (uint32_t i = threadIdx.x; i < kOffP / 16; i += kThreads) sts128(sb + i * 16u, 0u, 0u, 0u, 0u);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tm_s = tmem, tm_o = tmem + 256;
  const int R = prm.R;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      auto load_tile = [&](uint32_t dst, int plane0, int row0, uint32_t full_bar) {   // 12 hi + 12 lo plane slabs of 128 rows
        int n = R - row0;
        n = n > TQ ? TQ : n;
        const uint32_t bytes = (uint32_t)n * 16u;
        mbar_arrive_expect_tx(full_bar, bytes * 2u * kPlanesD);
        for (int pl = 0; pl < kPlanesD; ++pl) {
          const size_t off = ((size_t)(plane0 + pl) * R + row0) * 8;
          bulk_g2s(dst + (uint32_t)pl * kPitch, prm.hi + off, bytes, full_bar);
          bulk_g2s(dst + kHalfQ + (uint32_t)pl * kPitch, prm.lo + off, bytes, full_bar);
        }
      };
      load_tile(q_s, h * kPlanesD, start + q0, B(0));
      for (int j = 0; j < nk; ++j) {
        if (j > 0) mbar_wait(B(2), (uint32_t)(j - 1) & 1u, 60);
        load_tile(k_s, 24 + h * kPlanesD, start + j * TK, B(1));
        if (j > 0) mbar_wait(B(4), (uint32_t)(j - 1) & 1u, 61);
        load_tile(v_s, 48 + h * kPlanesD, start + j * TK, B(3));
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (warp-uniform, elected lane issues)
    const uint32_t id_s = idesc_f16(TK, 0), id_o = idesc_f16(D, 1);
    const uint32_t kk_hi = (uint32_t)(make_desc(0, kPitch, 128u) >> 32);          // K-major: LBO = plane pitch, SBO = 128 B
    const uint32_t kk_lo = (uint32_t)make_desc(0, kPitch, 128u);
    const uint32_t mn_hi = (uint32_t)(make_desc(0, 128u, kPitch) >> 32);          // MN-major V: LBO = 128 B (8 keys), SBO = plane pitch
    const uint32_t mn_lo = (uint32_t)make_desc(0, 128u, kPitch);
    auto issue_s = [&](int j) {                                                    // S[j & 1] = Q K_j^T, three terms
      const uint32_t d = tm_s + (uint32_t)(j & 1) * TK;
      uint32_t acc = 0;
#pragma unroll
      for (int s = 0; s < D / 16; ++s) {
        const uint32_t step = (uint32_t)(2 * s) * kPitch;
        const uint32_t a_h = kk_lo + ((q_s + step) >> 4), a_l = kk_lo + ((q_s + kHalfQ + step) >> 4);
        const uint32_t b_h = kk_lo + ((k_s + step) >> 4), b_l = kk_lo + ((k_s + kHalfQ + step) >> 4);
        tc_mma_f16_lohi(d, a_h, kk_hi, b_h, kk_hi, id_s, acc);
        tc_mma_f16_lohi(d, a_l, kk_hi, b_h, kk_hi, id_s, 1u);
        tc_mma_f16_lohi(d, a_h, kk_hi, b_l, kk_hi, id_s, 1u);
        acc = 1u;
      }
    };
    mbar_wait(B(0), 0, 62);
    mbar_wait(B(1), 0, 63);
    tc_fence_after();
    issue_s(0);
    tc_commit(B(5));
    tc_commit(B(2));
    for (int j = 0; j < nk; ++j) {
      if (j + 1 < nk) {
        const int jn = j + 1;
        mbar_wait(B(1), (uint32_t)jn & 1u, 64);
        if (jn >= 2) mbar_wait(B(7 + (jn & 1)), (uint32_t)((jn >> 1) - 1) & 1u, 65);
        tc_fence_after();
        issue_s(jn);
        tc_commit(B(5 + (jn & 1)));
        tc_commit(B(2));
      }
      mbar_wait(B(9), (uint32_t)j & 1u, 66);
      mbar_wait(B(3), (uint32_t)j & 1u, 67);
      tc_fence_after();
      uint32_t acc = j > 0 ? 1u : 0u;
#pragma unroll
      for (int s = 0; s < TK / 16; ++s) {                                          // O += P V_j, three terms
        const uint32_t a_h = kk_lo + ((p_s + (uint32_t)(2 * s) * kPitch) >> 4), a_l = kk_lo + ((p_s + kHalfP + (uint32_t)(2 * s) * kPitch) >> 4);
        const uint32_t b_h = mn_lo + ((v_s + (uint32_t)s * 256u) >> 4), b_l = mn_lo + ((v_s + kHalfQ + (uint32_t)s * 256u) >> 4);
        tc_mma_f16_lohi(tm_o, a_h, kk_hi, b_h, mn_hi, id_o, acc);
        tc_mma_f16_lohi(tm_o, a_l, kk_hi, b_h, mn_hi, id_o, 1u);
        tc_mma_f16_lohi(tm_o, a_h, kk_hi, b_l, mn_hi, id_o, 1u);
        acc = 1u;
      }
      tc_commit(B(10));
      tc_commit(B(4));
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ softmax warps (thread = query row = TMEM lane)
    const int q = warp & 3;
    const int lrow = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const float c2 = rsqrtf((float)D) * 1.4426950408889634f;                       // scores -> log2 domain
    float m_ref = -INFINITY, l = 0.f;
    for (int j = 0; j < nk; ++j) {
      mbar_wait(B(5 + (j & 1)), (uint32_t)(j >> 1) & 1u, 68);
      tc_fence_after();
      uint32_t v[TK];
      const uint32_t ts = tm_s + lane_off + (uint32_t)(j & 1) * TK;
#pragma unroll
      for (int cch = 0; cch < TK / 32; ++cch) tmem_ld32_nowait(ts + 32u * cch, v + 32 * cch);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(B(7 + (j & 1)));                                  // S is in registers: the buffer may be overwritten
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
      if (j > 0) {
        mbar_wait(B(10), (uint32_t)(j - 1) & 1u, 69);                              // PV_{j-1} done: P may be overwritten, O is current
        tc_fence_after();
        if (__any_sync(0xffffffffu, factor != 1.f)) {                              // rare: rescale this warp's 32 rows of O in TMEM
          uint32_t o[D];
#pragma unroll
          for (int cch = 0; cch < D / 32; ++cch) tmem_ld32_nowait(tm_o + lane_off + 32u * cch, o + 32 * cch);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < D; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * factor);
#pragma unroll
          for (int cch = 0; cch < D / 32; ++cch) tmem_st32_nowait(tm_o + lane_off + 32u * cch, o + 32 * cch);
          tmem_st_wait();
        }
      }
      const uint32_t prow = p_s + (uint32_t)lrow * 16u;
#pragma unroll
      for (int g = 0; g < TK / 8; ++g) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float p0 = ex2(__uint_as_float(v[8 * g + 2 * e]) - m_ref), p1 = ex2(__uint_as_float(v[8 * g + 2 * e + 1]) - m_ref);
          l += p0 + p1;
          hi[e] = pack_f16x2(p0, p1);
          const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi[e]));
          lo[e] = pack_f16x2(p0 - hf.x, p1 - hf.y);
        }
        sts128(prow + (uint32_t)g * kPitch, hi[0], hi[1], hi[2], hi[3]);
        sts128(prow + kHalfP + (uint32_t)g * kPitch, lo[0], lo[1], lo[2], lo[3]);
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(B(9));
    }
    mbar_wait(B(10), (uint32_t)(nk - 1) & 1u, 70);
    tc_fence_after();
    uint32_t o[D];
#pragma unroll
    for (int cch = 0; cch < D / 32; ++cch) tmem_ld32_nowait(tm_o + lane_off + 32u * cch, o + 32 * cch);
    tmem_ld_wait();
    if (q0 + lrow < T) {
      const size_t r = (size_t)start + q0 + lrow;
      float4* dst = reinterpret_cast<float4*>(prm.o_main + r * kHidden + h * D);
#pragma unroll
      for (int e = 0; e < D / 4; ++e)
        dst[e] = make_float4(__uint_as_float(o[4 * e]), __uint_as_float(o[4 * e + 1]), __uint_as_float(o[4 * e + 2]), __uint_as_float(o[4 * e + 3]));
      prm.m_out[r * kHeads + h] = m_ref;
      prm.l_out[r * kHeads + h] = l;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// fp32 row-major qkv [R][576] -> planar fp16 hi / lo [72][R][8]   (x = hi + lo to 22 bits)
__global__ void qkv_to_planar_hilo_kernel(const float* __restrict__ x, __half* __restrict__ hi, __half* __restrict__ lo, int R) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;      // one thread per (plane, row)
  if (i >= kQkvPlanes * R) return;
  const int pl = i / R, r = i % R;
  const float4 a = *reinterpret_cast<const float4*>(x + (size_t)r * 3 * kHidden + pl * 8);
  const float4 c = *reinterpret_cast<const float4*>(x + (size_t)r * 3 * kHidden + pl * 8 + 4);
  const float f[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
  uint32_t h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    h[e] = pack_f16x2(f[2 * e], f[2 * e + 1]);
    const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&h[e]));
    l[e] = pack_f16x2(f[2 * e] - hf.x, f[2 * e + 1] - hf.y);
  }
  *reinterpret_cast<uint4*>(hi + (size_t)i * 8) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(lo + (size_t)i * 8) = make_uint4(l[0], l[1], l[2], l[3]);
}

// The 9-wide band: relative-key scores, relative-value term, normalisation.  One warp per (row, head); lane owns channels
// lane, lane + 32, lane + 64.
__global__ void __launch_bounds__(128) rel_band_fixup_kernel(VsRows rows, const float* __restrict__ qkv, const float* __restrict__ ek,
                                                             const float* __restrict__ ev, const float* __restrict__ o_main,
                                                             const float* __restrict__ m_in, const float* __restrict__ l_in,
                                                             float* __restrict__ out, int R) {
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x % 32;
  if (gw >= R * kHeads) return;
  const int r = gw / kHeads, h = gw % kHeads;
  const int u = rows.row_utt[r];
  float* dst = out + (size_t)r * kHidden + h * D;
  if (u < 0) {                                                   // gap rows of the output are zero
#pragma unroll
    for (int c = 0; c < 3; ++c) dst[lane + 32 * c] = 0.f;
    return;
  }
  const int start = rows.utt_start[u], T = rows.utt_len[u], i = r - start;
  const int ld = 3 * kHidden;
  const float scale = rsqrtf((float)D), log2e = 1.4426950408889634f;
  float qs[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) qs[c] = qkv[(size_t)r * ld + h * D + lane + 32 * c] * scale;
  float s2[kRel], b2[kRel];                                      // plain and relative-key scores of the band, log2 domain
  bool ok[kRel];
#pragma unroll
  for (int w = 0; w < kRel; ++w) {
    const int j = i + w - kWindow;
    ok[w] = j >= 0 && j < T;
    float ps = 0.f, pb = 0.f;
    if (ok[w]) {
      const float* kr = qkv + (size_t)(start + j) * ld + kHidden + h * D;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        ps = fmaf(qs[c], kr[lane + 32 * c], ps);
        pb = fmaf(qs[c], __ldg(ek + w * D + lane + 32 * c), pb);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ps += __shfl_xor_sync(0xffffffffu, ps, o);
      pb += __shfl_xor_sync(0xffffffffu, pb, o);
    }
    s2[w] = ps * log2e;
    b2[w] = pb * log2e;
  }
  const float m = m_in[(size_t)r * kHeads + h], l = l_in[(size_t)r * kHeads + h];
  float m2 = m;
#pragma unroll
  for (int w = 0; w < kRel; ++w)
    if (ok[w]) m2 = fmaxf(m2, s2[w] + b2[w]);
  const float a = exp2f(m - m2);
  float den = l * a, acc[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) acc[c] = o_main[(size_t)r * kHidden + h * D + lane + 32 * c] * a;
#pragma unroll
  for (int w = 0; w < kRel; ++w) {
    if (!ok[w]) continue;
    const float e1 = exp2f(s2[w] + b2[w] - m2), e0 = exp2f(s2[w] - m2);
    den += e1 - e0;
    const float* vr = qkv + (size_t)(start + i + w - kWindow) * ld + 2 * kHidden + h * D;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      acc[c] = fmaf(e1 - e0, vr[lane + 32 * c], fmaf(e1, __ldg(ev + w * D + lane + 32 * c), acc[c]));
  }
  const float inv = 1.f / den;
#pragma unroll
  for (int c = 0; c < 3; ++c) dst[lane + 32 * c] = acc[c] * inv;
}

}  // namespace

int64_t attention_umma_ws_floats(int R) {
  // planar hi + lo (R x 576 halves each = R x 576 floats together) + O (R x 192) + m, l (R x 2 each), 256-byte aligned pieces
  return (int64_t)R * (3 * kHidden + kHidden + 2 * kHeads) + 5 * 64;
}

int rel_attention_umma(const VsRows& rows, const float* qkv, const float* ek, const float* ev, float* out, Workspace& ws,
                       cudaStream_t st) {
  const int R = rows.n_rows;
  VS_REQUIRE(rows.max_len > 0 && rows.max_len <= R, "rel_attention_umma: bad max_len %d", rows.max_len);
  __half* hi = ws.take<__half>((int64_t)R * 3 * kHidden);
  __half* lo = ws.take<__half>((int64_t)R * 3 * kHidden);
  float* o_main = ws.take<float>((int64_t)R * kHidden);
  float* m = ws.take<float>((int64_t)R * kHeads);
  float* l = ws.take<float>((int64_t)R * kHeads);
  if (!ws.ok) { set_error("rel_attention_umma: workspace too small"); return VS_ERR_WORKSPACE; }
  qkv_to_planar_hilo_kernel<<<(kQkvPlanes * R + 255) / 256, 256, 0, st>>>(qkv, hi, lo, R);
  VS_LAUNCH_CHECK();
  VS_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(attention_umma_kernel), (int)kSmemBytes));
  Params prm;
  prm.rows = rows; prm.hi = hi; prm.lo = lo; prm.o_main = o_main; prm.m_out = m; prm.l_out = l; prm.R = R;
  dim3 grid((rows.max_len + TQ - 1) / TQ, kHeads, rows.n_utt);
  attention_umma_kernel<<<grid, kThreads, kSmemBytes, st>>>(prm);
  VS_LAUNCH_CHECK();
  rel_band_fixup_kernel<<<(R * kHeads * 32 + 127) / 128, 128, 0, st>>>(rows, qkv, ek, ev, o_main, m, l, out, R);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

}  // namespace vs
