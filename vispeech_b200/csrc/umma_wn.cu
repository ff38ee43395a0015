// One WN layer (reference modules.py:148-176, the body of the flow's coupling layers and of the posterior encoder) as ONE
// tcgen05 kernel, plain TF32 (calls with >= tf32_min_rows rows; smaller calls keep the two-kernel 3xTF32 path):
//
//     acts = tanh(in_t(h) + g_t) * sigmoid(in_s(h) + g_s)         in_layer: k = 5, 192 -> 384        (commons.py:100-107)
//     rs   = res_skip(acts)                                       1x1, 192 -> 384 (192 in the last layer)
//     h'   = (h + rs[:, :192]) * mask ;  skip (+)= rs[:, 192:]
//
// umma_tf32.cu runs this as two launches (in_layer + gate epilogue, res_skip + update epilogue) with `acts` travelling
// through HBM and - the expensive part - the activation tile staged (fp32 -> TF32, K-major smem) once per 192-column n-block:
// its MMA warp waits for the loaders 40-64 % of the time.  Here, per 128-row tile:
//   loaders   stage h (132 rows x 192 channels) ONCE                                            -> smem A (2 x 96 channels)
//   MMA       in_layer for BOTH n-blocks against that tile                                      -> TMEM acc[0..383]
//   epilogue  gate (+ bias + per-speaker cond) -> TF32 -> smem, in the K-major layout, OVER the dead input tile
//   MMA       res_skip against that smem tile                                                   -> TMEM acc (same columns)
//   epilogue  h' = h + res (h is double-buffered: neighbouring tiles still read h's halo rows), skip (+)= skip
// Weight slabs (24 KB: 32 channels x 192 columns) stream through a 5-deep TMA ring for both GEMMs.
// h and skip live in a PLANAR fp32 layout [C/4][R][4] inside the WN stack (rows_to_planar4 / planar4_to_rows at its ends):
// a plane-slab of a row tile is contiguous in HBM and IS one K-chunk column of the K-major smem layout, so the input tile
// arrives by 24 plain TMA bulk copies per stage (no loader warps, no st.shared; the MMA reads the top 19 bits of the fp32
// words), and the update epilogue's per-row float4 accesses are 512 contiguous bytes per warp instruction instead of 32
// different 128-byte lines (the row-major form of this kernel spent 20 % of its time there and 13 % in the loaders).
#include "umma_tf32.cuh"
#include "umma_common.cuh"
#include "umma_conv.cuh"

namespace vs {
namespace {

using namespace umma;

constexpr int kLoaderWarps = 4, kEpiWarps = 8;
constexpr int kThreads = 32 * (kLoaderWarps + 2 + kEpiWarps);
constexpr int H = kHidden, TAPS = 5, HALO = 2, ROWS_A = kTileM + TAPS - 1;     // 132
constexpr int NBLK = 192, SB = 5, kPF = 24;     // a whole 96-channel stage is in flight per thread while it waits for the region
constexpr uint32_t A_STAGE = 24u * ROWS_A * 16u;                 // 96 channels = 24 planes of [rows][4 floats]
constexpr uint32_t A_BYTES = 2u * A_STAGE;                       // 101,376
constexpr uint32_t ACT_PLANE = kTileM * 16u;                     // acts: 48 planes x 128 rows x 16 B = 98,304 <= A_BYTES
constexpr uint32_t B_SLAB = 32u * NBLK * 4u;                     // 24,576
constexpr uint32_t OFF_B = A_BYTES, OFF_BAR = OFF_B + SB * B_SLAB;
constexpr uint32_t SMEM_BYTES = OFF_BAR + 8u * (4 + 2 * SB + 4) + 16u;
static_assert(48u * ACT_PLANE <= A_BYTES, "acts must fit over the input stages");
static_assert(SMEM_BYTES <= 227u * 1024, "smem");

struct Params { UmmaWn c; int n_tiles; long long* dbg; };   // dbg: wait-clock counters of the MMA warp (-DVS_UMMA_TIMING only)
#ifdef VS_UMMA_TIMING
#define WN_TIMED(var, stmt)                         \
  do {                                              \
    const long long _t0 = prm.dbg ? clock64() : 0;  \
    stmt;                                           \
    if (prm.dbg) var += clock64() - _t0;            \
  } while (0)
#else
#define WN_TIMED(var, stmt) stmt
#endif

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, pe;\n\t.reg .b64 da, db;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

__device__ __forceinline__ float gate_fast(float a, float b) {
  const float ea = __expf(-2.f * fminf(fmaxf(a, -40.f), 40.f));     // clamped: e^80 stays finite in fp32
  const float eb = __expf(-fminf(fmaxf(b, -80.f), 80.f));
  return __fdividef(1.f - ea, (1.f + ea) * (1.f + eb));
}

__global__ void __launch_bounds__(kThreads, 1) umma_wn_kernel(const __grid_constant__ Params prm) {
  extern __shared__ __align__(128) uint8_t smem[];
  const UmmaWn& c = prm.c;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t a_base = smem_base, b_base = smem_base + OFF_B, bar = smem_base + OFF_BAR;
  auto a_full = [&](int i) { return bar + 8u * i; };
  auto a_empty = [&](int i) { return bar + 8u * (2 + i); };
  auto b_full = [&](int i) { return bar + 8u * (4 + i); };
  auto b_empty = [&](int i) { return bar + 8u * (4 + SB + i); };
  const uint32_t acc1_full = bar + 8u * (4 + 2 * SB), acts_full = acc1_full + 8u, acc2_full = acc1_full + 16u,
                 acc2_empty = acc1_full + 24u;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR + 8 * (4 + 2 * SB + 4));

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(a_full(i), 1); mbar_init(a_empty(i), 1); }
    for (int i = 0; i < SB; ++i) { mbar_init(b_full(i), 1); mbar_init(b_empty(i), 1); }
    mbar_init(acc1_full, 1); mbar_init(acts_full, kEpiWarps); mbar_init(acc2_full, 1); mbar_init(acc2_empty, kEpiWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kLoaderWarps + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nb2 = c.last ? 1 : 2;                      // res_skip n-blocks: [res | skip], or skip only in the last layer

  if (warp < kLoaderWarps) {
    // ------------------------------------------------------------- input tile producer (warp 0): TMA bulk copies per plane
    if (warp == 0) {
      uint32_t i = 0;
      for (int tile = blockIdx.x; tile < prm.n_tiles; tile += gridDim.x, ++i) {
        const int row_lo = tile * kTileM - HALO, row_hi = row_lo + ROWS_A;
        const int c_lo = row_lo < 0 ? 0 : row_lo, c_hi = row_hi > c.R ? c.R : row_hi;
        const int n_zero_lo = c_lo - row_lo, n_zero_hi = row_hi - c_hi;
        const uint32_t bytes = (uint32_t)(c_hi - c_lo) * 16u;
        for (int ka = 0; ka < 2; ++ka) {
          mbar_wait(a_empty(ka), (i & 1u) ^ 1u, 41);
          const uint32_t stage = a_base + ka * A_STAGE;
          if (n_zero_lo > 0 || n_zero_hi > 0) {        // rows outside [0, R): the conv's zero padding
            const int per_plane = n_zero_lo + n_zero_hi;
            for (int k = lane; k < 24 * per_plane; k += 32) {
              const int pl = k / per_plane, j = k % per_plane;
              const int row = j < n_zero_lo ? j : (ROWS_A - n_zero_hi + (j - n_zero_lo));
              asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(stage + (uint32_t)(pl * ROWS_A + row) * 16u), "r"(0) : "memory");
            }
            fence_proxy_async();
          }
          __syncwarp();
          if (lane == 0) mbar_arrive_expect_tx(a_full(ka), 24u * bytes);
          __syncwarp();
          if (lane < 24)
            bulk_g2s(stage + (uint32_t)(lane * ROWS_A + n_zero_lo) * 16u, c.h_in + ((size_t)(ka * 24 + lane) * c.R + c_lo) * 4, bytes,
                     a_full(ka));
        }
      }
    }
  } else if (warp == kLoaderWarps) {
    // ------------------------------------------------------------- weight slabs (TMA bulk), in the MMA warp's order
    uint32_t slot = 0, phase = 0;
    auto fetch = [&](const float* src) {
      mbar_wait(b_empty(slot), phase ^ 1u, 42);
      if (lane == 0) {
        mbar_arrive_expect_tx(b_full(slot), B_SLAB);
        bulk_g2s(b_base + slot * B_SLAB, src, B_SLAB, b_full(slot));
      }
      if (++slot == SB) { slot = 0; phase ^= 1u; }
    };
    for (int tile = blockIdx.x; tile < prm.n_tiles; tile += gridDim.x) {
      for (int ka = 0; ka < 2; ++ka)
        for (int nb = 0; nb < 2; ++nb)
          for (int t = 0; t < TAPS; ++t)
            for (int j = 0; j < 3; ++j) fetch(c.w_in + (size_t)((((nb * 2 + ka) * TAPS + t) * 3 + j)) * (B_SLAB / 4));
      for (int nb = 0; nb < nb2; ++nb)
        for (int s6 = 0; s6 < 6; ++s6) fetch(c.w_rs + (size_t)(nb * 6 + s6) * (B_SLAB / 4));
    }
    __syncwarp();
  } else if (warp == kLoaderWarps + 1) {
    // ------------------------------------------------------------- MMA issuer
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NBLK >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
    const uint32_t a1_lbo = ROWS_A * 16u, a2_lbo = ACT_PLANE, b_lbo = NBLK * 16u;
    const uint32_t a1_hi = (uint32_t)(make_desc(0, a1_lbo, 128u) >> 32), a1_fix = (uint32_t)make_desc(0, a1_lbo, 128u);
    const uint32_t a2_hi = (uint32_t)(make_desc(0, a2_lbo, 128u) >> 32), a2_fix = (uint32_t)make_desc(0, a2_lbo, 128u);
    const uint32_t b_hi = (uint32_t)(make_desc(0, b_lbo, 128u) >> 32), b_fix = (uint32_t)make_desc(0, b_lbo, 128u);
    constexpr uint32_t a1_kstep = 2u * ROWS_A, a2_kstep = 2u * kTileM, b_kstep = 2u * NBLK;     // two planes, in 16-byte units
    uint32_t slot = 0, phase = 0, i = 0;
#ifdef VS_UMMA_TIMING
    long long tw_a = 0, tw_b = 0, tw_acts = 0, tw_acc2 = 0;
    const long long t_start = prm.dbg ? clock64() : 0;
#endif
    for (int tile = blockIdx.x; tile < prm.n_tiles; tile += gridDim.x, ++i) {
      const uint32_t par = i & 1u;
      WN_TIMED(tw_acc2, mbar_wait(acc2_empty, par ^ 1u, 43));   // the previous tile's res/skip accumulators have been drained
      tc_fence_after();
      for (int ka = 0; ka < 2; ++ka) {
        WN_TIMED(tw_a, mbar_wait(a_full(ka), par, 44));
        tc_fence_after();
        const uint32_t a_stage16 = (a_base + ka * A_STAGE) >> 4;
        for (int nb = 0; nb < 2; ++nb) {
          const uint32_t d = tmem_base + (uint32_t)nb * NBLK;
          for (int t = 0; t < TAPS; ++t)
            for (int j = 0; j < 3; ++j) {
              WN_TIMED(tw_b, mbar_wait(b_full(slot), phase, 45));
              tc_fence_after();
              const uint32_t a_lo = a1_fix + a_stage16 + (uint32_t)j * 8u * ROWS_A + (uint32_t)t;
              const uint32_t b_lo = b_fix + ((b_base + slot * B_SLAB) >> 4);
              const uint32_t first = (ka == 0 && t == 0 && j == 0) ? 0u : 1u;
#pragma unroll
              for (int k8 = 0; k8 < 4; ++k8)
                mma_tf32(d, a_lo + k8 * a1_kstep, a1_hi, b_lo + k8 * b_kstep, b_hi, idesc, k8 ? 1u : first);
              tc_commit(b_empty(slot));
              if (++slot == SB) { slot = 0; phase ^= 1u; }
            }
        }
      }
      tc_commit(acc1_full);
      WN_TIMED(tw_acts, mbar_wait(acts_full, par, 46));   // gate output is in smem (over the input tile), acc columns are free
      tc_fence_after();
      const uint32_t act16 = a_base >> 4;
      for (int nb = 0; nb < nb2; ++nb) {
        const uint32_t d = tmem_base + (uint32_t)nb * NBLK;
        for (int s6 = 0; s6 < 6; ++s6) {
          WN_TIMED(tw_b, mbar_wait(b_full(slot), phase, 47));
          tc_fence_after();
          const uint32_t a_lo = a2_fix + act16 + (uint32_t)s6 * 8u * kTileM;
          const uint32_t b_lo = b_fix + ((b_base + slot * B_SLAB) >> 4);
#pragma unroll
          for (int k8 = 0; k8 < 4; ++k8)
            mma_tf32(d, a_lo + k8 * a2_kstep, a2_hi, b_lo + k8 * b_kstep, b_hi, idesc, (s6 == 0 && k8 == 0) ? 0u : 1u);
          tc_commit(b_empty(slot));
          if (++slot == SB) { slot = 0; phase ^= 1u; }
        }
      }
      tc_commit(acc2_full);
      tc_commit(a_empty(0));                         // the activation region may receive the next tile
      tc_commit(a_empty(1));
    }
#ifdef VS_UMMA_TIMING
    if (prm.dbg && lane == 0) {
      long long* o = prm.dbg + 148 * 16 + (size_t)blockIdx.x * 8;    // behind the region the other tcgen05 kernels use
      o[0] = clock64() - t_start; o[1] = tw_a; o[2] = tw_b; o[3] = tw_acts; o[4] = tw_acc2; o[5] = i;
    }
#endif
  } else {
    // ------------------------------------------------------------- epilogues (8 warps): gate, then res/skip update
    const int ew = warp - (kLoaderWarps + 2);
    const int q = warp & 3, hsel = ew >> 2;
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    const int rl = q * 32 + lane;                      // row within the tile
    uint32_t i = 0;
    for (int tile = blockIdx.x; tile < prm.n_tiles; tile += gridDim.x, ++i) {
      const uint32_t par = i & 1u;
      const int r = tile * kTileM + rl;
      const bool in_range = r < c.R;
      int utt = -1;
      if (in_range) utt = c.row_utt ? c.row_utt[r] : 0;
      const bool valid = utt >= 0;
      const float* ub = nullptr;
      if (valid && c.cond) ub = c.cond + (size_t)(c.cond_idx ? c.cond_idx[utt] : utt) * c.cond_ld;
      // ---- gate: chunk cc = [16 tanh pre-activations | 16 sigmoid pre-activations] of channels 16cc..16cc+15
      mbar_wait(acc1_full, par, 48);
      tc_fence_after();
      for (int cc = hsel; cc < 12; cc += 2) {
        const int col0 = cc * 32;
        // bias + per-speaker cond of this chunk first: the st.shared below are asm volatile, so loads issued inside the g loop
        // would each wait out a full L1/L2 round trip (measured: 17 k clk per tile in this epilogue before, tools/wn_timing.py)
        float4 bsum[8];                              // [0..3] tanh half, [4..7] sigmoid half
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          bsum[g] = __ldg(reinterpret_cast<const float4*>(c.b_in + col0 + 4 * g));
          if (ub) {
            const float4 u4 = __ldg(reinterpret_cast<const float4*>(ub + col0 + 4 * g));
            bsum[g].x += u4.x; bsum[g].y += u4.y; bsum[g].z += u4.z; bsum[g].w += u4.w;
          }
        }
        uint32_t v[32];
        tmem_ld32(t_row + (uint32_t)(cc * 32), v);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
          if (valid) {
            const float4 bt = bsum[g], bs = bsum[4 + g];
            y.x = to_tf32(gate_fast(__uint_as_float(v[4 * g]) + bt.x, __uint_as_float(v[16 + 4 * g]) + bs.x));
            y.y = to_tf32(gate_fast(__uint_as_float(v[4 * g + 1]) + bt.y, __uint_as_float(v[16 + 4 * g + 1]) + bs.y));
            y.z = to_tf32(gate_fast(__uint_as_float(v[4 * g + 2]) + bt.z, __uint_as_float(v[16 + 4 * g + 2]) + bs.z));
            y.w = to_tf32(gate_fast(__uint_as_float(v[4 * g + 3]) + bt.w, __uint_as_float(v[16 + 4 * g + 3]) + bs.w));
          }
          // channels 16cc + 4g .. +3 = plane 4cc + g of the K-major tile
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a_base + (uint32_t)(4 * cc + g) * ACT_PLANE + (uint32_t)rl * 16u),
                       "f"(y.x), "f"(y.y), "f"(y.z), "f"(y.w)
                       : "memory");
        }
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acts_full);
      // ---- res/skip: h' = h + res (double-buffered), skip (+)= skip; gap rows: h' = 0, skip zeroed when assigned
      mbar_wait(acc2_full, par, 49);
      tc_fence_after();
      for (int cc = hsel; cc < 6 * nb2; cc += 2) {
        uint32_t v[32];
        tmem_ld32(t_row + (uint32_t)(cc * 32), v);
        if (!in_range) continue;
        const bool to_h = !c.last && cc < 6;
        const int col = (cc % 6) * 32;
        const float* bias = c.b_rs + cc * 32;
        // all loads of the chunk before its first store (the compiler may not move a load across a store to a possibly
        // aliasing pointer: one round trip per float4 otherwise - 20 k clk per tile in this epilogue before)
        float4 bb[8], old[8];
        const size_t plane = (size_t)c.R * 4;          // floats between consecutive 4-channel planes
        const size_t off = (size_t)(col / 4) * plane + (size_t)r * 4;
        const float* src = (to_h ? c.h_in : c.skip) + off;
        const bool need_old = valid && (to_h || !c.first);
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          bb[g] = __ldg(reinterpret_cast<const float4*>(bias + 4 * g));
          old[g] = need_old ? *reinterpret_cast<const float4*>(src + g * plane) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float* dst = (to_h ? c.h_out : c.skip) + off;
        if (valid) {
#pragma unroll
          for (int g = 0; g < 8; ++g)
            *reinterpret_cast<float4*>(dst + g * plane) =
                make_float4(__uint_as_float(v[4 * g]) + bb[g].x + old[g].x, __uint_as_float(v[4 * g + 1]) + bb[g].y + old[g].y,
                            __uint_as_float(v[4 * g + 2]) + bb[g].z + old[g].z, __uint_as_float(v[4 * g + 3]) + bb[g].w + old[g].w);
        } else if (to_h || c.first) {                  // gap rows: h' = 0; skip zeroed when assigned, untouched otherwise
#pragma unroll
          for (int g = 0; g < 8; ++g) *reinterpret_cast<float4*>(dst + g * plane) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc2_empty);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kLoaderWarps + 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// [R][C] fp32 row-major <-> planar [C/4][R][4]: a block moves 32 rows; both sides of the transpose are coalesced
__global__ void __launch_bounds__(256) rows_planar4_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int C,
                                                           int to_planar) {
  __shared__ float4 tile[32][49];                      // C / 4 <= 48 planes (+1: conflict-free column reads)
  const int r0 = blockIdx.x * 32, P = C / 4;
  if (to_planar) {
    for (int k = threadIdx.x; k < 32 * P; k += blockDim.x) {
      const int r = k / P, p = k % P;
      tile[r][p] = (r0 + r < R) ? *reinterpret_cast<const float4*>(in + (size_t)(r0 + r) * C + 4 * p) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 32 * P; k += blockDim.x) {
      const int p = k / 32, r = k % 32;
      if (r0 + r < R) *reinterpret_cast<float4*>(out + ((size_t)p * R + r0 + r) * 4) = tile[r][p];
    }
  } else {
    for (int k = threadIdx.x; k < 32 * P; k += blockDim.x) {
      const int p = k / 32, r = k % 32;
      tile[r][p] = (r0 + r < R) ? *reinterpret_cast<const float4*>(in + ((size_t)p * R + r0 + r) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 32 * P; k += blockDim.x) {
      const int r = k / P, p = k % P;
      if (r0 + r < R) *reinterpret_cast<float4*>(out + (size_t)(r0 + r) * C + 4 * p) = tile[r][p];
    }
  }
}

}  // namespace

int rows_to_planar4(const float* in, float* out, int R, int C, cudaStream_t st) {
  VS_REQUIRE(C % 4 == 0 && C <= 192 && R > 0, "rows_to_planar4: C=%d", C);
  rows_planar4_kernel<<<(R + 31) / 32, 256, 0, st>>>(in, out, R, C, 1);
  VS_LAUNCH_CHECK();
  return VS_OK;
}
int planar4_to_rows(const float* in, float* out, int R, int C, cudaStream_t st) {
  VS_REQUIRE(C % 4 == 0 && C <= 192 && R > 0, "planar4_to_rows: C=%d", C);
  rows_planar4_kernel<<<(R + 31) / 32, 256, 0, st>>>(in, out, R, C, 0);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

int umma_wn_layer(const UmmaWn& c, cudaStream_t st) {
  VS_REQUIRE(c.h_in && c.skip && c.w_in && c.b_in && c.w_rs && c.b_rs && (c.last || c.h_out), "umma_wn_layer: null pointer");
  VS_REQUIRE(c.R > 0 && c.h_in != c.h_out, "umma_wn_layer: h must be double-buffered");
  VS_REQUIRE(((reinterpret_cast<uintptr_t>(c.h_in) | reinterpret_cast<uintptr_t>(c.h_out) | reinterpret_cast<uintptr_t>(c.skip) |
               reinterpret_cast<uintptr_t>(c.b_in) | reinterpret_cast<uintptr_t>(c.b_rs) | reinterpret_cast<uintptr_t>(c.cond)) & 15) == 0 &&
                 c.cond_ld % 4 == 0,
             "umma_wn_layer: pointers must be 16-byte aligned");
  int n_sm = 0;
  VS_TRY(device_sm_count(&n_sm));
  VS_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(umma_wn_kernel), (int)SMEM_BYTES));
  Params prm;
  prm.c = c;
  prm.dbg = static_cast<long long*>(umma_conv_timing_buffer());
  prm.n_tiles = (c.R + kTileM - 1) / kTileM;
  const int grid = n_sm < prm.n_tiles ? n_sm : prm.n_tiles;
  umma_wn_kernel<<<grid, kThreads, SMEM_BYTES, st>>>(prm);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

}  // namespace vs
