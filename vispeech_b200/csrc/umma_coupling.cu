// One mean-only residual coupling layer of the flow (reference modules.py:324-343) with its whole WN stack (modules.py:148-176:
// 4 layers of k = 5 gated conv + 1x1 res/skip) as ONE tcgen05 kernel:
//
//     h = pre(x0) ; for l: acts = tanh . sigmoid (in_l(h) + cond_l(g)) ; [res | skip] = rs_l(acts) ; h += res ; skip_sum += skip
//     m = post(skip_sum) ;  x1 <- (x1 -/+ m) * mask
//
// Before: 9 launches per coupling layer (pre, layout change, 4 x WN layer, layout change, post, update), each WN layer re-staging
// a 101 KB fp32 tile: 374 us per coupling layer at the C2 size against ~50 us of tensor time.  Here a CTA owns a 128-row tile
// (112 rows written: the 4 x 2 rows of receptive field at either end are recomputed) for the whole layer:
//   * the residual stream h (192 columns) and m (96 columns) live in TMEM in fp32 for the whole stack: res_skip accumulates
//     straight onto them (the residual add and the skip sum are free and never rounded); the skip tensor itself is never
//     formed - post is folded into the skip half of every res_skip (m = sum_l acts_l (W_skip_l W_post), packing.py pack_coupling);
//   * only the conv OPERANDS are fp16 in shared memory, written by the epilogues in the UMMA K-major layout: h16 (the next
//     in_layer's input, masked: gap rows are the conv's zero padding) and acts16; fp16 has TF32's 11-bit significand, which is the
//     precision regime this path already ran in (calls >= tf32_min_rows rows; smaller calls keep the fp32-accurate kernels);
//   * in_layer runs as four n-blocks of 96 gate-interleaved columns ([48 tanh | 48 sigmoid]) into two alternating TMEM
//     accumulators, so the gate epilogue of block nb overlaps the MMAs of block nb + 1;
//   * weights stream as 36 KB slabs ([K = 192][N = 96] halves: one tap of one n-block, 12 MMAs) through a 3-slot TMA ring out
//     of L2 in a fixed order (92 slabs per tile);
//   * bias + per-speaker cond_layer(g) of every layer come from one fp32 table row per (speaker, layer).
// TMEM: h [0,192) m [192,288) acc0 [288,384) acc1 [384,480).  Roles: warp 0 weight producer, warp 1 MMA issuer, warps 2-9
// epilogue crew (x0 staging, h -> h16, gate, final update).
#include "umma_conv.cuh"
#include "umma_common.cuh"
#include "umma_tf32.cuh"

namespace vs {
namespace {

using namespace umma;

constexpr int H = 192, HALF = 96, TAPS = 5, L = 4, NBLK = 96;
constexpr int kHalo = 2 * L, kValid = kTileM - 2 * kHalo;                 // 8, 112
constexpr int kEpiWarps = 8, kThreads = 32 * (2 + kEpiWarps);
constexpr int ROWS_H = kTileM + 4;                                        // 2 zero pad rows either side of h16
constexpr uint32_t kSlabBytes = 24u * NBLK * 16u;                         // 36,864
constexpr uint32_t kPreBytes = 12u * NBLK * 16u;                          // K = 96
constexpr int kSlabsPerTile = 2 + L * (4 * TAPS) + (L - 1) * 3 + 1;       // 92
constexpr int kWSlots = 3;
constexpr uint32_t OFF_H16 = 0, H16_BYTES = 24u * ROWS_H * 16u;           // 50,688
constexpr uint32_t OFF_ACT = OFF_H16 + H16_BYTES, ACT_BYTES = 24u * kTileM * 16u;   // 49,152 (x0_16: its first 12 planes)
constexpr uint32_t OFF_W = OFF_ACT + ACT_BYTES;
constexpr uint32_t OFF_BAR = OFF_W + kWSlots * kSlabBytes;
constexpr int kNumBars = 2 * kWSlots + 8;
constexpr uint32_t SMEM_BYTES = OFF_BAR + 8u * kNumBars + 16u;
static_assert(SMEM_BYTES <= 227u * 1024, "umma_coupling: shared memory");
constexpr uint32_t TM_H = 0, TM_M = 192, TM_ACC = 288;

struct Params {
  UmmaCoupling c;
  int n_tiles;
  long long* dbg;        // wait-clock counters of the MMA issuer (-DVS_UMMA_TIMING builds, option "umma_timing_buffer")
};
#ifdef VS_UMMA_TIMING
#define VS_TIMED(var, stmt)                         \
  do {                                              \
    const long long _t0 = prm.dbg ? clock64() : 0;  \
    stmt;                                           \
    if (prm.dbg) var += clock64() - _t0;            \
  } while (0)
#else
#define VS_TIMED(var, stmt) stmt
#endif

__device__ __forceinline__ float gate_fast(float a, float b) {
  const float ea = __expf(-2.f * fminf(fmaxf(a, -40.f), 40.f));     // clamped: e^80 stays finite in fp32
  const float eb = __expf(-fminf(fmaxf(b, -80.f), 80.f));
  return __fdividef(1.f - ea, (1.f + ea) * (1.f + eb));
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(kThreads, 1) umma_coupling_kernel(const __grid_constant__ Params prm) {
  extern __shared__ __align__(128) uint8_t smem[];
  pdl_trigger();
  const UmmaCoupling& c = prm.c;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t h16 = smem_base + OFF_H16, act16 = smem_base + OFF_ACT, w_base = smem_base + OFF_W, bar = smem_base + OFF_BAR;
  auto w_full = [&](uint32_t i) { return bar + 8u * i; };
  auto w_empty = [&](uint32_t i) { return bar + 8u * (kWSlots + i); };
  const uint32_t x0_ready = bar + 8u * (2 * kWSlots), h_done = x0_ready + 8u, h16_ready = x0_ready + 16u, acts_ready = x0_ready + 24u;
  auto acc_full = [&](uint32_t i) { return x0_ready + 32u + 8u * i; };
  auto acc_empty = [&](uint32_t i) { return x0_ready + 48u + 8u * i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR + 8 * kNumBars);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kWSlots; ++i) { mbar_init(w_full(i), 1); mbar_init(w_empty(i), 1); }
    mbar_init(x0_ready, kEpiWarps); mbar_init(h_done, 1); mbar_init(h16_ready, kEpiWarps); mbar_init(acts_ready, kEpiWarps);
    for (int i = 0; i < 2; ++i) { mbar_init(acc_full(i), 1); mbar_init(acc_empty(i), kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // the pad rows of h16 (2 at either end of every plane) stay zero: rows outside the tile read as zero padding
  for (int i = threadIdx.x; i < 24 * 4; i += kThreads) {
    const int pl = i / 4, r = i % 4;
    sts128(h16 + (uint32_t)(pl * ROWS_H + (r < 2 ? r : ROWS_H - 4 + r)) * 16u, 0u, 0u, 0u, 0u);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();                              // the prologue above overlapped the previous kernel's tail
  const uint32_t tmem = *tmem_slot;
  const int R = c.R;

  if (warp == 0) {
    // ------------------------------------------------------------------ weight producer: the same 92 slabs for every tile
    if (lane == 0) {
      uint32_t wi = 0;
      const uint8_t* src0 = reinterpret_cast<const uint8_t*>(c.w);
      for (int tile = blockIdx.x; tile < prm.n_tiles; tile += gridDim.x)
        for (int s = 0; s < kSlabsPerTile; ++s, ++wi) {
          const uint32_t slot = wi % kWSlots, ph = (wi / kWSlots) & 1u;
          const uint32_t bytes = s < 2 ? kPreBytes : kSlabBytes;
          mbar_wait(w_empty(slot), ph ^ 1u, 1);
          mbar_arrive_expect_tx(w_full(slot), bytes);
          bulk_g2s(w_base + slot * kSlabBytes, src0 + (size_t)s * kSlabBytes, bytes, w_full(slot));
        }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = make_idesc(NBLK);
    const uint32_t b_hi = (uint32_t)(make_desc(0, NBLK * 16u, 128u) >> 32), b_fixed = (uint32_t)make_desc(0, NBLK * 16u, 128u);
    const uint32_t ah_hi = (uint32_t)(make_desc(0, ROWS_H * 16u, 128u) >> 32), ah_fixed = (uint32_t)make_desc(0, ROWS_H * 16u, 128u);
    const uint32_t aa_hi = (uint32_t)(make_desc(0, kTileM * 16u, 128u) >> 32), aa_fixed = (uint32_t)make_desc(0, kTileM * 16u, 128u);
    constexpr uint32_t b_kstep = 2u * NBLK, ah_kstep = 2u * ROWS_H, aa_kstep = 2u * kTileM;
    uint32_t wi = 0, n_h16 = 0, n_acts = 0, n_acc[2] = {0, 0}, n_tile = 0;
#ifdef VS_UMMA_TIMING
    long long tw_w = 0, tw_h16 = 0, tw_acc = 0, tw_acts = 0, tw_x0 = 0;
    const long long t_start = clock64();
#endif
    // one slab = NK MMAs (K = 16 each) of A (start a_lo, K step a_kstep) against the slab, into d
    auto slab_mmas = [&](uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t a_kstep, int nk, uint32_t accumulate) {
      const uint32_t slot = wi % kWSlots, ph = (wi / kWSlots) & 1u;
      VS_TIMED(tw_w, mbar_wait(w_full(slot), ph, 2));
      tc_fence_after();
      uint32_t b_lo = b_fixed + ((w_base + slot * kSlabBytes) >> 4);
#pragma unroll 4
      for (int k = 0; k < nk; ++k, a_lo += a_kstep, b_lo += b_kstep) {
        tc_mma_f16_lohi(d, a_lo, a_hi, b_lo, b_hi, idesc, accumulate);
        accumulate = 1;
      }
      tc_commit(w_empty(slot));
      ++wi;
    };
    for (int tile = blockIdx.x; tile < prm.n_tiles; tile += gridDim.x, ++n_tile) {
      VS_TIMED(tw_x0, mbar_wait(x0_ready, n_tile & 1u, 3));
      tc_fence_after();
      for (int s = 0; s < 2; ++s)                                        // h = pre(x0)
        slab_mmas(tmem + TM_H + (uint32_t)s * NBLK, aa_fixed + (act16 >> 4), aa_hi, aa_kstep, HALF / 16, 0u);
      tc_commit(h_done);
      for (int l = 0; l < L; ++l) {
        VS_TIMED(tw_h16, mbar_wait(h16_ready, n_h16 & 1u, 4));
        ++n_h16;
        tc_fence_after();
        for (int nb = 0; nb < 4; ++nb) {                                 // in_layer, n-block nb -> acc[nb & 1]
          const uint32_t a = nb & 1;
          VS_TIMED(tw_acc, mbar_wait(acc_empty(a), (n_acc[a] & 1u) ^ 1u, 5));
          ++n_acc[a];
          tc_fence_after();
          const uint32_t d = tmem + TM_ACC + a * NBLK;
          for (int t = 0; t < TAPS; ++t) slab_mmas(d, ah_fixed + ((h16 + (uint32_t)t * 16u) >> 4), ah_hi, ah_kstep, H / 16, t ? 1u : 0u);
          tc_commit(acc_full(a));
        }
        VS_TIMED(tw_acts, mbar_wait(acts_ready, n_acts & 1u, 6));
        ++n_acts;
        tc_fence_after();
        if (l < L - 1) {                                                  // h += res (the residual add happens in the accumulator)
          for (int s = 0; s < 2; ++s) slab_mmas(tmem + TM_H + (uint32_t)s * NBLK, aa_fixed + (act16 >> 4), aa_hi, aa_kstep, H / 16, 1u);
          tc_commit(h_done);                                              // the crew turns h into the next operand while m is updated
        }
        slab_mmas(tmem + TM_M, aa_fixed + (act16 >> 4), aa_hi, aa_kstep, H / 16, l ? 1u : 0u);     // m += acts (W_skip W_post)
        if (l == L - 1) tc_commit(h_done);
      }
    }
#ifdef VS_UMMA_TIMING
    if (prm.dbg && lane == 0) {
      long long* o = prm.dbg + (size_t)blockIdx.x * 8;
      o[0] = clock64() - t_start; o[1] = tw_w; o[2] = tw_h16; o[3] = tw_acc; o[4] = tw_acts; o[5] = tw_x0;
    }
#endif
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue crew: thread = one row, two warps per lane quarter
    const int q = warp & 3, hh = (warp - 2) >> 2;                        // hh: which half of the columns this warp takes
    const int j = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const float* hb = c.bias;                                            // [L][192]
    const float* mb = c.bias + L * H;                                    // [96]
    const float* cg = c.bias + L * H + HALF;                             // [n_spk][L][384]
    uint32_t n_hdone = 0, n_acc[2] = {0, 0};
    // this warp's 48 channels of x0 of a tile's row, fetched one tile ahead (the loads of tile n + 1 fly during tile n's last layer)
    float4 xr[12];
    auto fetch_x0 = [&](int tile_) {
      const int g_ = tile_ * kValid - kHalo + j;
      const bool ok = g_ >= 0 && g_ < R && c.row_utt[g_] >= 0;
      const float4* src = reinterpret_cast<const float4*>(c.z + (size_t)(ok ? g_ : 0) * H + c.in_off + 48 * hh);
#pragma unroll
      for (int i = 0; i < 12; ++i) xr[i] = ok ? src[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    if ((int)blockIdx.x < prm.n_tiles) fetch_x0(blockIdx.x);
    for (int tile = blockIdx.x; tile < prm.n_tiles; tile += gridDim.x) {
      const int g = tile * kValid - kHalo + j;
      const bool in_seq = g >= 0 && g < R;
      const int utt = in_seq ? c.row_utt[g] : -1;
      const bool valid = utt >= 0;
      const uint32_t keep = valid ? 0xFFFFFFFFu : 0u;
      // ---- x0 (fp32 rows) -> x0_16 in the first 12 planes of the acts buffer
      {
#pragma unroll
        for (int p6 = 0; p6 < 6; ++p6) {
          const float4 a = xr[2 * p6], b = xr[2 * p6 + 1];
          sts128(act16 + (uint32_t)((6 * hh + p6) * kTileM + j) * 16u, pack_f16x2(a.x, a.y), pack_f16x2(a.z, a.w), pack_f16x2(b.x, b.y),
                 pack_f16x2(b.z, b.w));
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(x0_ready);
      }
      const float* cgrow = cg + (size_t)(valid ? c.sid[utt] : 0) * (L * 2 * H);
      for (int l = 0; l < L; ++l) {
        // ---- h (TMEM, fp32) + cumulative bias -> masked f16 operand of in_layer l; this warp's 96 columns
        mbar_wait(h_done, n_hdone & 1u, 7);
        ++n_hdone;
        tc_fence_after();
        {
          const float* b = hb + l * H + 96 * hh;
          const uint32_t src = tmem + lane_off + TM_H + 96u * hh;
          const uint32_t dst = h16 + (uint32_t)((12 * hh) * ROWS_H + 2 + j) * 16u;
#pragma unroll
          for (int cc = 0; cc < 3; ++cc) {
            uint32_t v[32];
            tmem_ld32(src + 32u * cc, v);
#pragma unroll
            for (int gq = 0; gq < 4; ++gq) {
              float y[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) y[e] = __uint_as_float(v[8 * gq + e]) + __ldg(b + 32 * cc + 8 * gq + e);
              sts128(dst + (uint32_t)((4 * cc + gq) * ROWS_H) * 16u, pack_f16x2(y[0], y[1]) & keep, pack_f16x2(y[2], y[3]) & keep,
                     pack_f16x2(y[4], y[5]) & keep, pack_f16x2(y[6], y[7]) & keep);
            }
          }
          fence_proxy_async();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(h16_ready);
        }
        // ---- gate epilogues: acc[nb & 1] (96 columns = 48 tanh | 48 sigmoid) -> acts16 planes 6 nb .. 6 nb + 5; this warp's 24 channels
        for (int nb = 0; nb < 4; ++nb) {
          const uint32_t a = nb & 1;
          const float* bt = cgrow + l * (2 * H) + 96 * nb + 24 * hh;        // tanh part; sigmoid part 48 further
          float bias[48];
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            const float4 t4 = __ldg(reinterpret_cast<const float4*>(bt) + i), s4 = __ldg(reinterpret_cast<const float4*>(bt + 48) + i);
            bias[4 * i] = t4.x; bias[4 * i + 1] = t4.y; bias[4 * i + 2] = t4.z; bias[4 * i + 3] = t4.w;
            bias[24 + 4 * i] = s4.x; bias[24 + 4 * i + 1] = s4.y; bias[24 + 4 * i + 2] = s4.z; bias[24 + 4 * i + 3] = s4.w;
          }
          mbar_wait(acc_full(a), n_acc[a] & 1u, 8);
          ++n_acc[a];
          tc_fence_after();
          const uint32_t src = tmem + lane_off + TM_ACC + a * NBLK + 24u * hh;
          uint32_t vt[24], vs_[24];
#pragma unroll
          for (int i = 0; i < 3; ++i) { tmem_ld8(src + 8u * i, vt + 8 * i); tmem_ld8(src + 48u + 8u * i, vs_ + 8 * i); }
          tmem_wait_ld();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty(a));                       // the accumulator is in registers
#pragma unroll
          for (int gq = 0; gq < 3; ++gq) {
            float y[8];
#pragma unroll
            for (int e = 0; e < 8; ++e)
              y[e] = gate_fast(__uint_as_float(vt[8 * gq + e]) + bias[8 * gq + e], __uint_as_float(vs_[8 * gq + e]) + bias[24 + 8 * gq + e]);
            sts128(act16 + (uint32_t)((6 * nb + 3 * hh + gq) * kTileM + j) * 16u, pack_f16x2(y[0], y[1]), pack_f16x2(y[2], y[3]),
                   pack_f16x2(y[4], y[5]), pack_f16x2(y[6], y[7]));
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(acts_ready);
      }
      // ---- m (TMEM) + bias -> x1 <- (x1 + sign m) * mask, rows this tile owns; this warp's 48 channels
      if (tile + (int)gridDim.x < prm.n_tiles) fetch_x0(tile + gridDim.x);      // x0 is the half of z this kernel does not write
      mbar_wait(h_done, n_hdone & 1u, 9);
      ++n_hdone;
      tc_fence_after();
      {
        const uint32_t src = tmem + lane_off + TM_M + 48u * hh;
        uint32_t v[48];
#pragma unroll
        for (int i = 0; i < 6; ++i) tmem_ld8(src + 8u * i, v + 8 * i);
        tmem_wait_ld();
        tc_fence_before();
        if (valid && j >= kHalo && j < kTileM - kHalo) {
          float4* x1 = reinterpret_cast<float4*>(c.z + (size_t)g * H + c.upd_off + 48 * hh);
          const float4* b4 = reinterpret_cast<const float4*>(mb + 48 * hh);
#pragma unroll
          for (int i = 0; i < 12; ++i) {
            float4 x = x1[i];
            const float4 b = __ldg(b4 + i);
            x.x += c.sign * (__uint_as_float(v[4 * i]) + b.x); x.y += c.sign * (__uint_as_float(v[4 * i + 1]) + b.y);
            x.z += c.sign * (__uint_as_float(v[4 * i + 2]) + b.z); x.w += c.sign * (__uint_as_float(v[4 * i + 3]) + b.w);
            x1[i] = x;
          }
        }
      }
      // the next tile's x0 staging overwrites the acts buffer: every MMA that read it has completed (h_done of the last layer)
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

}  // namespace

int umma_coupling(const UmmaCoupling& c, cudaStream_t st) {
  VS_REQUIRE(c.z && c.w && c.bias && c.row_utt && c.sid && c.R > 0, "umma_coupling: null pointer");
  VS_REQUIRE((c.in_off == 0 && c.upd_off == HALF) || (c.in_off == HALF && c.upd_off == 0), "umma_coupling: bad channel halves");
  Params prm;
  prm.c = c;
  prm.n_tiles = (c.R + kValid - 1) / kValid;
  prm.dbg = static_cast<long long*>(umma_conv_timing_buffer());
  int n_sm = 0;
  VS_TRY(device_sm_count(&n_sm));
  VS_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(umma_coupling_kernel), (int)SMEM_BYTES));
  const int grid = prm.n_tiles < n_sm ? prm.n_tiles : n_sm;
  VS_CUDA_CHECK(launch_pdl<8>(umma_coupling_kernel, dim3(grid), dim3(kThreads), SMEM_BYTES, st, prm));
  VS_LAUNCH_CHECK();
  return VS_OK;
}

}  // namespace vs
