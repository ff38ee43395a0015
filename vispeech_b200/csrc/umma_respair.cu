// One ResBlock1 iteration (reference modules.py:211-220) as ONE tcgen05 kernel for the narrow decoder stages:
//
//     y = c2( lrelu( c1( lrelu(x) ) + b1 ) ) + b2 + x        (c1: k taps, dilation d;  c2: k taps, dilation 1)
//
// The unfused chain (umma_conv.cu) moves 6 activation-sized tensors through HBM per iteration (read lrelu(x), write
// lrelu(c1), read it back, read x, write y and lrelu(y)); stages 2-3 (C = 64, 32) are HBM-bound on that traffic.
// Here only the raw stream x -> y touches HBM (2 tensors): lrelu(x) is produced in shared memory, the intermediate
// lrelu(c1) is written by the first epilogue straight into the second conv's A-operand tile in shared memory, and the
// residual is taken from the already staged x tile.
//
// Per CTA, per super tile of L = 128*MT conv1 rows (sequential phases; 2-3 CTAs per SM overlap each other):
//   producer : TMA bulk copies of raw x rows [s0-h1, s0+L+h1) (planar bf16), zero fill outside [0,R)        -> XR
//   epilogue : A1 = lrelu(XR)                         (smem -> smem, fence.proxy.async)
//   MMA      : conv1, MT interleaved accumulators     (taps = descriptor offsets t*d rows into A1)          -> TMEM
//   epilogue : TMEM -> +b1 -> lrelu -> validity mask -> bf16 -> A2 (the layout conv2 reads)
//   MMA      : conv2 over A2 (taps = offsets t)       outputs o in [0, L-2*h2) are valid                    -> TMEM
//   epilogue : TMEM -> +b2 + x (from XR) [+ running MRF sum] -> raw / activated outputs, 16 B coalesced stores
// Weights of both convs stay resident in shared memory for the whole kernel (C <= 64).
#include "umma_conv.cuh"
#include "umma_common.cuh"

namespace vs {
namespace {

using namespace umma;

constexpr int kThreads = 192;     // warp 0 producer, warp 1 MMA, warps 2..5 epilogue (one per TMEM lane quarter)
constexpr int kEpiWarps = 4;

struct Plan {
  int MT, L, Lout, h1, h2, planes, rows_x, rows_a2, n_super, tmem_cols, ctas_per_sm, row_div_shift;
  uint32_t xr_bytes, a2_bytes, w_bytes, smem_bytes;
  uint32_t off_a1, off_a2, off_w1, off_w2, off_bar, off_bias;
};
struct Params {
  UmmaPair c;
  Plan p;
};

__global__ void __launch_bounds__(kThreads, 3) umma_respair_kernel(const __grid_constant__ Params prm) {
  extern __shared__ __align__(128) uint8_t smem[];
  const UmmaPair& c = prm.c;
  const Plan& p = prm.p;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int N = c.C;

  const uint32_t smem_base = smem_u32(smem);
  const uint32_t xr = smem_base, a1 = smem_base + p.off_a1, a2 = smem_base + p.off_a2;
  const uint32_t w1 = smem_base + p.off_w1, w2 = smem_base + p.off_w2, bar = smem_base + p.off_bar;
  const uint32_t w_full = bar, xr_full = bar + 8, xr_empty = bar + 16, a1_full = bar + 24, acc1_full = bar + 32,
                 a2_full = bar + 40, acc2_full = bar + 48;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + p.off_bar + 64);
  float* bias_s = reinterpret_cast<float*>(smem + p.off_bias);      // [2][C]

  if (threadIdx.x == 0) {
    mbar_init(w_full, 1); mbar_init(xr_full, 1); mbar_init(xr_empty, kEpiWarps); mbar_init(a1_full, kEpiWarps);
    mbar_init(acc1_full, 1); mbar_init(a2_full, kEpiWarps); mbar_init(acc2_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < N; i += kThreads) { bias_s[i] = c.b1[i]; bias_s[N + i] = c.b2[i]; }
  // rows [L, L + 2*h2) of A2 are read by the last taps of conv2 (their outputs are discarded): keep them zero
  for (int i = threadIdx.x; i < p.planes * 2 * p.h2; i += kThreads) {
    const int pl = i / (2 * p.h2), j = i % (2 * p.h2);
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(a2 + (uint32_t)(pl * p.rows_a2 + p.L + j) * 16u), "r"(0)
                 : "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc1 = tmem_base, acc2 = tmem_base;     // same columns: conv2 starts after epilogue 1 drained them

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    if (lane == 0) {
      mbar_arrive_expect_tx(w_full, 2 * p.w_bytes);
      bulk_g2s(w1, c.w1, p.w_bytes, w_full);
      bulk_g2s(w2, c.w2, p.w_bytes, w_full);
    }
    uint32_t it = 0;
    for (int super = blockIdx.x; super < p.n_super; super += gridDim.x, ++it) {
      mbar_wait(xr_empty, (it & 1) ^ 1, 21);
      const int s0 = super * p.Lout - p.h2;
      const int row_lo = s0 - p.h1, row_hi = row_lo + p.rows_x;
      const int c_lo = row_lo < 0 ? 0 : row_lo, c_hi = row_hi > c.R ? c.R : row_hi;
      const int n_zero_lo = c_lo - row_lo, n_zero_hi = row_hi - c_hi;
      if (n_zero_lo > 0 || n_zero_hi > 0) {
        const int per_plane = n_zero_lo + n_zero_hi;
        for (int i = lane; i < p.planes * per_plane; i += 32) {
          const int pl = i / per_plane, j = i % per_plane;
          const int row = j < n_zero_lo ? j : (p.rows_x - n_zero_hi + (j - n_zero_lo));
          asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(xr + (uint32_t)(pl * p.rows_x + row) * 16u), "r"(0)
                       : "memory");
        }
      }
      __syncwarp();
      const int n_rows = c_hi - c_lo;                      // may be <= 0 for a tile hanging past the end
      const uint32_t bytes = n_rows > 0 ? (uint32_t)n_rows * 16u : 0u;
      if (lane == 0) mbar_arrive_expect_tx(xr_full, bytes * p.planes);
      __syncwarp();
      if (bytes)
        for (int pl = lane; pl < p.planes; pl += 32)
          bulk_g2s(xr + (uint32_t)(pl * p.rows_x + n_zero_lo) * 16u, c.x + ((size_t)pl * c.R + c_lo) * 8, bytes, xr_full);
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = make_idesc(N);
    const uint32_t b_lbo = (uint32_t)N * 16u;
    const uint32_t b_hi = (uint32_t)(make_desc(0, b_lbo, 128u) >> 32), b_lo_fixed = (uint32_t)make_desc(0, b_lbo, 128u);
    const uint32_t a1_lbo = (uint32_t)p.rows_x * 16u, a2_lbo = (uint32_t)p.rows_a2 * 16u;
    const uint32_t a1_hi = (uint32_t)(make_desc(0, a1_lbo, 128u) >> 32), a1_lo_fixed = (uint32_t)make_desc(0, a1_lbo, 128u);
    const uint32_t a2_hi = (uint32_t)(make_desc(0, a2_lbo, 128u) >> 32), a2_lo_fixed = (uint32_t)make_desc(0, a2_lbo, 128u);
    const uint32_t b_kstep = 2u * (uint32_t)N;
    const int ksteps = N / 16, taps = c.taps, MT = p.MT;
    mbar_wait(w_full, 0, 22);
    tc_fence_after();
    uint32_t it = 0;
    for (int super = blockIdx.x; super < p.n_super; super += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      // conv1: A1 rows j + t*d, MT independent accumulators interleaved per K-step
      mbar_wait(a1_full, ph, 23);
      tc_fence_after();
      {
        uint32_t b_lo = b_lo_fixed + (w1 >> 4);
        uint32_t a_tap = a1_lo_fixed + (a1 >> 4);
        uint32_t accumulate = 0;
        for (int t = 0; t < taps; ++t, a_tap += (uint32_t)c.dil) {
          uint32_t a_lo = a_tap;
          for (int ks = 0; ks < ksteps; ++ks) {
#pragma unroll
            for (int m = 0; m < 2; ++m)
              if (m < MT)
                tc_mma_bf16_lohi(acc1 + (uint32_t)(m * N), a_lo + (uint32_t)(m * kTileM), a1_hi, b_lo, b_hi, idesc, accumulate);
            accumulate = 1;
            a_lo += 2u * (uint32_t)p.rows_x;
            b_lo += b_kstep;
          }
        }
        tc_commit(acc1_full);
      }
      // conv2: A2 rows o + t (dilation 1)
      mbar_wait(a2_full, ph, 24);
      tc_fence_after();
      {
        uint32_t b_lo = b_lo_fixed + (w2 >> 4);
        uint32_t a_tap = a2_lo_fixed + (a2 >> 4);
        uint32_t accumulate = 0;
        for (int t = 0; t < taps; ++t, a_tap += 1u) {
          uint32_t a_lo = a_tap;
          for (int ks = 0; ks < ksteps; ++ks) {
#pragma unroll
            for (int m = 0; m < 2; ++m)
              if (m < MT)
                tc_mma_bf16_lohi(acc2 + (uint32_t)(m * N), a_lo + (uint32_t)(m * kTileM), a2_hi, b_lo, b_hi, idesc, accumulate);
            accumulate = 1;
            a_lo += 2u * (uint32_t)p.rows_a2;
            b_lo += b_kstep;
          }
        }
        tc_commit(acc2_full);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue / transform warps
    const int q = warp & 3;
    const int et = (warp - 2) * 32 + lane;              // 0..127
    const int n_chunks = N / 32;
    uint32_t it = 0;
    for (int super = blockIdx.x; super < p.n_super; super += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      const int s0 = super * p.Lout - p.h2;
      // ---- A1 = lrelu(XR)
      mbar_wait(xr_full, ph, 25);
      {
        const __nv_bfloat162 slope2 = __float2bfloat162_rn(c.in_slope);
        const int n_chunks16 = p.planes * p.rows_x;
#pragma unroll 2
        for (int i = et; i < n_chunks16; i += 32 * kEpiWarps) {
          uint32_t w[4];
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(xr + (uint32_t)i * 16u));
#pragma unroll
          for (int e = 0; e < 4; ++e) {           // leaky_relu(x) = max(x, slope*x) for 0 < slope < 1, two bf16 per op
            __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&w[e]);
            v = __hmax2(v, __hmul2(v, slope2));
            w[e] = *reinterpret_cast<uint32_t*>(&v);
          }
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a1 + (uint32_t)i * 16u), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3])
                       : "memory");
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(a1_full);

      // ---- epilogue 1: conv1 accumulators -> A2
      mbar_wait(acc1_full, ph, 26);
      tc_fence_after();
      for (int m = 0; m < p.MT; ++m) {
        const int j = m * kTileM + q * 32 + lane;         // A2 local row = conv1 output row
        const int g = s0 + j;
        bool valid = g >= 0 && g < c.R;
        if (valid && c.row_utt) valid = c.row_utt[g >> p.row_div_shift] >= 0;
        const uint32_t t_row = acc1 + ((uint32_t)(q * 32) << 16) + (uint32_t)(m * N);
        for (int cc = 0; cc < n_chunks; ++cc) {
          uint32_t v[32];
          tmem_ld32(t_row + (uint32_t)(cc * 32), v);
#pragma unroll
          for (int gq = 0; gq < 4; ++gq) {
            const int co0 = cc * 32 + gq * 8;
            uint32_t o0 = 0, o1 = 0, o2 = 0, o3 = 0;
            if (valid) {
              float y[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) y[e] = lrelu(__uint_as_float(v[8 * gq + e]) + bias_s[co0 + e], c.in_slope);
              o0 = pack_bf16x2(y[0], y[1]); o1 = pack_bf16x2(y[2], y[3]); o2 = pack_bf16x2(y[4], y[5]); o3 = pack_bf16x2(y[6], y[7]);
            }
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a2 + (uint32_t)((co0 >> 3) * p.rows_a2 + j) * 16u),
                         "r"(o0), "r"(o1), "r"(o2), "r"(o3)
                         : "memory");
          }
        }
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a2_full);

      // ---- epilogue 2: conv2 accumulators + residual -> HBM
      mbar_wait(acc2_full, ph, 27);
      tc_fence_after();
      for (int m = 0; m < p.MT; ++m) {
        const int o = m * kTileM + q * 32 + lane;         // conv2 output position within the super tile
        const int g = s0 + p.h2 + o;
        const bool in_tile = o < p.Lout && g < c.R;       // g >= 0 always (s0 + h2 = super * Lout)
        int utt = -1;
        if (in_tile) utt = c.row_utt ? c.row_utt[g >> p.row_div_shift] : 0;
        const bool valid = utt >= 0;
        const uint32_t t_row = acc2 + ((uint32_t)(q * 32) << 16) + (uint32_t)(m * N);
        const uint32_t xr_row = xr + (uint32_t)(p.h1 + p.h2 + o) * 16u;
        for (int cc = 0; cc < n_chunks; ++cc) {
          uint4 rv2[4];
          if (valid && c.res2) {
#pragma unroll
            for (int gq = 0; gq < 4; ++gq)
              rv2[gq] = *reinterpret_cast<const uint4*>(c.res2 + ((size_t)(cc * 4 + gq) * c.R + g) * 8);
          }
          uint32_t v[32];
          tmem_ld32(t_row + (uint32_t)(cc * 32), v);
          if (in_tile) {
#pragma unroll
            for (int gq = 0; gq < 4; ++gq) {
              const int co0 = cc * 32 + gq * 8;
              const size_t go = ((size_t)(co0 >> 3) * c.R + g) * 8;
              uint4 raw = make_uint4(0, 0, 0, 0), act = make_uint4(0, 0, 0, 0);
              if (valid) {
                uint32_t x0, x1, x2, x3;
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3)
                             : "r"(xr_row + (uint32_t)((co0 >> 3) * p.rows_x) * 16u));
                float xf[8], y[8];
                unpack_bf16x8(make_uint4(x0, x1, x2, x3), xf);
#pragma unroll
                for (int e = 0; e < 8; ++e) y[e] = __uint_as_float(v[8 * gq + e]) + bias_s[N + co0 + e] + xf[e];
                if (c.res2) {
                  float f[8];
                  unpack_bf16x8(rv2[gq], f);
#pragma unroll
                  for (int e = 0; e < 8; ++e) y[e] += f[e];
                }
                raw = make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
                if (c.out_act) {
                  float z[8];
#pragma unroll
                  for (int e = 0; e < 8; ++e) z[e] = lrelu(y[e] * c.act_scale, c.act_slope);
                  act = make_uint4(pack_bf16x2(z[0], z[1]), pack_bf16x2(z[2], z[3]), pack_bf16x2(z[4], z[5]), pack_bf16x2(z[6], z[7]));
                }
              }
              if (c.out_raw) *reinterpret_cast<uint4*>(c.out_raw + go) = raw;
              if (c.out_act) *reinterpret_cast<uint4*>(c.out_act + go) = act;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(xr_empty);               // XR (and with it A1/A2/TMEM) may be reused
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

int make_plan(const UmmaPair& c, Plan* out) {
  Plan p{};
  VS_REQUIRE(c.C == 32 || c.C == 64, "umma_respair: C=%d (only 32 and 64 are fused)", c.C);
  VS_REQUIRE(c.taps % 2 == 1 && c.taps >= 3 && c.dil >= 1 && c.R > 0, "umma_respair: bad shape");
  p.h1 = c.dil * (c.taps - 1) / 2;
  p.h2 = (c.taps - 1) / 2;
  p.planes = c.C / 8;
  p.w_bytes = (uint32_t)c.taps * c.C * c.C * 2u;
  int s = 0;
  while ((1 << s) < c.row_div) ++s;
  VS_REQUIRE((1 << s) == c.row_div, "umma_respair: row_div=%d must be a power of two", c.row_div);
  p.row_div_shift = s;
  const uint32_t fixed = 64 + 16 + 128 + 2u * c.C * 4u + 256;
  const uint32_t half_sm = 110u * 1024, full_sm = 222u * 1024;
  p.MT = 0;
  for (int pass = 0; pass < 2 && !p.MT; ++pass)
    for (int mt = 2; mt >= 1 && !p.MT; --mt) {
      const uint32_t L = 128u * mt;
      const uint32_t need = 2u * p.planes * (L + 2 * p.h1) * 16u + p.planes * (L + 2 * p.h2) * 16u + 2 * p.w_bytes + fixed;
      if (need <= (pass == 0 ? half_sm : full_sm)) p.MT = mt;
    }
  VS_REQUIRE(p.MT > 0, "umma_respair: C=%d k=%d d=%d does not fit in shared memory", c.C, c.taps, c.dil);
  p.L = 128 * p.MT;
  p.Lout = p.L - 2 * p.h2;
  p.rows_x = p.L + 2 * p.h1;
  p.rows_a2 = p.L + 2 * p.h2;
  p.xr_bytes = (uint32_t)p.planes * p.rows_x * 16u;
  p.a2_bytes = (uint32_t)p.planes * p.rows_a2 * 16u;
  p.off_a1 = p.xr_bytes;
  p.off_a2 = p.off_a1 + p.xr_bytes;
  p.off_w1 = p.off_a2 + p.a2_bytes;
  p.off_w2 = p.off_w1 + p.w_bytes;
  p.off_bar = (p.off_w2 + p.w_bytes + 127u) & ~127u;
  p.off_bias = p.off_bar + 128u;
  p.smem_bytes = p.off_bias + 2u * c.C * 4u;
  // conv1 and conv2 reuse the same TMEM columns (their phases never overlap).  (Splitting each row tile over several
  // accumulators by K-step to shorten the dependent MMA chains was measured slower: the 4 epilogue warps pay more
  // for the extra tcgen05.ld + adds than the MMAs gain.)
  int cols = 32;
  while (cols < p.MT * c.C) cols *= 2;
  p.tmem_cols = cols;
  int per_sm = (int)((227u * 1024) / (p.smem_bytes + 1024));
  if (per_sm > 512 / p.tmem_cols) per_sm = 512 / p.tmem_cols;
  if (per_sm > 3) per_sm = 3;
  if (per_sm < 1) per_sm = 1;
  const uint32_t min_smem = (227u * 1024) / (uint32_t)(per_sm + 1) + 1024u;
  if (p.smem_bytes < min_smem) p.smem_bytes = min_smem;
  p.ctas_per_sm = per_sm;
  p.n_super = (c.R + p.Lout - 1) / p.Lout;
  *out = p;
  return VS_OK;
}

}  // namespace

static bool g_enabled = false;
void umma_respair_enable(bool on) { g_enabled = on; }

bool umma_respair_supported(int C, int taps, int dil) {
  UmmaPair c;
  c.C = C; c.taps = taps; c.dil = dil; c.R = 1024; c.row_div = 1;
  Plan p;
  // Experimental: after the epilogue specialisation of umma_conv.cu the unfused pair runs at ~90 % of HBM peak
  // (C=32,k=3: 0.31 + 0.65 ms) and this sequential-phase kernel (0.90 ms) no longer wins; its phases are bound by
  // the dependent-MMA latency per super tile.  Kept (with its parity test) as the base for a software-pipelined
  // version; off unless vs_set_option("fused_respair", 1).
  if (!g_enabled || (C != 32 && C != 64)) return false;
  return make_plan(c, &p) == VS_OK;
}

int umma_respair(const UmmaPair& c, cudaStream_t st) {
  Params prm;
  prm.c = c;
  VS_REQUIRE(c.x && c.w1 && c.w2 && c.b1 && c.b2 && (c.out_raw || c.out_act), "umma_respair: null pointer");
  VS_TRY(make_plan(c, &prm.p));
  static int n_sm = 0;
  static bool configured = false;
  if (!configured) {
    int dev = 0;
    VS_CUDA_CHECK(cudaGetDevice(&dev));
    VS_CUDA_CHECK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    VS_CUDA_CHECK(cudaFuncSetAttribute(umma_respair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  int grid = n_sm * prm.p.ctas_per_sm;
  if (grid > prm.p.n_super) grid = prm.p.n_super;
  umma_respair_kernel<<<grid, kThreads, prm.p.smem_bytes, st>>>(prm);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

}  // namespace vs
