// One ResBlock1 iteration (reference modules.py:211-220) as ONE tcgen05 kernel for the narrow decoder stages:
//
//     y = c2( lrelu( c1( lrelu(x) ) + b1 ) ) + b2 + x        (c1: k taps, dilation d;  c2: k taps, dilation 1)
//
// The unfused chain (umma_conv.cu) moves 6 activation-sized tensors through HBM per iteration (read lrelu(x), write
// lrelu(c1), read it back, read x, write y and lrelu(y)) and stages 2-3 (C = 64, 32) sit on the HBM roofline of that
// traffic.  Here ONE tensor is read and ONE written per iteration:
//   * only the ACTIVATED stream a = lrelu(x) travels between iterations.  It is the conv's A operand as is (TMA bulk copy
//     straight into the UMMA layout), and the residual is recovered in the last epilogue by inverting the leaky-relu,
//     x = min(a, a / slope): in bf16 that costs the same relative rounding as storing x itself;
//   * the intermediate lrelu(c1 + b1) is written by the first epilogue directly into shared memory in the layout the
//     second conv reads (rows outside the sequence / in gaps are zeroed there: they are c2's zero padding).
// Per CTA and super tile of L = 128*MT conv1 rows (outputs valid on L - (k-1) rows):
//   producer  : XA[s] <- rows [s0-h1, s0+L+h1) of a (2-deep ring, zero fill outside [0,R))
//   MMA       : conv1 (taps = descriptor offsets t*d rows into XA)   -> TMEM      | each row tile accumulates into S
//   epilogue  : TMEM -> +b1 -> lrelu -> mask -> bf16 -> A2 (smem)                 | accumulators round-robin by K-step:
//   MMA       : conv2 over A2 (offsets t)                             -> TMEM      | dependent MMA chains on one
//   epilogue  : TMEM -> +b2 + x(from XA) [+ MRF sum] -> lrelu -> HBM               | accumulator cost ~250 cycles a link
// Both convs' weights stay resident in shared memory; 2-3 CTAs share an SM and overlap each other's phases.
#include "umma_conv.cuh"
#include "umma_common.cuh"

namespace vs {
namespace {

using namespace umma;

constexpr int kThreads = 192;     // warp 0 producer, warp 1 MMA, warps 2..5 epilogue (one per TMEM lane quarter)
constexpr int kEpiWarps = 4;
constexpr int kXA = 2;            // input ring depth

struct Plan {
  int MT, L, Lout, h1, h2, planes, rows_x, rows_a2, n_super, tmem_cols, ctas_per_sm, row_div_shift, S;
  uint32_t xa_bytes, a2_bytes, w_bytes, smem_bytes;
  uint32_t off_a2, off_w1, off_w2, off_bar, off_bias;
};
struct Params {
  UmmaPair c;
  Plan p;
};

template <int N, int S>
__global__ void __launch_bounds__(kThreads, 3) umma_respair_kernel(const __grid_constant__ Params prm) {
  extern __shared__ __align__(128) uint8_t smem[];
  const UmmaPair& c = prm.c;
  const Plan& p = prm.p;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;

  const uint32_t smem_base = smem_u32(smem);
  const uint32_t xa = smem_base, a2 = smem_base + p.off_a2;
  const uint32_t w1 = smem_base + p.off_w1, w2 = smem_base + p.off_w2, bar = smem_base + p.off_bar;
  const uint32_t w_full = bar, acc1_full = bar + 8, a2_full = bar + 16, acc2_full = bar + 24, tmem_free = bar + 32;
  auto xa_full = [&](int i) { return bar + 40u + 8u * i; };
  auto xa_empty = [&](int i) { return bar + 56u + 8u * i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + p.off_bar + 80);
  float* bias_s = reinterpret_cast<float*>(smem + p.off_bias);      // [2][N]

  if (threadIdx.x == 0) {
    mbar_init(w_full, 1); mbar_init(acc1_full, 1); mbar_init(a2_full, kEpiWarps); mbar_init(acc2_full, 1);
    mbar_init(tmem_free, kEpiWarps);
    for (int i = 0; i < kXA; ++i) { mbar_init(xa_full(i), 1); mbar_init(xa_empty(i), kEpiWarps); }
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
  const uint32_t tmem_base = *tmem_slot;     // accumulator (m, split) lives at column (m*S + split) * N; conv1 and conv2
                                             // reuse the same columns (conv2 starts after epilogue 1 drained them)
  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    if (lane == 0) {
      mbar_arrive_expect_tx(w_full, 2 * p.w_bytes);
      bulk_g2s(w1, c.w1, p.w_bytes, w_full);
      bulk_g2s(w2, c.w2, p.w_bytes, w_full);
    }
    uint32_t slot = 0, phase = 0;
    for (int super = blockIdx.x; super < p.n_super; super += gridDim.x) {
      mbar_wait(xa_empty(slot), phase ^ 1, 21);
      const uint32_t stage = xa + slot * p.xa_bytes;
      const int s0 = super * p.Lout - p.h2;
      const int row_lo = s0 - p.h1, row_hi = row_lo + p.rows_x;
      const int c_lo = row_lo < 0 ? 0 : row_lo, c_hi = row_hi > c.R ? c.R : row_hi;
      const int n_zero_lo = c_lo - row_lo, n_zero_hi = row_hi - c_hi;
      if (n_zero_lo > 0 || n_zero_hi > 0) {
        const int per_plane = n_zero_lo + n_zero_hi;
        for (int i = lane; i < p.planes * per_plane; i += 32) {
          const int pl = i / per_plane, j = i % per_plane;
          const int row = j < n_zero_lo ? j : (p.rows_x - n_zero_hi + (j - n_zero_lo));
          asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(stage + (uint32_t)(pl * p.rows_x + row) * 16u), "r"(0)
                       : "memory");
        }
        fence_proxy_async();
      }
      __syncwarp();
      const uint32_t bytes = (uint32_t)(c_hi - c_lo) * 16u;
      if (lane == 0) mbar_arrive_expect_tx(xa_full(slot), bytes * p.planes);
      __syncwarp();
      for (int pl = lane; pl < p.planes; pl += 32)
        bulk_g2s(stage + (uint32_t)(pl * p.rows_x + n_zero_lo) * 16u, c.x + ((size_t)pl * c.R + c_lo) * 8, bytes, xa_full(slot));
      if (++slot == kXA) { slot = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (warp-uniform, elected lane issues)
    const uint32_t idesc = make_idesc(N);
    const uint32_t b_lbo = (uint32_t)N * 16u;
    const uint32_t b_hi = (uint32_t)(make_desc(0, b_lbo, 128u) >> 32), b_lo_fixed = (uint32_t)make_desc(0, b_lbo, 128u);
    const uint32_t a1_lbo = (uint32_t)p.rows_x * 16u, a2_lbo = (uint32_t)p.rows_a2 * 16u;
    const uint32_t a1_hi = (uint32_t)(make_desc(0, a1_lbo, 128u) >> 32), a1_lo_fixed = (uint32_t)make_desc(0, a1_lbo, 128u);
    const uint32_t a2_hi = (uint32_t)(make_desc(0, a2_lbo, 128u) >> 32), a2_lo_fixed = (uint32_t)make_desc(0, a2_lbo, 128u);
    constexpr uint32_t b_kstep = 2u * N;
    constexpr int ksteps = N / 16;
    const int taps = c.taps, MT = p.MT;
    mbar_wait(w_full, 0, 22);
    tc_fence_after();
    uint32_t slot = 0, phase = 0, it = 0;
    for (int super = blockIdx.x; super < p.n_super; super += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      mbar_wait(tmem_free, ph ^ 1, 28);           // epilogue 2 of the previous super tile has drained the accumulators
      mbar_wait(xa_full(slot), phase, 23);
      tc_fence_after();
      {  // conv1: XA rows j + t*d
        uint32_t b_lo = b_lo_fixed + (w1 >> 4);
        uint32_t a_tap = a1_lo_fixed + ((xa + slot * p.xa_bytes) >> 4);
        int step = 0;
        for (int t = 0; t < taps; ++t, a_tap += (uint32_t)c.dil) {
          uint32_t a_lo = a_tap;
#pragma unroll
          for (int ks = 0; ks < ksteps; ++ks, ++step) {
            const uint32_t sp = (uint32_t)(step % S);
            const uint32_t accumulate = step >= S ? 1u : 0u;
#pragma unroll
            for (int m = 0; m < 2; ++m)
              if (m < MT)
                tc_mma_bf16_lohi(tmem_base + (uint32_t)(m * S) * N + sp * N, a_lo + (uint32_t)(m * kTileM), a1_hi, b_lo, b_hi,
                                 idesc, accumulate);
            a_lo += 2u * (uint32_t)p.rows_x;
            b_lo += b_kstep;
          }
        }
        tc_commit(acc1_full);
      }
      mbar_wait(a2_full, ph, 24);
      tc_fence_after();
      {  // conv2: A2 rows o + t (dilation 1)
        uint32_t b_lo = b_lo_fixed + (w2 >> 4);
        uint32_t a_tap = a2_lo_fixed + (a2 >> 4);
        int step = 0;
        for (int t = 0; t < taps; ++t, a_tap += 1u) {
          uint32_t a_lo = a_tap;
#pragma unroll
          for (int ks = 0; ks < ksteps; ++ks, ++step) {
            const uint32_t sp = (uint32_t)(step % S);
            const uint32_t accumulate = step >= S ? 1u : 0u;
#pragma unroll
            for (int m = 0; m < 2; ++m)
              if (m < MT)
                tc_mma_bf16_lohi(tmem_base + (uint32_t)(m * S) * N + sp * N, a_lo + (uint32_t)(m * kTileM), a2_hi, b_lo, b_hi,
                                 idesc, accumulate);
            a_lo += 2u * (uint32_t)p.rows_a2;
            b_lo += b_kstep;
          }
        }
        tc_commit(acc2_full);
      }
      if (++slot == kXA) { slot = 0; phase ^= 1; }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int q = warp & 3;
    constexpr int n_chunks = N / 32;
    const float slope = c.in_slope, inv_slope = 1.f / c.in_slope;
    const float oslope = c.act_slope, oscale = c.act_scale;
    const bool has_res2 = c.res2 != nullptr, has_raw = c.out_raw != nullptr, has_act = c.out_act != nullptr;
    uint32_t slot = 0, it = 0;
    // sum of the S split accumulators of row tile m, 32 columns from column cc*32
    auto load_acc = [&](int m, int cc, uint32_t (&v)[32]) {
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(m * S) * N + (uint32_t)(cc * 32);
      tmem_ld32(t_row, v);
#pragma unroll
      for (int sp = 1; sp < S; ++sp) {
        uint32_t u[32];
        tmem_ld32(t_row + (uint32_t)sp * N, u);
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) + __uint_as_float(u[e]));
      }
    };
    for (int super = blockIdx.x; super < p.n_super; super += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      const int s0 = super * p.Lout - p.h2;
      const uint32_t stage = xa + slot * p.xa_bytes;
      // ---- epilogue 1: conv1 accumulators -> A2 = lrelu(c1 + b1), zero where c2 must see padding
      mbar_wait(acc1_full, ph, 26);
      tc_fence_after();
      for (int m = 0; m < p.MT; ++m) {
        const int j = m * kTileM + q * 32 + lane;         // A2 local row = conv1 output row
        const int g = s0 + j;
        bool valid = g >= 0 && g < c.R;
        if (valid && c.row_utt) valid = c.row_utt[g >> p.row_div_shift] >= 0;
#pragma unroll
        for (int cc = 0; cc < n_chunks; ++cc) {
          uint32_t v[32];
          load_acc(m, cc, v);
#pragma unroll
          for (int gq = 0; gq < 4; ++gq) {
            const int co0 = cc * 32 + gq * 8;
            uint32_t o0 = 0, o1 = 0, o2 = 0, o3 = 0;
            if (valid) {
              float y[8];
              const float4 b0 = *reinterpret_cast<const float4*>(bias_s + co0);
              const float4 b1 = *reinterpret_cast<const float4*>(bias_s + co0 + 4);
              const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float t = __uint_as_float(v[8 * gq + e]) + bb[e];
                y[e] = fmaxf(t, t * slope);
              }
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

      // ---- epilogue 2: conv2 accumulators + b2 + x (inverse lrelu of the staged input) [+ MRF sum] -> HBM
      mbar_wait(acc2_full, ph, 27);
      tc_fence_after();
      for (int m = 0; m < p.MT; ++m) {
        const int o = m * kTileM + q * 32 + lane;         // conv2 output position within the super tile
        const int g = s0 + p.h2 + o;
        const bool in_tile = o < p.Lout && g < c.R;       // g >= 0 always (s0 + h2 = super * Lout)
        int utt = -1;
        if (in_tile) utt = c.row_utt ? c.row_utt[g >> p.row_div_shift] : 0;
        const bool valid = utt >= 0;
        const uint32_t xa_row = stage + (uint32_t)(p.h1 + p.h2 + o) * 16u;
#pragma unroll
        for (int cc = 0; cc < n_chunks; ++cc) {
          uint4 rv2[4];
          if (valid && has_res2) {
#pragma unroll
            for (int gq = 0; gq < 4; ++gq)
              rv2[gq] = *reinterpret_cast<const uint4*>(c.res2 + ((size_t)(cc * 4 + gq) * c.R + g) * 8);
          }
          uint32_t v[32];
          load_acc(m, cc, v);
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
                             : "r"(xa_row + (uint32_t)((co0 >> 3) * p.rows_x) * 16u));
                float xf[8], y[8];
                unpack_bf16x8(make_uint4(x0, x1, x2, x3), xf);
                const float4 b0 = *reinterpret_cast<const float4*>(bias_s + N + co0);
                const float4 b1 = *reinterpret_cast<const float4*>(bias_s + N + co0 + 4);
                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int e = 0; e < 8; ++e)     // x = lrelu^-1(a) = min(a, a / slope)
                  y[e] = __uint_as_float(v[8 * gq + e]) + bb[e] + fminf(xf[e], xf[e] * inv_slope);
                if (has_res2) {
                  float f[8];
                  unpack_bf16x8(rv2[gq], f);
#pragma unroll
                  for (int e = 0; e < 8; ++e) y[e] += f[e];
                }
                if (has_raw)
                  raw = make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
                if (has_act) {
                  float z[8];
#pragma unroll
                  for (int e = 0; e < 8; ++e) {
                    const float t = y[e] * oscale;
                    z[e] = fmaxf(t, t * oslope);
                  }
                  act = make_uint4(pack_bf16x2(z[0], z[1]), pack_bf16x2(z[2], z[3]), pack_bf16x2(z[4], z[5]), pack_bf16x2(z[6], z[7]));
                }
              }
              if (has_raw) *reinterpret_cast<uint4*>(c.out_raw + go) = raw;
              if (has_act) *reinterpret_cast<uint4*>(c.out_act + go) = act;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(xa_empty(slot)); mbar_arrive(tmem_free); }
      if (++slot == kXA) slot = 0;
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
  VS_REQUIRE(c.in_slope > 0.f && c.in_slope < 1.f && c.act_slope > 0.f && c.act_slope <= 1.f, "umma_respair: bad slopes");
  p.h1 = c.dil * (c.taps - 1) / 2;
  p.h2 = (c.taps - 1) / 2;
  p.planes = c.C / 8;
  p.w_bytes = (uint32_t)c.taps * c.C * c.C * 2u;
  int s = 0;
  while ((1 << s) < c.row_div) ++s;
  VS_REQUIRE((1 << s) == c.row_div, "umma_respair: row_div=%d must be a power of two", c.row_div);
  p.row_div_shift = s;
  p.S = (c.taps * (c.C / 16) >= 12) ? 2 : 1;       // split long dependent MMA chains over two accumulators
  const uint32_t fixed = 128 + 2u * c.C * 4u + 256;
  const uint32_t caps[3] = {75u * 1024, 110u * 1024, 222u * 1024};     // 3, 2, 1 CTAs per SM
  p.MT = 0;
  for (int pass = 0; pass < 3 && !p.MT; ++pass)
    for (int mt = 2; mt >= 1 && !p.MT; --mt) {
      const uint32_t L = 128u * mt;
      const uint32_t need = kXA * p.planes * (L + 2 * p.h1) * 16u + p.planes * (L + 2 * p.h2) * 16u + 2 * p.w_bytes + fixed;
      if (need <= caps[pass] && mt * p.S * c.C <= 512 / (3 - pass)) p.MT = mt;
    }
  VS_REQUIRE(p.MT > 0, "umma_respair: C=%d k=%d d=%d does not fit in shared memory", c.C, c.taps, c.dil);
  p.L = 128 * p.MT;
  p.Lout = p.L - 2 * p.h2;
  p.rows_x = p.L + 2 * p.h1;
  p.rows_a2 = p.L + 2 * p.h2;
  p.xa_bytes = (uint32_t)p.planes * p.rows_x * 16u;
  p.a2_bytes = (uint32_t)p.planes * p.rows_a2 * 16u;
  p.off_a2 = kXA * p.xa_bytes;
  p.off_w1 = p.off_a2 + p.a2_bytes;
  p.off_w2 = p.off_w1 + p.w_bytes;
  p.off_bar = (p.off_w2 + p.w_bytes + 127u) & ~127u;
  p.off_bias = p.off_bar + 128u;
  p.smem_bytes = p.off_bias + 2u * c.C * 4u;
  int cols = 32;
  while (cols < p.MT * p.S * c.C) cols *= 2;
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

int g_mode = 1;      // 0 off, 1 where faster (default), 2 everywhere it fits

}  // namespace

void umma_respair_enable(int mode) { g_mode = mode; }

// Which ResBlock iterations run fused in the decoder.  Measured on B200 at the C2 size (profiles/decoder_convs_r1.txt):
// in isolation the fused kernel wins only for C = 32, k = 3 (0.56 ms vs 0.31 + 0.65 ms); for k = 7 / 11 (1.05 / 1.2-1.4
// ms vs 0.97 / 1.16 ms) and C = 64 (0.97 vs 0.87 ms) its per-CTA phase sequence (conv1 -> epilogue -> conv2 -> epilogue
// on 4 epilogue warps, 2-3 CTAs per SM) is latency-bound.  Inside the full decoder an A/B on one box showed no net gain
// (33.6 vs 33.8 ms), so the default is OFF; vs_set_option("fused_respair", 1 | 2) turns it on (parity-tested).
bool umma_respair_supported(int C, int taps, int dil) {
  if (g_mode == 0) return false;
  if (g_mode == 1 && !(C == 32 && taps == 3)) return false;
  if (!(C == 32 || (C == 64 && taps == 3))) return false;
  UmmaPair c;
  c.C = C; c.taps = taps; c.dil = dil; c.R = 1024; c.row_div = 1; c.in_slope = 0.1f;
  Plan p;
  return make_plan(c, &p) == VS_OK;
}

int umma_respair(const UmmaPair& c, cudaStream_t st) {
  Params prm;
  prm.c = c;
  VS_REQUIRE(c.x && c.w1 && c.w2 && c.b1 && c.b2 && (c.out_raw || c.out_act), "umma_respair: null pointer");
  VS_TRY(make_plan(c, &prm.p));
  static int n_sm = 0;
  if (!n_sm) {
    int dev = 0;
    VS_CUDA_CHECK(cudaGetDevice(&dev));
    VS_CUDA_CHECK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  }
  int grid = n_sm * prm.p.ctas_per_sm;
  if (grid > prm.p.n_super) grid = prm.p.n_super;
#define VS_PAIR_CASE(NN, SS)                                                                                          \
  if (c.C == NN && prm.p.S == SS) {                                                                                   \
    static bool cfg = false;                                                                                          \
    if (!cfg) {                                                                                                       \
      VS_CUDA_CHECK(cudaFuncSetAttribute(umma_respair_kernel<NN, SS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
      cfg = true;                                                                                                     \
    }                                                                                                                 \
    umma_respair_kernel<NN, SS><<<grid, kThreads, prm.p.smem_bytes, st>>>(prm);                                       \
  }
  VS_PAIR_CASE(32, 1) else VS_PAIR_CASE(32, 2) else VS_PAIR_CASE(64, 1) else VS_PAIR_CASE(64, 2) else {
    set_error("umma_respair: no instantiation for C=%d S=%d", c.C, prm.p.S);
    return VS_ERR_INVALID;
  }
#undef VS_PAIR_CASE
  VS_LAUNCH_CHECK();
  return VS_OK;
}

}  // namespace vs
