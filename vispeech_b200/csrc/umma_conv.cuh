// tcgen05 / TMEM implicit-GEMM conv1d over planar f16 ragged rows (sm_100a).  See umma_conv.cu.
#pragma once
#include "common.cuh"

namespace vs {

// Activations: planar f16 [C/8][R][8]  (plane p holds channels 8p..8p+7 of every row, 16 B per row).
// Weights:     f16 [NB][taps][Cin/KC][KC/8][Nblk][8], KC = min(Cin, 64), Nblk = min(N, 256), N = NB*Nblk:
//              element (nb, t, kc, p, n, e) = W[t][kc*KC + 8p + e][nb*Nblk + n].  One (nb,t,kc) slab is one
//              cp.async.bulk of KC*Nblk*2 bytes and lands in smem already in the UMMA K-major no-swizzle layout.
struct UmmaConv {
  const __half* in = nullptr;   // planar [Cin/8][R][8]
  const __half* w = nullptr;    // packed as above
  const float* bias = nullptr;         // [N % upsample-aware: bias[gn % Cout]] or null
  const float* ubias = nullptr;        // optional per-speaker table [n_spk][N]
  const int32_t* ubias_idx = nullptr;  // [n_utt] -> row of ubias
  const __half* res = nullptr;  // optional residual, planar like out (same rows/channels)
  float res_inv_slope = 0.f;           // != 0: `res` holds a = lrelu(x, 1/res_inv_slope); the residual added is
                                       // x = min(a, a * res_inv_slope) (exact inverse up to f16 rounding), so only the
                                       // activated stream has to be stored between ResBlock iterations
  const __half* res2 = nullptr; // optional second residual (MRF running sum)
  __half* out_raw = nullptr;    // optional: y
  __half* out_lo = nullptr;     // optional (needs out_raw): fp16(y - fp16(y)), so that out_raw + out_lo carries y to 22 bits
  __half* out_act = nullptr;    // optional: lrelu(y * act_scale, act_slope)
  const int32_t* row_utt = nullptr;    // validity of OUTPUT row: row_utt[orow / row_div] >= 0 (also gives utt for ubias)
  int row_div = 1;
  int R = 0;                           // input rows
  int Cin = 0, N = 0;                  // N = total GEMM columns (= Cout * up for transposed conv)
  int taps = 1, dil = 1, pad_l = 0;
  int up = 1;                          // ConvTranspose1d polyphase: column gn -> phase gn / Cout, channel gn % Cout;
                                       // output row = up*r + phase; output planes have R*up rows
  float act_slope = 1.f, act_scale = 1.f;
};

int umma_conv1d(const UmmaConv& a, cudaStream_t st);
// The same conv on a CTA pair (tcgen05 cta_group::2, weights resident and split between the two CTAs; umma_pair.cu): Cin = N = 128
// (any k of the decoder) and 256 (k = 3).  umma_conv1d routes to it when option "pair_conv" is on.
bool umma_pair_supported(const UmmaConv& c);
int umma_pair_conv(const UmmaConv& c, cudaStream_t st);
void* umma_conv_timing_buffer();   // option "umma_timing_buffer": >= 296*12 int64 of per-CTA wait clocks, or null

// One fused ResBlock1 iteration y = c2(lrelu(c1(a))) + lrelu^-1(a), a = lrelu(x) (umma_respair.cu), C in {32, 64}.
struct UmmaPair {
  const __half* x = nullptr;      // ACTIVATED input a = lrelu(x, in_slope), planar [C/8][R][8]
  const __half* w1 = nullptr;     // c1 weights, slabs as in UmmaConv (k taps, dilation dil)
  const __half* w2 = nullptr;     // c2 weights (k taps, dilation 1)
  const float* b1 = nullptr;             // device
  const float* b2 = nullptr;
  const float* b1_host = nullptr;        // optional host copies of b1 / b2 (they travel in the kernel's parameter block);
  const float* b2_host = nullptr;        // without them the call fetches the device arrays and synchronises the stream
  const __half* res2 = nullptr;   // optional running MRF sum added to y
  __half* out_raw = nullptr;      // y
  __half* out_act = nullptr;      // lrelu(y * act_scale, act_slope)
  const int32_t* row_utt = nullptr; int row_div = 1;
  int R = 0, C = 0, taps = 3, dil = 1;
  float in_slope = 0.1f;                 // LRELU_SLOPE of modules.py:17 (both inner leaky-relus)
  float act_slope = 1.f, act_scale = 1.f;
};
// The whole last MRF stage (C = 32): 3 ResBlock1 + sum/3 + lrelu(0.01) + conv_post + tanh in one kernel (umma_mrf.cu).
struct UmmaMrf {
  const __half* x_hi = nullptr;          // ups[3] output x0 = hi + lo, planar [4][R][8] each (NOT activated)
  const __half* x_lo = nullptr;
  const __half* w[3][3][2] = {};         // [resblock k=3,7,11][iteration d=1,3,5][c1 | c2], slabs as in UmmaConv
  const float* b1_host[3][3] = {};       // host copies of the biases (32 floats each): they travel in the parameter block
  const float* b2_host[3][3] = {};
  const float* post_w_host = nullptr;    // conv_post weights [7][32] (host)
  float* wave = nullptr;                 // [R] tanh(conv_post(...)), zeros on gap rows
  const int32_t* row_utt = nullptr; int row_div = 1;
  int R = 0;
};
int umma_mrf(const UmmaMrf& c, cudaStream_t st);

// fp32-accurate conv over planar fp16 hi / lo operands (umma_split.cu): D += a_hi w_hi + a_lo w_hi + a_hi w_lo.
struct UmmaSplit {
  const __half* in_hi = nullptr;    // planar [Cin/8][R][8]
  const __half* in_lo = nullptr;
  const __half* w = nullptr;        // packing.py pack_split16: [K-slice][NB][tap][cs / KC][hi|lo][KC/8][Nblk][8]
  const float* bias = nullptr;      // [N] or null (added by K-slice 0)
  float* out32 = nullptr;           // optional fp32 row-major [R][out32_ld]; K-slice s writes out32 + s * out32_slice
  int out32_ld = 0;
  int64_t out32_slice = 0;
  __half* out_hi = nullptr;         // optional planar [N/8][R][8] hi / lo of the result (k_slices == 1 only)
  __half* out_lo = nullptr;
  const int32_t* row_utt = nullptr; // validity: zeros are written on rows with row_utt < 0
  int R = 0, Cin = 0, N = 0, taps = 1, dil = 1, pad_l = 0;
  int act = 0;                      // 0 none, 1 ReLU
  int k_slices = 1;                 // Cin / k_slices channels (<= 192) per slice
};
int umma_split(const UmmaSplit& c, cudaStream_t st);
// fp32 rows -> planar hi / lo; n_sum > 1 sums that many partials (x + s * x_stride) first; zeros on invalid rows
int rows_to_split(const float* x, int ld, int64_t x_stride, int n_sum, const int32_t* row_utt, __half* hi, __half* lo, int R, int C,
                  cudaStream_t st);

int sum_partials(const float* x, int64_t stride, int n, float* out, int64_t total, cudaStream_t st);   // out = sum_s x[s * stride + .]

// A whole ResBlock1 of the C = 64 stage's k = 3 branch in one kernel, residual stream in fp32 in TMEM (umma_resblock.cu).
struct UmmaResBlock {
  const __half* a = nullptr;             // ACTIVATED input a = lrelu(x0, 0.1), planar [8][R][8]; x0 is recovered as min(a, 10 a)
  const __half* w[3][2] = {};            // [iteration d = 1, 3, 5][c1 | c2], slabs as in UmmaConv
  const float* b1_host[3] = {};          // host copies of the biases (64 floats each): they travel in the parameter block
  const float* b2_host[3] = {};
  __half* out_raw = nullptr;             // the ResBlock's output, planar [8][R][8], zeros on gap rows
  const int32_t* row_utt = nullptr; int row_div = 1;
  int R = 0;
};
bool umma_resblock_supported(int C, int taps);
int umma_resblock(const UmmaResBlock& c, cudaStream_t st);

bool umma_respair_supported(int C, int taps, int dil);
// the same fused iteration on a CTA pair (tcgen05 cta_group::2; umma_pairfused.cu): C = 64 (every k) and C = 128 (k = 3)
bool umma_pairfused_supported(int C, int taps, int dil);
int umma_pairfused(const UmmaPair& c, cudaStream_t st);
int umma_respair(const UmmaPair& c, cudaStream_t st);

}  // namespace vs
