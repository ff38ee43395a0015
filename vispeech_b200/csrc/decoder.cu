// HiFi-GAN Generator.forward (reference models.py:271-290; ResBlock1.forward modules.py:210-223).
// decode_f32: the same math on the fp32 CUDA-core conv - a test-only cross-check of the f16 tcgen05
// decoder (decoder_umma.cu), selected explicitly with precision=1; never an automatic fallback.
#include "decoder.cuh"
#include "ops_misc.cuh"

namespace vs {

int resolve_decoder(const FetchFn& fetch, int n_speakers, DecoderW* d) {
#define DF32(field, name, numel) VS_TRY(fetch(name, numel, VS_DTYPE_F32, reinterpret_cast<const void**>(&(field))))
#define DF16(field, name, numel) VS_TRY(fetch(name, numel, VS_DTYPE_F16, reinterpret_cast<const void**>(&(field))))
  DF32(d->pre.w, "dec.pre.w", 7 * 192 * 512);
  DF32(d->pre.b, "dec.pre.b", 512);
  DF32(d->cond_tab, "dec.cond_tab", (int64_t)n_speakers * 512);
  DF32(d->post_w, "dec.post.w", 7 * 32);
  for (int i = 0; i < kDecStages; ++i) {
    const int cin = kStageC[i], cout = kStageC[i + 1];
    const std::string p = "dec.ups." + std::to_string(i);
    DF32(d->ups[i].w, p + ".w", (int64_t)kUpKernel[i] * cin * cout);
    DF32(d->ups[i].b, p + ".b", cout);
    for (int j = 0; j < kDecKernels; ++j) {
      const int n = i * kDecKernels + j;
      for (int mth = 0; mth < kDecDils; ++mth) {
        const std::string q = "dec.rb." + std::to_string(n) + ".";
        const int64_t numel = (int64_t)kResK[j] * cout * cout;
        DF32(d->c1[n][mth].w, q + "c1." + std::to_string(mth) + ".w", numel);
        DF32(d->c1[n][mth].b, q + "c1." + std::to_string(mth) + ".b", cout);
        DF32(d->c2[n][mth].w, q + "c2." + std::to_string(mth) + ".w", numel);
        DF32(d->c2[n][mth].b, q + "c2." + std::to_string(mth) + ".b", cout);
      }
    }
  }
  return resolve_decoder_f16(fetch, d);
#undef DF32
#undef DF16
}

// the widest stage buffers are stage 2 (256 rows x 64 ch) and stage 3 (512 x 32): 16384 floats per frame
// fp32 cross-check: 5 buffers x 16384 floats per frame; f16 path: 6 buffers x 16384 f16 (fits in the same bound)
int64_t decoder_ws_floats(int rf) { return (int64_t)rf * (5 * 16384 + 192 + 8) + 4096; }

int decode_f32(const DecoderW& w, const VsRows& rows, const float* z, int max_len, float* wave, Workspace& ws,
               cudaStream_t st) {
  const int R = rows.n_rows;
  int32_t* valid = ws.take<int32_t>(R);
  float* zin = ws.take<float>((int64_t)R * kHidden);
  float* buf[5];
  for (int i = 0; i < 5; ++i) buf[i] = ws.take<float>((int64_t)R * 16384);
  if (!ws.ok) { set_error("decode_f32: workspace too small"); return VS_ERR_WORKSPACE; }
  float *X = buf[0], *T = buf[1], *A = buf[2], *B = buf[3], *S = buf[4];

  VS_TRY(mask_frames(rows, max_len, valid, st));                        // (z * x_mask)[:, :, :max_len]  models.py:720
  VS_TRY(masked_copy(z, valid, zin, R, kHidden, st));

  ConvF32 c;
  c.R = R; c.row_utt = valid; c.in = zin; c.in_ld = kHidden; c.Cin = kHidden; c.w = w.pre.w; c.bias = w.pre.b;
  c.ubias = w.cond_tab; c.ubias_ld = 512; c.ubias_idx = rows.sid; c.out = S; c.out_ld = 512; c.Cout = 512; c.k = 7; c.pad_l = 3;
  VS_TRY(conv1d_f32(c, st));                                             // conv_pre + cond(g)  models.py:272-274

  int mul = 1;                                                           // rows per frame at the current stage
  for (int i = 0; i < kDecStages; ++i) {
    const int cin = kStageC[i], cout = kStageC[i + 1], s = kUpRate[i], K = kUpKernel[i], p = kUpPad[i];
    const int Rin = R * mul, Rout = Rin * s, taps = K / s;
    // ConvTranspose1d as s polyphase convs: out[s*q+ph] = sum_t W[ph+p-s*d] in[q+d], d = dmin+t (models.py:277-278)
    for (int ph = 0; ph < s; ++ph) {
      const int dmax = (ph + p) / s, dmin = dmax - taps + 1;
      c = ConvF32();
      c.R = Rin; c.in = S; c.in_ld = cin; c.Cin = cin; c.in_slope = (i == 0) ? 0.1f : 0.1f;
      c.w = w.ups[i].w + (size_t)ph * taps * cin * cout; c.bias = w.ups[i].b;
      c.out = X; c.out_ld = cout; c.Cout = cout; c.k = taps; c.pad_l = -dmin;
      c.out_row_mul = s; c.out_row_off = ph; c.R_out = Rout; c.row_utt = valid; c.row_div = mul * s;
      VS_TRY(conv1d_f32(c, st));
    }
    mul *= s;
    for (int j = 0; j < kDecKernels; ++j) {                              // MRF: xs = sum_j ResBlock1_j(x)  models.py:279-285
      const int n = i * kDecKernels + j, k = kResK[j];
      const float* cur = X;
      for (int mth = 0; mth < kDecDils; ++mth) {
        const bool last = (mth == kDecDils - 1);
        c = ConvF32();
        c.R = Rout; c.row_utt = valid; c.row_div = mul; c.in_ld = cout; c.Cin = cout; c.out_ld = cout; c.Cout = cout;
        c.k = k; c.in_slope = 0.1f;
        c.in = cur; c.w = w.c1[n][mth].w; c.bias = w.c1[n][mth].b; c.out = T; c.dil = kResD[mth];
        c.pad_l = (k - 1) / 2;
        VS_TRY(conv1d_f32(c, st));                                       // xt = c1(lrelu(x))
        float* dst = last ? S : (mth == 0 ? A : B);
        c.in = T; c.w = w.c2[n][mth].w; c.bias = w.c2[n][mth].b; c.out = dst; c.dil = 1;
        c.res = cur; c.res_ld = cout;                                    // x = c2(lrelu(xt)) + x
        if (last) { c.out_scale = 1.f / kDecKernels; c.accumulate = (j > 0); }   // x = xs / num_kernels
        VS_TRY(conv1d_f32(c, st));
        cur = dst;
      }
    }
  }
  // x = leaky_relu(x) [default slope 0.01, quirk Q3]; conv_post (no bias); tanh   models.py:286-288
  c = ConvF32();
  c.R = R * mul; c.row_utt = valid; c.row_div = mul; c.in = S; c.in_ld = 32; c.Cin = 32; c.in_slope = 0.01f;
  c.w = w.post_w; c.out = wave; c.out_ld = 1; c.Cout = 1; c.k = 7; c.pad_l = 3; c.act = 2;
  VS_TRY(conv1d_f32(c, st));
  return VS_OK;
}

}  // namespace vs
