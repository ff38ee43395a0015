// HiFi-GAN Generator (reference models.py:244-297, modules.py:187-229) - weights view + entry points.
#pragma once
#include <functional>
#include <string>
#include "common.cuh"

namespace vs {

constexpr int kDecStages = 4;
constexpr int kDecKernels = 3;                       // resblock_kernel_sizes = [3,7,11]
constexpr int kDecDils = 3;                          // dilations (1,3,5)
constexpr int kResK[kDecKernels] = {3, 7, 11};
constexpr int kResD[kDecDils] = {1, 3, 5};
constexpr int kUpRate[kDecStages] = {8, 8, 4, 2};
constexpr int kUpKernel[kDecStages] = {16, 16, 4, 4};
constexpr int kUpPad[kDecStages] = {4, 4, 0, 1};     // (k - u) / 2, models.py:259
constexpr int kStageC[kDecStages + 1] = {512, 256, 128, 64, 32};

struct ConvW { const float* w; const float* b; };                 // fp32 [k][Cin][Cout], [Cout]
struct ConvW16 { const __half* w; const float* b; };       // f16 UMMA-packed (umma_conv.cu), fp32 bias

struct DecoderW {
  // fp32 (precision=1 cross-check path)
  ConvW pre;                                  // [7][192][512]
  const float* cond_tab;                      // [n_spk][512] = dec.cond(emb_g) incl. bias (models.py:273-274)
  ConvW ups[kDecStages];                      // [s][K/s][Cin][Cout] polyphase taps
  ConvW c1[kDecStages * kDecKernels][kDecDils], c2[kDecStages * kDecKernels][kDecDils];
  const float* post_w;                        // [7][32][1]
  // f16 tcgen05 path
  ConvW16 pre16, ups16[kDecStages];
  ConvW16 c1_16[kDecStages * kDecKernels][kDecDils], c2_16[kDecStages * kDecKernels][kDecDils];
  // host copies of the ResBlock biases of the stages with fused kernels (C <= 128): they travel in the kernels' parameter blocks
  float bias_host[kDecStages * kDecKernels][kDecDils][2][128];
  float post_w_host[7 * 32];                  // conv_post weights for umma_mrf.cu (kernel parameters)
};

using FetchFn = std::function<int(const std::string&, int64_t, int32_t, const void**)>;
int resolve_decoder(const FetchFn& fetch, int n_speakers, DecoderW* out);
int resolve_decoder_f16(const FetchFn& fetch, DecoderW* out);
int ups_taps(int stage, int* pad_l);   // taps of the polyphase form of ups[stage] (union over phases)
int64_t decoder_ws_floats(int n_rows_frame);
int decode_f32(const DecoderW& w, const VsRows& rows, const float* z, int max_len, float* wave, Workspace& ws,
               cudaStream_t st);
bool decoder_two_streams(int frame_rows);      // option decoder_streams resolved for a call of this many frame rows (decoder_umma.cu)
int decode_f16(const DecoderW& w, const VsRows& rows, const float* z, int max_len, float* wave, Workspace& ws,
                cudaStream_t st);

}  // namespace vs
