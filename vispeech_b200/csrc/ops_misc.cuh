// Declarations of the small memory-bound kernels in ops_misc.cu.
#pragma once
#include "common.cuh"

namespace vs {
int embed_rows(const int32_t* ids, const float* emb, float* x, int R, int n_vocab, cudaStream_t st);
int add_speaker_rows(const float* x, const float* tab, const VsRows& rows, float* out, int C, cudaStream_t st);
int duration_rows(const float* logw, const double* ctrl, int mode, float scale, const VsRows& rows, double* dur,
                  cudaStream_t st);
int pitch_rows(const float* pred, const float* ctrl, int mode, float scale, const VsRows& rows, float* lf0, float* f0,
               cudaStream_t st);
int energy_rows(const float* pred, const float* ctrl, int mode, float scale, const VsRows& rows, float* norm,
                float* energy, cudaStream_t st);
int prenet_add(float* x, const float* v, const float* w, const float* b, const VsRows& rows, cudaStream_t st);
int lr_count(const VsRows& rows, const double* dur, int32_t* cum, int32_t* frames, cudaStream_t st);
int lr_gather(const VsRows& rp, const VsRows& rf, const float* xp, const int32_t* cum, float* xf, int32_t* lr_index,
              cudaStream_t st);
int prior_sample(const float* stats, const float* noise /*null: Philox(seed)*/, uint64_t seed, float ns, const VsRows& rows,
                 float* m_p, float* logs_p, float* z_p, cudaStream_t st);
int randn_fill(float* out, int64_t n, uint64_t seed, cudaStream_t st);
int wn_gate(const float* a, const float* cond, int cond_ld, int cond_off, const VsRows& rows, float* acts,
            cudaStream_t st);
int wn_update(const float* rs, int rs_ld, int last, int first, const VsRows& rows, float* h, float* skip,
              cudaStream_t st);
int coupling_update(float* z, int z_off, const float* m, float sign, const VsRows& rows, cudaStream_t st);
int mask_frames(const VsRows& rows, int max_len, int32_t* row_utt_out, cudaStream_t st);
int masked_copy(const float* x, const int32_t* row_utt, float* out, int R, int C, cudaStream_t st);
int unpack_rows(const VsRows& rows, const float* x, int C, int mul, int t_max, float* out, cudaStream_t st);
int pcm16(const float* x, int B, int T, const int32_t* n_samples, int decimate, const float* fir, int ntaps, int16_t* out,
          int T_out, cudaStream_t st);
int mel_frame_rows(const VsRows& rows, const float* wave, int t_max, const int32_t* n_samples, int hop, int ld, int pad,
                   float* out, cudaStream_t st);
int mel_magnitude(const float* dft, int ld_in, int im_off, int n_bins, int ld_out, int R, float* mag, cudaStream_t st);
int mel_unpack(const VsRows& rows, const float* x, int ld, int C, int t_max, int take_log, float* out, cudaStream_t st);
const char* last_error();
}  // namespace vs
