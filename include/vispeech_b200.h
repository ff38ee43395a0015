/*
 * vispeech_b200.h - C ABI of the B200-native SynthesizerTrn.infer hot path.
 *
 * The reference (innnky/vispeech) has no FFI / plugin layer: its boundary is the Python method
 * SynthesizerTrn.infer (models.py:672-722) fed by utils.load_checkpoint (utils.py:21-51).
 * This header is the boundary a maintainer would bind instead (INTEGRATION.md shows the ctypes
 * stub); vispeech_b200/synthesizer.py is that binding, keeping the reference's call surface.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - every call is stream-ordered on `stream` (a cudaStream_t passed as void*), never
 *     synchronises, never allocates: the caller owns outputs and the workspace;
 *   - return value 0 = ok; otherwise a VS_ERR_* code, text via vs_last_error();
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 *
 * Ragged rows ("VsRows"): a batch is a single long sequence of rows.  Utterance b owns rows
 * [utt_start[b], utt_start[b]+utt_len[b]); between utterances lie >= `gap` rows that are kept
 * exactly zero, so a convolution across the gap sees the zero padding a batch-1 reference call
 * would see (per-utterance batch-1 semantics).  row_utt[r] is the owning
 * utterance or -1.  Phoneme-level and frame-level tensors each have their own VsRows.
 * fp32 activations are row-major [n_rows][C].
 */
#ifndef VISPEECH_B200_H
#define VISPEECH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VS_OK 0
#define VS_ERR_INVALID 1      /* bad argument / unsupported shape */
#define VS_ERR_MISSING 2      /* weight tensor not registered */
#define VS_ERR_CUDA 3         /* CUDA runtime error (text in vs_last_error) */
#define VS_ERR_WORKSPACE 4    /* workspace too small */

#define VS_DTYPE_F32 0
#define VS_DTYPE_F16 1
#define VS_DTYPE_I64 2
#define VS_DTYPE_F64 3

typedef struct VsModel VsModel;

/* configs/config.json:39-90 + inference.py:26-33.  v1 kernels are specialised for the reference
 * config; vs_model_create rejects anything else with VS_ERR_INVALID. */
typedef struct {
  int32_t n_vocab;            /* 519  text/symbols.py:22 */
  int32_t hidden;             /* 192 */
  int32_t filter;             /* 768 */
  int32_t n_heads;            /* 2 */
  int32_t n_layers;           /* 4  (text encoder, frame prior) */
  int32_t pitch_layers;       /* 6  models.py:498 */
  int32_t window;             /* 4  attentions.py:14 */
  int32_t gin;                /* 256 */
  int32_t n_speakers;         /* 200 */
  int32_t flow_layers;        /* 4 */
  int32_t n_flows;            /* 4 */
  int32_t upsample_initial;   /* 512 */
  int32_t hop;                /* 512 = 8*8*4*2 */
} VsConfig;

typedef struct {
  int32_t n_utt;
  int32_t n_rows;             /* allocated rows; every kernel writes all of them (zeros where invalid) */
  int32_t max_len;            /* max over utt_len (host-known; sizes attention grids) */
  int32_t reserved;
  const int32_t* row_utt;     /* [n_rows]  utterance index or -1 */
  const int32_t* utt_start;   /* [n_utt] */
  const int32_t* utt_len;     /* [n_utt] */
  const int32_t* sid;         /* [n_utt]   speaker id (emb_g row, models.py:674) */
} VsRows;

const char* vs_last_error(void);
int vs_version(void);
int64_t vs_launch_count(void);
/* process-wide knobs. "tf32_min_rows": convs over at least this many rows run on the tensor cores in TF32 (default
 * 4096; fewer rows stay on the fp32 CUDA-core kernel).  Used by the parity tests to force either path.
 * "x3_min_rows": convs over at least this many rows (and below tf32_min_rows, or phoneme level) run in the error-
 * compensated three-term form on the tensor cores (fp16 hi/lo, or 3xTF32 with split16 = 0; fp32-level accuracy); default 256.
 * "tf32_prior": 0 (default) = the frame prior network and the (m_p, logs_p) projection use 3xTF32 at every size (prior
 * sampling amplifies their error), 1 = plain TF32 above tf32_min_rows like the flow (A/B measurements).
 * "fused_respair": 0 = never, 1 = only the C=32 stage's ResBlock iterations run as one fused conv-pair kernel,
 * 2 (default) = wherever both weight sets fit in shared memory (C=32 and C=64, all k; C=64 k=11 in its TIGHT form).
 * "attention_mma": 1 (default) = sequences of >= 128 rows use the tcgen05 attention kernel (csrc/attention_umma.cu: fp16
 * hi/lo operands, S and O in TMEM, band terms by a CUDA-core fix-up), shorter ones the fp32 CUDA-core kernel; 0 = always
 * CUDA cores, 2 = always the 3xTF32 mma.sync kernel, 3 = the same in plain TF32 (A/B only), 4 = always tcgen05.
 * "wn_fused": 1 (default) = on the plain-TF32 route every WN layer is ONE kernel on planar fp32 state (csrc/umma_wn.cu),
 * 0 = in_layer and res_skip as two launches of the generic TF32 conv.
 * "mrf_fused": 1 (default) = the decoder's last MRF stage (C = 32: three ResBlocks, sum, conv_post, tanh) is ONE kernel with the
 * residual stream and the MRF sum in fp32 in TMEM (csrc/umma_mrf.cu), 0 = the chain of fused conv-pair kernels + conv_post (A/B).
 * "pair_conv": the decoder's wide convs on a CTA pair (tcgen05 cta_group::2: weights resident, split between the two CTAs' shared
 * memories; csrc/umma_pair.cu): 1 (default) = Cin = Cout = 128, 2 = also Cin = Cout = 256 at k = 3 (0.107 -> 0.083 ms per launch alone, no
 * gain inside the decoder), 0 = the single-CTA kernel (A/B).
 * "pdl": bit mask of the kernel groups launched with programmatic dependent launch (the next kernel's launch, CTA scheduling and prologue
 * overlap its predecessor's tail; every such kernel executes griddepcontrol.wait before it touches global memory): 1 = three-term conv,
 * LayerNorm, rows_to_split, the fp32 cluster conv; 32 = the small element-wise kernels of the latent stages; 16 = CUDA-core attention,
 * row_dot; 2 / 64 / 128 = the frame-level attention kernel / its tile re-layout / its band fix-up; 4 = decoder (always on for calls below
 * 2048 frame rows); 8 = flow.  Default 33 (the groups that measured faster; 64 alone costs the frame prior 0.3 ms), 0 = every launch
 * fully serialised (A/B).
 * "conv_spread": 1 (default) = a decoder conv with streamed weights and fewer 128-row tiles than half the SMs is also cut along its output
 * channels (n-blocks narrowed to >= 64 columns, one CTA per (row tile, n-block group)), so the first ConvTranspose of a 4-second call
 * runs on 96 CTAs instead of 3; results are bit-identical to 0 (= row tiles only, A/B).  Batch-size calls never take it.
 * "attention_small": 1 (default) = the CUDA-core attention kernel takes 16 queries per CTA instead of 64 when the batch is too small to
 * fill half the SMs with 64-query CTAs (the batch-1 latency path), 0 = always 64.
 * "tap_pairs": the fused ResBlock iterations of the C = 64 stage may issue their conv taps in PAIRS as N = 128 MMAs (half the shared-memory
 * operand traffic per tap; the epilogue re-aligns the odd taps' half by a lane shuffle + a small exchange between lane quarters): 0 (default)
 * = never (measured slower in the whole decoder), 1 = at k = 11 only, 2 = at every k (csrc/umma_respair.cu).
 * "coupling_min_rows": frame rows from which the flow takes the one-kernel coupling layer.  Default 1 = always: 76 -> 4 launches per flow
 * pass on the batch-1 path (C1 3.66 -> 3.18 ms per call), z within 4.2e-4 of the reference on the golden utterances (bar 1e-2), and an
 * utterance's flow no longer changes regime with the batch it is in; 4096 = small calls keep the fp32 / three-term kernels (z within 1e-5).
 * "coupling_fused": 1 (default) = in the plain-TF32 regime (>= tf32_min_rows frame rows) every coupling layer of the flow is ONE kernel
 * (pre, the 4-layer WN stack, post and the x1 update; residual stream and skip sum in fp32 in TMEM, fp16 operands; csrc/umma_coupling.cu),
 * 0 = pre / per-layer WN kernels / post / update as separate launches (A/B).
 * "pair_fused": 1 (default) = the k = 3 ResBlock iterations of the C = 128 stage run as ONE kernel per iteration on a CTA pair (conv1,
 * leaky-ReLU, conv2, residual; the intermediate stays in shared memory; csrc/umma_pairfused.cu), 2 = also every iteration of the C = 64
 * stage (measured: no gain over the single-CTA fused kernel there), 0 = two CTA-pair conv launches per iteration (A/B).
 * "resblock_fused": 1 (default) = the k = 3 ResBlock of the C = 64 stage is ONE kernel with its residual stream in fp32 in TMEM
 * (csrc/umma_resblock.cu), 0 = three fused conv-pair launches (A/B).
 * "split16": 1 (default) = every conv in the 3xTF32 regime of the encoders / predictors / projection runs as the three-term fp16
 * hi/lo conv on tcgen05 kind::f16 (csrc/umma_split.cu: same fp32-level accuracy, TMA-fed planar operands), 0 = 3xTF32 (A/B).
 * "tf32_cluster": 1 (default) | 2 = two CTAs of a cluster share every weight slab of the TF32 conv by TMA multicast
 * (bit-identical; no gain measured).  "decoder_streams": 2 = the k=11 ResBlock chains of each decoder stage run on a side
 * stream (fork / join by events, graph-capturable; bit-identical), 1 = one stream, 0 (default) = 2 for calls below 2048 frame rows (where
 * the kernels are small enough to run side by side: chunked 60 s decode 32.8 -> 28.1 ms) and 1 above (no gain at batch size).  "respair_grid_div": co-scheduling experiments (tools/cosched_pairs.py).
 * "umma_timing_buffer": diagnostics - a device pointer (or 0) to >= 148*24 int64 where the tcgen05 kernels built with
 * -DVS_UMMA_TIMING leave per-CTA clocks spent waiting on each mbarrier (tools/conv_timing.py, tf32_timing.py, wn_timing.py).
 * vs_set_option sets the PROCESS DEFAULT of an option; vs_model_set_option overrides it for one model handle.  Every entry
 * point that takes a model freezes "defaults overlaid with that model's overrides" into a thread-local snapshot for the
 * duration of the call, so models on different threads / devices never see each other's settings, and per-kernel launch
 * attributes are kept per device (a process may hold models on several GPUs). */
int vs_set_option(const char* name, int64_t value);
int vs_model_set_option(VsModel* m, const char* name, int64_t value);

/* ---- weights: replaces utils.load_checkpoint (utils.py:21-51) + the implicit weight-norm fold.
 * Tensors are registered by name in the PACKED layouts listed in vispeech_b200/packing.py
 * (folded weight norm, [tap][Cin][Cout] fp32 conv weights, per-speaker conditioning tables,
 * planar f16 decoder weights).  The model keeps the pointers; the caller keeps ownership. */
int vs_model_create(const VsConfig* cfg, VsModel** out);
void vs_model_destroy(VsModel* m);
int vs_model_set_tensor(VsModel* m, const char* name, const void* ptr, int64_t numel, int32_t dtype);
int vs_model_finalize(VsModel* m);                       /* VS_ERR_MISSING if any tensor is absent */

/* workspace bytes needed by any single call below for these row counts */
int64_t vs_workspace_bytes(const VsModel* m, int32_t n_rows_phoneme, int32_t n_rows_frame);
/* the same split by consumer: every call except vs_hifigan_decode | vs_hifigan_decode at the given precision */
int64_t vs_workspace_bytes_latent(const VsModel* m, int32_t n_rows_phoneme, int32_t n_rows_frame);
int64_t vs_workspace_bytes_decoder(const VsModel* m, int32_t n_rows_frame, int32_t precision);

/* ---- a3-a7: TextEncoder.forward (models.py:168-174) = embedding*sqrt(H) + 4-layer attentions.Encoder */
int vs_text_encode(const VsModel* m, const VsRows* rows_p, const int32_t* ids_rows /*[n_rows], -1 in gaps*/,
                   float* x_out /*[n_rows][192]*/, void* ws, int64_t ws_bytes, void* stream);

/* ---- a8-a11: duration / pitch / energy predictors + prenets (models.py:681-708).
 * *_mode: 0 = predict, scale by *_scale (None -> 1, scalar control);  2 = per-phoneme override in
 * *_ctrl (durations in frames, F0 in Hz, raw energy).  x is updated in place (x += prenet(..)). */
int vs_variance_adapter(const VsModel* m, const VsRows* rows_p, float* x /*[n_rows][192] in/out*/,
                        int32_t dur_mode, float dur_scale, const double* dur_ctrl /*[n_rows]*/,
                        int32_t pitch_mode, float pitch_scale, const float* pitch_ctrl,
                        int32_t energy_mode, float energy_scale, const float* energy_ctrl,
                        double* duration_out /*[n_rows] as models.py:688 returns it*/,
                        float* f0_out, float* energy_out, void* ws, int64_t ws_bytes, void* stream);

/* ---- a12-a13: LengthRegulator (models.py:398-427).  Step 1: n_i = max(int(d_i),0), inclusive
 * per-utterance prefix sum and per-utterance frame counts (host reads frames[] to lay out frame rows).
 * Step 2: gather.  lr_index[r] = phoneme index within its utterance, -1 in gaps: bit-exact with the
 * reference expansion for identical durations. */
int vs_length_regulate_count(const VsRows* rows_p, const double* duration /*[n_rows]*/,
                             int32_t* cum_out /*[n_rows] inclusive*/, int32_t* frames_out /*[n_utt]*/, void* stream);
int vs_length_regulate_gather(const VsRows* rows_p, const VsRows* rows_f, const float* x_p, const int32_t* cum,
                              float* x_f /*[rows_f.n_rows][192]*/, int32_t* lr_index /*[rows_f.n_rows]*/, void* stream);

/* ---- a14-a16: FramePriorNet + Projection + prior sample (models.py:715-718).
 * noise: the eps of models.py:718 ([n_rows][192], injected for parity runs) or NULL = drawn inside the sampling kernel from
 * a counter-based Philox4x32-10 stream keyed by noise_seed (N(0,1) by Box-Muller; the same seed gives the same eps). */
int vs_frame_prior(const VsModel* m, const VsRows* rows_f, const float* x_f, const float* noise, uint64_t noise_seed,
                   float noise_scale, float* x_frame_out, float* m_p, float* logs_p, float* z_p,
                   void* ws, int64_t ws_bytes, void* stream);

/* N(0,1) samples of the same Philox stream as a tensor: out[i] = what vs_frame_prior(noise = NULL, noise_seed = seed) would use
 * for element i.  For callers that replay a captured CUDA graph (kernel arguments are frozen: eps is read from memory). */
int vs_randn(float* out, int64_t n, uint64_t seed, void* stream);

/* ---- a17: ResidualCouplingBlock reverse (models.py:202-209), in place on z ([n_rows][192]) */
int vs_flow_reverse(const VsModel* m, const VsRows* rows_f, float* z, void* ws, int64_t ws_bytes, void* stream);

/* ---- 8(f) voice_conversion (models.py:724-732): PosteriorEncoder.forward (models.py:233-241) on a linear spectrogram
 * laid out as ragged rows [n_rows][c_in] (c_in = spec channels zero-padded to a multiple of 96, see packing.py), and the
 * flow in forward direction (models.py:203-205), in place.  Needs the enc_q.* tensors (VS_ERR_MISSING otherwise). */
int vs_posterior_encode(const VsModel* m, const VsRows* rows_f, const float* spec, const float* noise /*[n_rows][192] or NULL*/,
                        uint64_t noise_seed, float* z, float* m_q, float* logs_q, void* ws, int64_t ws_bytes, void* stream);
int vs_flow_forward(const VsModel* m, const VsRows* rows_f, float* z, void* ws, int64_t ws_bytes, void* stream);

/* ---- a18: HiFi-GAN Generator (models.py:271-290).  max_len < 0 = no truncation (models.py:720).
 * wave_out is [rows_f.n_rows * hop] in ragged order.  precision: 0 = f16 tcgen05 path (product),
 * 1 = fp32 SIMT path (test-only cross-check of the same math, NOT a fallback). */
int vs_hifigan_decode(const VsModel* m, const VsRows* rows_f, const float* z, int32_t max_len, float* wave_out,
                      int32_t precision, void* ws, int64_t ws_bytes, void* stream);

/* ragged rows -> reference layout [n_utt][C][t_max] (zero padded), rows_per_step = samples per row entry */
int vs_unpack_rows(const VsRows* rows, const float* x /*[n_rows*rows_mul][C]*/, int32_t C, int32_t rows_mul,
                   int32_t t_max, float* out /*[n_utt][C][t_max]*/, void* stream);

/* ---- 8(f) waveform post-processing on the GPU (replaces scipy wavfile.write + `ffmpeg -ar 22050`, inference_api.py:50-51):
 * wave [n_utt][t_max] fp32 -> out [n_utt][t_out] s16 = clip(rint(32768 * FIR-decimated wave)); samples beyond
 * n_samples[b] count as zero.  decimate = 1 with n_taps = 0 is a plain conversion. */
int vs_wave_pcm16(const float* wave, int32_t n_utt, int32_t t_max, const int32_t* n_samples, int32_t decimate,
                  const float* fir /*[n_taps] or NULL*/, int32_t n_taps, int16_t* out, int32_t t_out, void* stream);

/* ---- 8(f): linear / log-mel spectrogram of waveforms (reference mel_processing.py:50-112: spectrogram_torch feeds
 * voice_conversion, mel_spectrogram_torch is the evaluation metric of train.py:303-313).  n_fft = win = 4 * hop, hann,
 * center = False after a reflect pad of (n_fft - hop) / 2.  rows: one entry per utterance with n_frames[b] + 3 rows
 * (frame j = rows j..j+3); wave [n_utt][t_max] fp32, n_samples[b] valid samples (>= (n_fft - hop) / 2 + 1).
 * dft_packed / mel_packed: the windowed DFT basis and the mel filterbank as 3xTF32 slabs (vispeech_b200/mel.py).
 * spec_out [n_utt][n_bins][frames_max] and/or mel_out [n_utt][n_mels][frames_max] (either may be NULL). */
int vs_mel_spectrogram(const VsRows* rows, const float* wave, int32_t t_max, const int32_t* n_samples, int32_t hop,
                       int32_t n_bins, int32_t n_mels, const float* dft_packed, const float* mel_packed, int32_t frames_max,
                       float* spec_out, float* mel_out, void* ws, int64_t ws_bytes, void* stream);

/* ---- op-level entry points (used by the parity tests; same kernels as above) */
int vs_op_conv1d_f32(const float* in, int32_t in_ld, const float* w /*[k][Cin][Cout]*/, const float* bias,
                     float* out, int32_t out_ld, int32_t n_rows, int32_t c_in, int32_t c_out, int32_t k, int32_t dil,
                     int32_t pad_l, float in_slope, int32_t act, const int32_t* row_utt, int32_t row_div, void* stream);
int vs_op_layernorm(const float* a, const float* b, const float* gamma, const float* beta, float* out,
                    int32_t n_rows, int32_t C, const int32_t* row_utt, void* stream);
/* ws: scratch for the tcgen05 kernel (>= n_rows * 8 KB + n_utt * 200 KB), or NULL = the register-accumulator kernels only */
int vs_op_rel_attention(const VsRows* rows, const float* qkv /*[n_rows][576]*/, const float* emb_rel_k,
                        const float* emb_rel_v, float* out /*[n_rows][192]*/, void* ws, int64_t ws_bytes, void* stream);
/* TF32 tcgen05 conv over fp32 row-major rows (csrc/umma_tf32.cu): out = act(conv(in) + bias), zeros on invalid rows.
 * split3 = 1: error-compensated 3xTF32 (weights packed with pack_tf32(split3=True)), fp32-level accuracy. */
int vs_op_conv1d_tf32(const float* in, int32_t in_ld, const float* w_packed, const float* bias, float* out, int32_t out_ld,
                      int32_t n_rows, int32_t c_in, int32_t c_out, int32_t taps, int32_t dil, int32_t pad_l, int32_t act,
                      int32_t split3, const int32_t* row_utt, void* stream);
/* One WN layer (modules.py:153-176) in one kernel, plain TF32 (csrc/umma_wn.cu): acts = tanh.sigmoid(in_layer(h) + cond[sid]);
 * rs = res_skip(acts); h_out = (h + rs[:, :192]) on valid rows, 0 on gap rows; skip = (first ? 0 : skip) + rs[:, 192:] (the
 * last layer's res_skip has 192 outputs, all skip; h_out may be NULL).  h_in / h_out / skip are row-major [n_rows][192] here
 * (the model keeps them planar across the stack); w_in_packed: gate-interleaved columns (packing.gate_columns), pack_tf32
 * slabs; cond: [n_spk][cond_ld] rows of this layer's 384 gate-interleaved columns; ws >= 3 * n_rows * 192 floats. */
/* fp32-accurate conv on tcgen05 kind::f16 with fp16 hi/lo operands (csrc/umma_split.cu; weights packed with pack_split16):
 * out = act(conv(in) + bias) over fp32 row-major rows, zeros on invalid rows; act: 0 none, 1 ReLU; c_in <= 192 or a multiple of
 * 192 (K-slices).  ws >= n_rows * (4 c_in + 4 c_out c_in/192) bytes. */
int vs_op_conv1d_split(const float* in, int32_t in_ld, const void* w_packed, const float* bias, float* out, int32_t out_ld,
                       int32_t n_rows, int32_t c_in, int32_t c_out, int32_t taps, int32_t dil, int32_t pad_l, int32_t act,
                       const int32_t* row_utt, void* ws, int64_t ws_bytes, void* stream);
int vs_op_wn_layer(const float* h_in, const float* w_in_packed, const float* b_in, const float* cond, int32_t cond_ld,
                   const int32_t* cond_idx, const float* w_rs_packed, const float* b_rs, const int32_t* row_utt,
                   int32_t n_rows, int32_t first, int32_t last, float* h_out, float* skip, void* ws, int64_t ws_bytes,
                   void* stream);
/* f16 tcgen05 implicit-GEMM conv on planar [C/8][n_rows][8] activations (csrc/umma_conv.cu):
 * y = conv(in) + bias + res;  out_raw = y;  out_act = leaky_relu(y*act_scale, act_slope); either output may be
 * null.  up > 1 = polyphase ConvTranspose1d (column gn -> phase gn/Cout, output row up*r+phase). */
int vs_op_conv1d_umma(const void* in_planar, const void* w_packed, const float* bias, const void* res_planar,
                      void* out_raw, void* out_act, int32_t n_rows, int32_t c_in, int32_t n_cols, int32_t taps,
                      int32_t dil, int32_t pad_l, int32_t up, float act_slope, float act_scale,
                      const int32_t* row_utt, int32_t row_div, void* stream);

/* the same with the c2 epilogue's extra inputs: res2 (the MRF running sum) and res_inv_slope != 0 (res holds a = lrelu(x, 1 / slope);
 * the residual added is x = min(a, a * res_inv_slope)).  Cin = N = 128 (and 256 at taps <= 3 with option pair_conv = 2) run on a CTA
 * pair (tcgen05 cta_group::2, csrc/umma_pair.cu) unless option "pair_conv" is 0. */
int vs_op_conv1d_umma2(const void* in_planar, const void* w_packed, const float* bias, const void* res_planar, const void* res2_planar,
                       float res_inv_slope, void* out_raw, void* out_act, int32_t n_rows, int32_t c_in, int32_t n_cols, int32_t taps,
                       int32_t dil, int32_t pad_l, int32_t up, float act_slope, float act_scale,
                       const int32_t* row_utt, int32_t row_div, void* stream);

/* fused ResBlock1 iteration y = c2(lrelu(c1(lrelu(x)))) + x (+ res2) on planar f16 rows (csrc/umma_respair.cu) */
int vs_op_respair(const void* x_planar, const void* w1_packed, const void* w2_packed, const float* b1, const float* b2,
                  const void* res2_planar, void* out_raw, void* out_act, int32_t n_rows, int32_t channels, int32_t taps,
                  int32_t dil, float act_slope, float act_scale, const int32_t* row_utt, int32_t row_div, void* stream);

/* The decoder's whole last MRF stage (C = 32) + conv_post + tanh in one kernel (csrc/umma_mrf.cu; reference models.py:279-288,
 * modules.py:210-223): x0 = x_hi + x_lo (planar f16 [4][n_rows][8] each) -> 3 ResBlock1 (k = 3, 7, 11; d = 1, 3, 5) -> sum / 3 ->
 * leaky_relu(0.01) -> conv_post (32 -> 1, k = 7) -> tanh -> wave [n_rows].  w_packed[18] / b_host[18]: per ResBlock j and
 * iteration m, (c1, c2) at index (3 j + m) * 2 + {0, 1}; weights are device pointers (pack_umma slabs), biases and post_w
 * ([7][32]) HOST pointers (they travel in the kernel's parameter block). */
int vs_op_mrf32(const void* x_hi, const void* x_lo, const void* const* w_packed, const float* const* b_host,
                const float* post_w_host, const int32_t* row_utt, int32_t row_div, int32_t n_rows, float* wave, void* stream);

/* A whole ResBlock1 of the C = 64 decoder stage's k = 3 branch in one kernel (csrc/umma_resblock.cu; reference modules.py:210-223:
 * three iterations x = x + c2(lrelu(c1(lrelu(x)))), dilations 1, 3, 5), residual stream in fp32 in TMEM.  a_planar = lrelu(x0, 0.1) as
 * planar f16 [8][n_rows][8]; w_packed[6] / b_host[6]: (c1, c2) of iteration m at index 2 m + {0, 1} (device pack_umma slabs, HOST biases);
 * out_raw = the ResBlock's output x3, planar f16, zeros on gap rows. */
int vs_op_resblock64(const void* a_planar, const void* const* w_packed, const float* const* b_host, const int32_t* row_utt,
                     int32_t row_div, int32_t n_rows, void* out_raw, void* stream);

#ifdef __cplusplus
}
#endif
#endif
