// Weight registry + subsystem orchestration + the C ABI of include/vispeech_b200.h.
// Each vs_* subsystem mirrors one block of SynthesizerTrn.infer (reference models.py:672-722).
#include <string>
#include <unordered_map>
#include <vector>

#include "ops_misc.cuh"
#include "decoder.cuh"
#include "umma_conv.cuh"
#include "umma_tf32.cuh"

namespace vs {

struct Tensor { const void* ptr; int64_t numel; int32_t dtype; };

struct EncLayer {
  const float *wqkv, *bqkv, *wo, *bo, *ek, *ev, *g1, *b1, *g2, *b2, *w1, *bf1, *w2, *bf2;
  const float *t_wqkv, *t_wo, *t_w1, *t_w2;            // TF32 slab copies (umma_tf32.cuh)
  const float *x_wqkv, *x_wo, *x_w1, *x_w2;            // 3xTF32 [hi|lo] slab copies
  const __half *s_wqkv, *s_wo, *s_w1, *s_w2;           // fp16 hi/lo slabs of umma_split.cu (packing.py pack_split16)
};
constexpr int kMaxWnLayers = 16;
// WN (modules.py:111-184) weights: flow coupling layers (4 layers) and the posterior encoder (16 layers)
struct WnW {
  int n_layers = 0;
  const float* cond_tab = nullptr;                       // [n_spk][2H*L] = cond_layer(emb_g) incl. bias
  const float* cond_tab_gate = nullptr;                  // same, gate-interleaved columns (packing.py gate_columns)
  const float *in_w[kMaxWnLayers], *in_b[kMaxWnLayers], *rs_w[kMaxWnLayers], *rs_b[kMaxWnLayers];
  const float *t_in[kMaxWnLayers], *t_rs[kMaxWnLayers], *t_in_gate[kMaxWnLayers], *in_gate_b[kMaxWnLayers];   // TF32
  const float *x_in[kMaxWnLayers], *x_rs[kMaxWnLayers], *x_in_gate[kMaxWnLayers];                             // 3xTF32
};
struct FlowW {
  const float *pre_w, *pre_b, *post_w, *post_b;
  const float *t_pre, *t_post, *x_pre, *x_post;
  const __half* c16_w = nullptr;                         // the whole coupling layer as one kernel (umma_coupling.cu; packing.py pack_coupling)
  const float* c16_b = nullptr;
  WnW wn;
};
struct PosteriorW {                                      // enc_q (models.py:212-241), only needed by voice_conversion
  bool present = false;
  int c_in = 0;                                          // spec channels padded to a multiple of 96 (packing.py)
  const float *pre_w, *pre_b, *proj_w, *proj_b, *t_pre, *x_pre, *t_proj, *x_proj;
  WnW wn;
};

// Precision / engine policy for the GEMM-shaped convs upstream of the decoder (all rows counts are per call):
//   rows >= tf32_min_rows (4096) and the conv is frame level  -> tcgen05 kind::tf32, operands rounded to TF32
//   rows >= x3_min_rows (256)                                  -> tcgen05 3xTF32 (hi/lo split, fp32-level accuracy)
//   otherwise                                                  -> fp32 CUDA cores
// Phoneme-level convs (text encoder, predictors) never take the plain-TF32 route: their error feeds back through the
// F0 / energy prenets (measured: z error 7.5e-3 with plain TF32 there vs 4.6e-3 without; bar 1e-2).
// Options "tf32_min_rows" (4096), "x3_min_rows" (256), "tf32_prior" (0) and "wn_fused" (1) come from opts() (common.cuh).
// The frame prior network and the projection to (m_p, logs_p) stay on 3xTF32 at every size unless tf32_prior = 1:
// z_p = m_p + eps * exp(logs_p) amplifies their error (measured on C4: plain TF32 there gives |dz| = 1.2e-2 > the 1e-2 bar,
// 3xTF32 1.4e-4).
static int conv_rows(const ConvF32& c, const float* w_tf32, const float* w_x3, cudaStream_t st) {
  const bool shape_ok = !c.res && !c.accumulate && c.in_slope == 1.f && c.out_row_mul == 1 && c.out_row_off == 0 &&
                        c.out_scale == 1.f && c.act <= 1 && c.row_div == 1 && c.Cout % 32 == 0;
  const bool use_tf32 = shape_ok && w_tf32 && c.R >= opts().v[OPT_TF32_MIN_ROWS];
  const bool use_x3 = shape_ok && !use_tf32 && w_x3 && c.R >= opts().v[OPT_X3_MIN_ROWS];
  if (!use_tf32 && !use_x3) return conv1d_f32(c, st);
  UmmaTf32 u;
  u.in = c.in; u.in_ld = c.in_ld; u.w = use_tf32 ? w_tf32 : w_x3; u.split3 = use_x3 ? 1 : 0; u.bias = c.bias;
  u.ubias = c.ubias; u.ubias_ld = c.ubias_ld; u.ubias_idx = c.ubias_idx; u.out = c.out; u.out_ld = c.out_ld;
  u.row_utt = c.row_utt; u.R = c.R; u.Cin = c.Cin; u.N = c.Cout; u.taps = c.k; u.dil = c.dil; u.pad_l = c.pad_l; u.act = c.act;
  return umma_tf32(u, st);
}

}  // namespace vs

struct VsModel {
  VsConfig cfg;
  vs::Options overrides;        // vs_model_set_option: kOptUnset = follow the process default
  std::unordered_map<std::string, vs::Tensor> tensors;
  bool finalized = false;
  // resolved views
  const float* emb = nullptr;
  std::vector<vs::EncLayer> enc_text, enc_pitch, enc_prior;
  const float *dp_cond, *dp_w1, *dp_b1, *dp_g1, *dp_be1, *dp_w2, *dp_b2, *dp_g2, *dp_be2, *dp_wp, *dp_bp;
  const float *pp_cond, *pp_wf0, *pp_bf0;
  const float *ep_cond, *ep_w1, *ep_b1, *ep_g1, *ep_be1, *ep_w2, *ep_b2, *ep_g2, *ep_be2, *ep_wl, *ep_bl;
  const float *pitch_pre_w, *pitch_pre_b, *energy_pre_w, *energy_pre_b;
  const float *proj_w, *proj_b, *t_proj_w, *x_proj_w, *x_dp_w1, *x_ep_w1, *x_ep_w2;
  const __half *s_proj_w, *s_dp_w1, *s_ep_w1, *s_ep_w2;
  std::vector<vs::FlowW> flows;
  vs::PosteriorW enc_q;
  vs::DecoderW dec;
};

namespace vs {

static int fetch(VsModel* m, const std::string& name, int64_t numel, int32_t dtype, const void** out) {
  auto it = m->tensors.find(name);
  if (it == m->tensors.end()) { set_error("weight '%s' was never registered", name.c_str()); return VS_ERR_MISSING; }
  if (it->second.numel != numel || it->second.dtype != dtype) {
    set_error("weight '%s': expected numel=%lld dtype=%d, got numel=%lld dtype=%d", name.c_str(), (long long)numel,
              dtype, (long long)it->second.numel, it->second.dtype);
    return VS_ERR_INVALID;
  }
  *out = it->second.ptr;
  return VS_OK;
}
#define FETCH_F32(field, name, numel) VS_TRY(fetch(m, name, numel, VS_DTYPE_F32, reinterpret_cast<const void**>(&(field))))
#define FETCH_F16(field, name, numel) VS_TRY(fetch(m, name, numel, VS_DTYPE_F16, reinterpret_cast<const void**>(&(field))))

static int resolve_encoder(VsModel* m, const std::string& p, int n_layers, std::vector<EncLayer>* out) {
  const int H = kHidden, F = kFilter;
  out->resize(n_layers);
  for (int i = 0; i < n_layers; ++i) {
    EncLayer& L = (*out)[i];
    const std::string q = p + "." + std::to_string(i) + ".";
    FETCH_F32(L.wqkv, q + "wqkv", (int64_t)H * 3 * H);  FETCH_F32(L.bqkv, q + "bqkv", 3 * H);
    FETCH_F32(L.wo, q + "wo", (int64_t)H * H);          FETCH_F32(L.bo, q + "bo", H);
    FETCH_F32(L.ek, q + "ek", kRel * kHeadDim);         FETCH_F32(L.ev, q + "ev", kRel * kHeadDim);
    FETCH_F32(L.g1, q + "g1", H);  FETCH_F32(L.b1, q + "b1", H);
    FETCH_F32(L.g2, q + "g2", H);  FETCH_F32(L.b2, q + "b2", H);
    FETCH_F32(L.w1, q + "w1", (int64_t)3 * H * F);      FETCH_F32(L.bf1, q + "bf1", F);
    FETCH_F32(L.w2, q + "w2", (int64_t)3 * F * H);      FETCH_F32(L.bf2, q + "bf2", H);
    FETCH_F32(L.t_wqkv, "tf32." + q + "wqkv", (int64_t)H * 3 * H);  FETCH_F32(L.t_wo, "tf32." + q + "wo", (int64_t)H * H);
    FETCH_F32(L.t_w1, "tf32." + q + "w1", (int64_t)3 * H * F);      FETCH_F32(L.t_w2, "tf32." + q + "w2", (int64_t)3 * F * H);
    FETCH_F32(L.x_wqkv, "x3." + q + "wqkv", (int64_t)2 * H * 3 * H);  FETCH_F32(L.x_wo, "x3." + q + "wo", (int64_t)2 * H * H);
    FETCH_F32(L.x_w1, "x3." + q + "w1", (int64_t)2 * 3 * H * F);      FETCH_F32(L.x_w2, "x3." + q + "w2", (int64_t)2 * 3 * F * H);
    FETCH_F16(L.s_wqkv, "s16." + q + "wqkv", (int64_t)2 * H * 3 * H);   FETCH_F16(L.s_wo, "s16." + q + "wo", (int64_t)2 * H * H);
    FETCH_F16(L.s_w1, "s16." + q + "w1", (int64_t)2 * 3 * H * F);       FETCH_F16(L.s_w2, "s16." + q + "w2", (int64_t)2 * 3 * F * H);
  }
  return VS_OK;
}

static int resolve_wn(VsModel* m, const std::string& p, int L, int S, WnW* w) {
  const int H = kHidden;
  VS_REQUIRE(L <= kMaxWnLayers, "WN with %d layers (max %d)", L, kMaxWnLayers);
  w->n_layers = L;
  FETCH_F32(w->cond_tab, p + "cond_tab", (int64_t)S * 2 * H * L);
  FETCH_F32(w->cond_tab_gate, p + "cond_tab_gate", (int64_t)S * 2 * H * L);
  for (int l = 0; l < L; ++l) {
    const std::string q = p + std::to_string(l) + ".";
    const int rs = (l < L - 1) ? 2 * H : H;
    FETCH_F32(w->in_w[l], q + "in.w", 5 * H * 2 * H);  FETCH_F32(w->in_b[l], q + "in.b", 2 * H);
    FETCH_F32(w->rs_w[l], q + "rs.w", H * rs);         FETCH_F32(w->rs_b[l], q + "rs.b", rs);
    FETCH_F32(w->t_in[l], "tf32." + q + "in.w", 5 * H * 2 * H);  FETCH_F32(w->t_rs[l], "tf32." + q + "rs.w", H * rs);
    FETCH_F32(w->t_in_gate[l], "tf32." + q + "in_gate.w", 5 * H * 2 * H);  FETCH_F32(w->in_gate_b[l], q + "in_gate.b", 2 * H);
    FETCH_F32(w->x_in[l], "x3." + q + "in.w", 2 * 5 * H * 2 * H);  FETCH_F32(w->x_rs[l], "x3." + q + "rs.w", 2 * H * rs);
    FETCH_F32(w->x_in_gate[l], "x3." + q + "in_gate.w", 2 * 5 * H * 2 * H);
  }
  return VS_OK;
}

static int finalize(VsModel* m) {
  const int H = kHidden, S = m->cfg.n_speakers;
  FETCH_F32(m->emb, "emb", (int64_t)m->cfg.n_vocab * H);
  VS_TRY(resolve_encoder(m, "enc_p.encoder", m->cfg.n_layers, &m->enc_text));
  VS_TRY(resolve_encoder(m, "pitch_predictor.pitch_net", m->cfg.pitch_layers, &m->enc_pitch));
  VS_TRY(resolve_encoder(m, "frame_prior_net.fft_block", m->cfg.n_layers, &m->enc_prior));
  FETCH_F32(m->dp_cond, "dp.cond_tab", (int64_t)S * H);
  FETCH_F32(m->dp_w1, "dp.w1", 3 * H * 256);   FETCH_F32(m->dp_b1, "dp.b1", 256);
  FETCH_F32(m->dp_g1, "dp.g1", 256);           FETCH_F32(m->dp_be1, "dp.be1", 256);
  FETCH_F32(m->dp_w2, "dp.w2", 3 * 256 * 256); FETCH_F32(m->dp_b2, "dp.b2", 256);
  FETCH_F32(m->dp_g2, "dp.g2", 256);           FETCH_F32(m->dp_be2, "dp.be2", 256);
  FETCH_F32(m->dp_wp, "dp.wp", 256);           FETCH_F32(m->dp_bp, "dp.bp", 1);
  FETCH_F32(m->pp_cond, "pp.cond_tab", (int64_t)S * H);
  FETCH_F32(m->pp_wf0, "pp.wf0", H);           FETCH_F32(m->pp_bf0, "pp.bf0", 1);
  FETCH_F32(m->ep_cond, "ep.cond_tab", (int64_t)S * H);
  FETCH_F32(m->ep_w1, "ep.w1", 3 * H * 768);   FETCH_F32(m->ep_b1, "ep.b1", 768);
  FETCH_F32(m->ep_g1, "ep.g1", 768);           FETCH_F32(m->ep_be1, "ep.be1", 768);
  FETCH_F32(m->ep_w2, "ep.w2", 3 * 768 * 768); FETCH_F32(m->ep_b2, "ep.b2", 768);
  FETCH_F32(m->ep_g2, "ep.g2", 768);           FETCH_F32(m->ep_be2, "ep.be2", 768);
  FETCH_F32(m->ep_wl, "ep.wl", 768);           FETCH_F32(m->ep_bl, "ep.bl", 1);
  FETCH_F32(m->pitch_pre_w, "pitch_prenet.w", H * 3);   FETCH_F32(m->pitch_pre_b, "pitch_prenet.b", H);
  FETCH_F32(m->energy_pre_w, "energy_prenet.w", H * 3); FETCH_F32(m->energy_pre_b, "energy_prenet.b", H);
  FETCH_F32(m->proj_w, "proj.w", H * 2 * H);   FETCH_F32(m->proj_b, "proj.b", 2 * H);
  FETCH_F32(m->t_proj_w, "tf32.proj.w", H * 2 * H);       FETCH_F32(m->x_proj_w, "x3.proj.w", 2 * H * 2 * H);
  FETCH_F32(m->x_dp_w1, "x3.dp.w1", 2 * 3 * H * 256);
  FETCH_F32(m->x_ep_w1, "x3.ep.w1", 2 * 3 * H * 768);     FETCH_F32(m->x_ep_w2, "x3.ep.w2", 2 * 3 * 768 * 768);
  FETCH_F16(m->s_proj_w, "s16.proj.w", 2 * H * 2 * H);    FETCH_F16(m->s_dp_w1, "s16.dp.w1", 2 * 3 * H * 256);
  FETCH_F16(m->s_ep_w1, "s16.ep.w1", 2 * 3 * H * 768);    FETCH_F16(m->s_ep_w2, "s16.ep.w2", 2 * 3 * 768 * 768);
  const int L = m->cfg.flow_layers;
  m->flows.resize(m->cfg.n_flows);
  for (int f = 0; f < m->cfg.n_flows; ++f) {
    FlowW& w = m->flows[f];
    const std::string p = "flow." + std::to_string(f) + ".";
    FETCH_F32(w.pre_w, p + "pre.w", (H / 2) * H);   FETCH_F32(w.pre_b, p + "pre.b", H);
    FETCH_F32(w.post_w, p + "post.w", H * (H / 2)); FETCH_F32(w.post_b, p + "post.b", H / 2);
    FETCH_F32(w.t_pre, "tf32." + p + "pre.w", (H / 2) * H);   FETCH_F32(w.t_post, "tf32." + p + "post.w", H * (H / 2));
    FETCH_F32(w.x_pre, "x3." + p + "pre.w", 2 * (H / 2) * H); FETCH_F32(w.x_post, "x3." + p + "post.w", 2 * H * (H / 2));
    VS_TRY(resolve_wn(m, p, L, S, &w.wn));
    if (L == 4 && m->tensors.count("c16." + p + "w")) {
      FETCH_F16(w.c16_w, "c16." + p + "w", (int64_t)92 * 24 * 96 * 8);
      FETCH_F32(w.c16_b, "c16." + p + "b", (int64_t)4 * H + H / 2 + (int64_t)S * 4 * 2 * H);
    }
  }
  // posterior encoder: optional (inference never touches it; voice_conversion does)
  m->enc_q.present = m->tensors.count("enc_q.pre.w") != 0;
  if (m->enc_q.present) {
    PosteriorW& q = m->enc_q;
    q.c_in = (int)(m->tensors["enc_q.pre.w"].numel / H);
    FETCH_F32(q.pre_w, "enc_q.pre.w", (int64_t)q.c_in * H);     FETCH_F32(q.pre_b, "enc_q.pre.b", H);
    FETCH_F32(q.proj_w, "enc_q.proj.w", H * 2 * H);              FETCH_F32(q.proj_b, "enc_q.proj.b", 2 * H);
    FETCH_F32(q.t_pre, "tf32.enc_q.pre.w", (int64_t)q.c_in * H); FETCH_F32(q.x_pre, "x3.enc_q.pre.w", (int64_t)2 * q.c_in * H);
    FETCH_F32(q.t_proj, "tf32.enc_q.proj.w", H * 2 * H);         FETCH_F32(q.x_proj, "x3.enc_q.proj.w", 2 * H * 2 * H);
    VS_TRY(resolve_wn(m, "enc_q.", 16, S, &q.wn));
  }
  VS_TRY(resolve_decoder(
      [m](const std::string& name, int64_t numel, int32_t dtype, const void** out) { return fetch(m, name, numel, dtype, out); },
      m->cfg.n_speakers, &m->dec));
  m->finalized = true;
  return VS_OK;
}

// ---- attentions.Encoder.forward (attentions.py:35-47), in place on x -------------------------------
static int64_t encoder_ws_floats(int R) {
  // qkv, att, y (4 K-slice partials), h (fp32, or hi + lo halves), planar hi / lo copies of x and att, attention scratch
  return (int64_t)R * (3 * kHidden + kHidden + 4 * kHidden + kFilter + 2 * kHidden) + attention_umma_ws_floats(R) + 16 * 64;
}

// Where the three-term fp16 conv (umma_split.cu) replaces 3xTF32: calls in the 3xTF32 regime (>= x3_min_rows rows, and not the
// plain-TF32 one), option "split16" (default 1).
static bool use_split16(int R, bool tf32_allowed) {
  if (!opts().v[OPT_SPLIT16] || R < opts().v[OPT_X3_MIN_ROWS]) return false;
  return !(tf32_allowed && R >= opts().v[OPT_TF32_MIN_ROWS]);
}

// one conv: fp32 rows in (converted to planar hi / lo in `scratch`) -> fp32 rows out
static int conv_rows_split(const ConvF32& c, const __half* w_s16, Workspace scratch, cudaStream_t st) {
  __half* hi = scratch.take<__half>((int64_t)c.R * c.Cin);
  __half* lo = scratch.take<__half>((int64_t)c.R * c.Cin);
  if (!scratch.ok) { set_error("conv_rows_split: workspace too small"); return VS_ERR_WORKSPACE; }
  VS_TRY(rows_to_split(c.in, c.in_ld, 0, 1, nullptr, hi, lo, c.R, c.Cin, st));
  UmmaSplit u;
  u.in_hi = hi; u.in_lo = lo; u.w = w_s16; u.bias = c.bias; u.out32 = c.out; u.out32_ld = c.out_ld; u.row_utt = c.row_utt;
  u.R = c.R; u.Cin = c.Cin; u.N = c.Cout; u.taps = c.k; u.dil = c.dil; u.pad_l = c.pad_l; u.act = c.act;
  return umma_split(u, st);
}

// attentions.Encoder.forward on the fp16 hi/lo tensor-core conv: x stays fp32 row-major (residual stream, LayerNorm), every
// conv operand is its planar hi / lo copy, written by the kernel that produced it where that kernel is ours to change
static int encoder_forward_split(const std::vector<EncLayer>& layers, const VsRows& rows, float* x, Workspace& ws, cudaStream_t st) {
  const int R = rows.n_rows, H = kHidden, F = kFilter;
  float* qkv = ws.take<float>((int64_t)R * 3 * H);
  float* att = ws.take<float>((int64_t)R * H);
  float* y = ws.take<float>((int64_t)4 * R * H);
  __half* xs_hi = ws.take<__half>((int64_t)R * H);
  __half* xs_lo = ws.take<__half>((int64_t)R * H);
  __half* as_hi = ws.take<__half>((int64_t)R * H);
  __half* as_lo = ws.take<__half>((int64_t)R * H);
  __half* h_hi = ws.take<__half>((int64_t)R * F);
  __half* h_lo = ws.take<__half>((int64_t)R * F);
  if (!ws.ok) { set_error("encoder: workspace too small"); return VS_ERR_WORKSPACE; }
  VS_TRY(rows_to_split(x, H, 0, 1, rows.row_utt, xs_hi, xs_lo, R, H, st));
  for (const EncLayer& L : layers) {
    UmmaSplit u;
    u.R = R; u.row_utt = rows.row_utt;
    u.in_hi = xs_hi; u.in_lo = xs_lo; u.Cin = H; u.w = L.s_wqkv; u.bias = L.bqkv; u.out32 = qkv; u.out32_ld = 3 * H; u.N = 3 * H;
    VS_TRY(umma_split(u, st));                                               // conv_q|k|v (attentions.py:139-141)
    bool planar = false;                                                     // the tcgen05 path writes the O conv's operand itself
    VS_TRY(rel_attention(rows, qkv, L.ek, L.ev, att, st, &ws, as_hi, as_lo, &planar, true));   // attentions.py:148-179
    if (!planar) VS_TRY(rows_to_split(att, H, 0, 1, rows.row_utt, as_hi, as_lo, R, H, st));   // masks gap rows itself
    u.in_hi = as_hi; u.in_lo = as_lo; u.w = L.s_wo; u.bias = L.bo; u.out32 = y; u.out32_ld = H; u.N = H;
    VS_TRY(umma_split(u, st));                                               // conv_o
    VS_TRY(layernorm_rows_ex(x, y, 1, 0, L.g1, L.b1, x, xs_hi, xs_lo, R, H, rows.row_utt, st));   // x = LN(x + y), + its hi / lo copy
    u = UmmaSplit();
    u.R = R; u.row_utt = rows.row_utt; u.taps = 3; u.pad_l = 1;
    u.in_hi = xs_hi; u.in_lo = xs_lo; u.Cin = H; u.w = L.s_w1; u.bias = L.bf1; u.act = 1; u.out_hi = h_hi; u.out_lo = h_lo; u.N = F;
    VS_TRY(umma_split(u, st));                                               // FFN conv_1 + relu -> planar hi / lo (attentions.py:278-282)
    u.in_hi = h_hi; u.in_lo = h_lo; u.Cin = F; u.w = L.s_w2; u.bias = L.bf2; u.act = 0; u.out_hi = nullptr; u.out_lo = nullptr;
    u.out32 = y; u.out32_ld = H; u.out32_slice = (int64_t)R * H; u.N = H; u.k_slices = F / H;
    VS_TRY(umma_split(u, st));                                               // FFN conv_2: four K-slices, fp32 partials
    VS_TRY(layernorm_rows_ex(x, y, F / H, (int64_t)R * H, L.g2, L.b2, x, xs_hi, xs_lo, R, H, rows.row_utt, st));
  }
  return VS_OK;
}

static int encoder_forward(const std::vector<EncLayer>& layers, const VsRows& rows, float* x, Workspace& ws,
                           cudaStream_t st, bool frame_level) {
  const int R = rows.n_rows, H = kHidden, F = kFilter;
  if (use_split16(R, frame_level)) return encoder_forward_split(layers, rows, x, ws, st);
  float* qkv = ws.take<float>((int64_t)R * 3 * H);
  float* att = ws.take<float>((int64_t)R * H);
  float* y = ws.take<float>((int64_t)R * H);
  float* hbuf = ws.take<float>((int64_t)R * F);
  if (!ws.ok) { set_error("encoder: workspace too small"); return VS_ERR_WORKSPACE; }
  for (const EncLayer& L : layers) {
    ConvF32 c;
    c.R = R; c.row_utt = rows.row_utt;
    c.in = x; c.in_ld = H; c.Cin = H; c.w = L.wqkv; c.bias = L.bqkv; c.out = qkv; c.out_ld = 3 * H; c.Cout = 3 * H;
    VS_TRY(conv_rows(c, frame_level ? L.t_wqkv : nullptr, L.x_wqkv, st));                                    // conv_q|k|v (attentions.py:139-141)
    VS_TRY(rel_attention(rows, qkv, L.ek, L.ev, att, st, &ws));            // attentions.py:148-179 (scratch: what is left of ws)
    c.in = att; c.w = L.wo; c.bias = L.bo; c.out = y; c.out_ld = H; c.Cout = H;
    VS_TRY(conv_rows(c, frame_level ? L.t_wo : nullptr, L.x_wo, st));                                      // conv_o
    VS_TRY(layernorm_rows(x, y, L.g1, L.b1, x, R, H, rows.row_utt, st));   // x = LN(x + y)
    c.in = x; c.Cin = H; c.w = L.w1; c.bias = L.bf1; c.out = hbuf; c.out_ld = F; c.Cout = F; c.k = 3; c.pad_l = 1; c.act = 1;
    VS_TRY(conv_rows(c, frame_level ? L.t_w1 : nullptr, L.x_w1, st));                                      // FFN conv_1 + relu (attentions.py:278-282)
    c.in = hbuf; c.in_ld = F; c.Cin = F; c.w = L.w2; c.bias = L.bf2; c.out = y; c.out_ld = H; c.Cout = H; c.act = 0;
    VS_TRY(conv_rows(c, frame_level ? L.t_w2 : nullptr, L.x_w2, st));                                      // FFN conv_2
    VS_TRY(layernorm_rows(x, y, L.g2, L.b2, x, R, H, rows.row_utt, st));
  }
  return VS_OK;
}

static int check_rows(const VsRows* r, const char* what) {
  VS_REQUIRE(r && r->n_utt > 0 && r->n_rows > 0 && r->row_utt && r->utt_start && r->utt_len && r->sid,
             "%s: incomplete VsRows", what);
  return VS_OK;
}

}  // namespace vs

using namespace vs;

extern "C" {

const char* vs_last_error(void) { return vs::last_error(); }
int vs_version(void) { return 1; }
int64_t vs_launch_count(void) { return (int64_t)vs::g_launch_count; }

int vs_set_option(const char* name, int64_t value) { return vs::option_set(nullptr, name, value); }

int vs_model_set_option(VsModel* m, const char* name, int64_t value) {
  VS_REQUIRE(m, "vs_model_set_option: null model");
  return vs::option_set(&m->overrides, name, value);
}

int vs_model_create(const VsConfig* cfg, VsModel** out) {
  VS_REQUIRE(cfg && out, "vs_model_create: null argument");
  VS_REQUIRE(cfg->hidden == kHidden && cfg->filter == kFilter && cfg->n_heads == kHeads && cfg->window == kWindow &&
                 cfg->hop == kHop && cfg->upsample_initial == 512 && cfg->gin == 256 && cfg->flow_layers <= 8,
             "vs_model_create: only the configs/config.json architecture is built (hidden 192, filter 768, 2 heads, "
             "window 4, hop 512)");
  int n_dev = 0;
  VS_CUDA_CHECK(cudaGetDeviceCount(&n_dev));
  VS_REQUIRE(n_dev > 0, "vs_model_create: no CUDA device; there is no CPU fallback");
  VsModel* m = new VsModel();
  m->cfg = *cfg;
  for (int i = 0; i < vs::OPT_COUNT; ++i) m->overrides.v[i] = vs::kOptUnset;
  *out = m;
  return VS_OK;
}

void vs_model_destroy(VsModel* m) { delete m; }

int vs_model_set_tensor(VsModel* m, const char* name, const void* ptr, int64_t numel, int32_t dtype) {
  VS_REQUIRE(m && name && ptr && numel > 0, "vs_model_set_tensor: bad argument");
  m->tensors[name] = vs::Tensor{ptr, numel, dtype};
  m->finalized = false;
  return VS_OK;
}

int vs_model_finalize(VsModel* m) {
  VS_REQUIRE(m, "vs_model_finalize: null model");
  return vs::finalize(m);
}

// Workspace sizes.  Latent stages (text encoder ... flow, posterior encoder) and the decoder are sized separately: the
// decoder's activation buffers dominate (fp32 cross-check decoder 5 x 16384 floats per frame row = 328 KB, the f16
// product decoder 6 x 16384 halves = 197 KB), the latent stages need ~10 KB per row - a side-stream latent workspace must
// not pay for a second set of decoder buffers.
int64_t vs_workspace_bytes_latent(const VsModel* m, int32_t rp, int32_t rf) {
  (void)m;
  const int64_t H = kHidden;
  const int64_t enc_p = encoder_ws_floats(rp), enc_f = encoder_ws_floats(rf);
  const int64_t variance = enc_p + (int64_t)rp * (H + 2 * 768 + 8 + 5 * 768);   // + the energy predictor's conv_2 operand and K-slice partials
  const int64_t prior = enc_f + (int64_t)rf * (2 * H + H);                   // stats + the projection's hi / lo operand copy
  const int64_t flow = (int64_t)rf * (H + 2 * H + H + 2 * H + H + 2 * H);   // also covers vs_posterior_encode
  int64_t mx = variance;
  if (prior > mx) mx = prior;
  if (flow > mx) mx = flow;
  return mx * 4 + (1 << 20);
}

int64_t vs_workspace_bytes_decoder(const VsModel* m, int32_t rf, int32_t precision) {
  if (precision == 1) return decoder_ws_floats(rf) * 4 + (1 << 20);
  vs::OptionScope option_scope(m ? &m->overrides : nullptr);
  const int64_t bufs = vs::decoder_two_streams(rf) ? 9 : 6;
  return (int64_t)rf * (bufs * 16384 * 2 + kHidden * 2 + 4) + (1 << 20);
}

/* any single call, either decoder precision */
int64_t vs_workspace_bytes(const VsModel* m, int32_t rp, int32_t rf) {
  const int64_t lat = vs_workspace_bytes_latent(m, rp, rf);
  int64_t dec = vs_workspace_bytes_decoder(m, rf, 1);
  const int64_t dec16 = vs_workspace_bytes_decoder(m, rf, 0);
  if (dec16 > dec) dec = dec16;
  return lat > dec ? lat : dec;
}

#define VS_ENTER(m, rows, what)                                                          \
  VS_REQUIRE((m) && (m)->finalized, what ": model not finalized");                       \
  VS_TRY(check_rows(rows, what));                                                        \
  vs::OptionScope option_scope(&(m)->overrides);                                         \
  cudaStream_t st = static_cast<cudaStream_t>(stream);                                   \
  Workspace W(ws, ws_bytes);

int vs_text_encode(const VsModel* m, const VsRows* rows, const int32_t* ids_rows, float* x_out, void* ws,
                   int64_t ws_bytes, void* stream) {
  VS_ENTER(m, rows, "vs_text_encode");
  VS_TRY(embed_rows(ids_rows, m->emb, x_out, rows->n_rows, m->cfg.n_vocab, st));
  return encoder_forward(m->enc_text, *rows, x_out, W, st, false);
}

int vs_variance_adapter(const VsModel* m, const VsRows* rows, float* x, int32_t dur_mode, float dur_scale,
                        const double* dur_ctrl, int32_t pitch_mode, float pitch_scale, const float* pitch_ctrl,
                        int32_t energy_mode, float energy_scale, const float* energy_ctrl, double* duration_out,
                        float* f0_out, float* energy_out, void* ws, int64_t ws_bytes, void* stream) {
  VS_ENTER(m, rows, "vs_variance_adapter");
  VS_REQUIRE((dur_mode == 0 || (dur_mode == 2 && dur_ctrl)) && (pitch_mode == 0 || (pitch_mode == 2 && pitch_ctrl)) &&
                 (energy_mode == 0 || (energy_mode == 2 && energy_ctrl)),
             "vs_variance_adapter: mode must be 0 (predict) or 2 (override, needs the control array)");
  const int R = rows->n_rows, H = kHidden;
  float* t = W.take<float>((int64_t)R * H);
  float* h1 = W.take<float>((int64_t)R * 768);
  float* h2 = W.take<float>((int64_t)R * 768);
  float* s0 = W.take<float>(R);
  float* s1 = W.take<float>(R);
  if (!W.ok) { set_error("vs_variance_adapter: workspace too small"); return VS_ERR_WORKSPACE; }
  ConvF32 c;

  // duration (models.py:681-688; DurationPredictor.forward :119-133) - uses x BEFORE the prenets touch it
  if (dur_mode == 0) {
    VS_TRY(add_speaker_rows(x, m->dp_cond, *rows, t, H, st));
    c = ConvF32(); c.R = R; c.row_utt = rows->row_utt; c.k = 3; c.pad_l = 1; c.act = 1;
    c.in = t; c.in_ld = H; c.Cin = H; c.w = m->dp_w1; c.bias = m->dp_b1; c.out = h1; c.out_ld = 256; c.Cout = 256;
    if (use_split16(R, false)) VS_TRY(conv_rows_split(c, m->s_dp_w1, W, st));
    else VS_TRY(conv_rows(c, nullptr, m->x_dp_w1, st));
    VS_TRY(layernorm_rows(h1, nullptr, m->dp_g1, m->dp_be1, h1, R, 256, rows->row_utt, st));
    c.in = h1; c.in_ld = 256; c.Cin = 256; c.w = m->dp_w2; c.bias = m->dp_b2; c.out = h2;
    VS_TRY(conv1d_f32(c, st));
    VS_TRY(layernorm_rows(h2, nullptr, m->dp_g2, m->dp_be2, h2, R, 256, rows->row_utt, st));
    VS_TRY(row_dot(h2, 256, m->dp_wp, m->dp_bp, s0, R, 256, rows->row_utt, st));
  }
  VS_TRY(duration_rows(s0, dur_ctrl, dur_mode, dur_scale, *rows, duration_out, st));

  // pitch (models.py:691-698; PitchPredictor.forward :505-514)
  if (pitch_mode == 0) {
    VS_TRY(add_speaker_rows(x, m->pp_cond, *rows, t, H, st));
    Workspace W2 = W;
    VS_TRY(encoder_forward(m->enc_pitch, *rows, t, W2, st, false));
    VS_TRY(row_dot(t, H, m->pp_wf0, m->pp_bf0, s0, R, H, rows->row_utt, st));
  }
  VS_TRY(pitch_rows(s0, pitch_ctrl, pitch_mode, pitch_scale, *rows, s1, f0_out, st));
  VS_TRY(prenet_add(x, s1, m->pitch_pre_w, m->pitch_pre_b, *rows, st));   // x += pitch_prenet(LF0)

  // energy (models.py:701-708; EnergyPredictor frame_prior_network.py:119-124) - sees x after the pitch prenet
  if (energy_mode == 0) {
    VS_TRY(add_speaker_rows(x, m->ep_cond, *rows, t, H, st));
    c = ConvF32(); c.R = R; c.row_utt = rows->row_utt; c.k = 3; c.pad_l = 1; c.act = 1;
    c.in = t; c.in_ld = H; c.Cin = H; c.w = m->ep_w1; c.bias = m->ep_b1; c.out = h1; c.out_ld = 768; c.Cout = 768;
    if (use_split16(R, false)) {
      // both convs on the three-term fp16 conv: LayerNorm 1 writes conv_2's planar hi / lo operand, conv_2 (Cin = 768) runs as four
      // K-slices with fp32 partials, and LayerNorm 2 sums them, applies the ReLU that had to wait for the sum, and normalises
      VS_TRY(conv_rows_split(c, m->s_ep_w1, W, st));
      Workspace W2 = W;
      __half* e_hi = W2.take<__half>((int64_t)R * 768);
      __half* e_lo = W2.take<__half>((int64_t)R * 768);
      float* part = W2.take<float>((int64_t)4 * R * 768);
      if (!W2.ok) { set_error("vs_variance_adapter: workspace too small"); return VS_ERR_WORKSPACE; }
      VS_TRY(layernorm_rows_ex(h1, nullptr, 1, 0, m->ep_g1, m->ep_be1, h1, e_hi, e_lo, R, 768, rows->row_utt, st));
      UmmaSplit u;
      u.R = R; u.row_utt = rows->row_utt; u.taps = 3; u.pad_l = 1; u.in_hi = e_hi; u.in_lo = e_lo; u.Cin = 768; u.w = m->s_ep_w2;
      u.bias = m->ep_b2; u.act = 0; u.out32 = part; u.out32_ld = 768; u.out32_slice = (int64_t)R * 768; u.N = 768; u.k_slices = 4;
      VS_TRY(umma_split(u, st));
      VS_TRY(layernorm_rows_ex(nullptr, part, 4, (int64_t)R * 768, m->ep_g2, m->ep_be2, h2, nullptr, nullptr, R, 768, rows->row_utt, st, 1));
    } else {
      VS_TRY(conv_rows(c, nullptr, m->x_ep_w1, st));
      VS_TRY(layernorm_rows(h1, nullptr, m->ep_g1, m->ep_be1, h1, R, 768, rows->row_utt, st));
      c.in = h1; c.in_ld = 768; c.Cin = 768; c.w = m->ep_w2; c.bias = m->ep_b2; c.out = h2;
      VS_TRY(conv_rows(c, nullptr, m->x_ep_w2, st));
      VS_TRY(layernorm_rows(h2, nullptr, m->ep_g2, m->ep_be2, h2, R, 768, rows->row_utt, st));
    }
    VS_TRY(row_dot(h2, 768, m->ep_wl, m->ep_bl, s0, R, 768, rows->row_utt, st));
  }
  VS_TRY(energy_rows(s0, energy_ctrl, energy_mode, energy_scale, *rows, s1, energy_out, st));
  VS_TRY(prenet_add(x, s1, m->energy_pre_w, m->energy_pre_b, *rows, st)); // x += energy_prenet(norm_energy)
  return VS_OK;
}

int vs_length_regulate_count(const VsRows* rows, const double* duration, int32_t* cum_out, int32_t* frames_out,
                             void* stream) {
  VS_TRY(check_rows(rows, "vs_length_regulate_count"));
  VS_REQUIRE(duration && cum_out && frames_out, "vs_length_regulate_count: null pointer");
  return lr_count(*rows, duration, cum_out, frames_out, static_cast<cudaStream_t>(stream));
}

int vs_length_regulate_gather(const VsRows* rows_p, const VsRows* rows_f, const float* x_p, const int32_t* cum,
                              float* x_f, int32_t* lr_index, void* stream) {
  VS_TRY(check_rows(rows_p, "vs_length_regulate_gather"));
  VS_TRY(check_rows(rows_f, "vs_length_regulate_gather"));
  VS_REQUIRE(rows_p->n_utt == rows_f->n_utt, "vs_length_regulate_gather: utterance counts differ");
  return lr_gather(*rows_p, *rows_f, x_p, cum, x_f, lr_index, static_cast<cudaStream_t>(stream));
}

int vs_frame_prior(const VsModel* m, const VsRows* rows, const float* x_f, const float* noise, uint64_t noise_seed,
                   float noise_scale, float* x_frame_out, float* m_p, float* logs_p, float* z_p, void* ws, int64_t ws_bytes,
                   void* stream) {
  VS_ENTER(m, rows, "vs_frame_prior");
  const int R = rows->n_rows, H = kHidden;
  float* stats = W.take<float>((int64_t)R * 2 * H);
  if (!W.ok) { set_error("vs_frame_prior: workspace too small"); return VS_ERR_WORKSPACE; }
  if (x_frame_out != x_f)
    VS_CUDA_CHECK(cudaMemcpyAsync(x_frame_out, x_f, sizeof(float) * (size_t)R * H, cudaMemcpyDeviceToDevice, st));
  const bool tf32_prior = opts().v[OPT_TF32_PRIOR] != 0;
  VS_TRY(encoder_forward(m->enc_prior, *rows, x_frame_out, W, st, tf32_prior));       // FramePriorNet.forward models.py:466-470
  ConvF32 c;                                                               // Projection.forward models.py:526-529
  c.R = R; c.row_utt = rows->row_utt; c.in = x_frame_out; c.in_ld = H; c.Cin = H; c.w = m->proj_w; c.bias = m->proj_b;
  c.out = stats; c.out_ld = 2 * H; c.Cout = 2 * H;
  if (use_split16(R, tf32_prior)) VS_TRY(conv_rows_split(c, m->s_proj_w, W, st));
  else VS_TRY(conv_rows(c, tf32_prior ? m->t_proj_w : nullptr, m->x_proj_w, st));
  return prior_sample(stats, noise, noise_seed, noise_scale, *rows, m_p, logs_p, z_p, st);
}

// WN.forward (modules.py:148-176): h is consumed (updated in place), skip receives the output (already masked).
struct WnBufs { float *a, *acts, *rs; };
static int wn_forward(const WnW& w, const VsRows& rows, float* h, float* skip, const WnBufs& b, cudaStream_t st) {
  const int R = rows.n_rows, H = kHidden, L = w.n_layers;
  const bool wn_tf32 = R >= opts().v[OPT_TF32_MIN_ROWS], wn_x3 = !wn_tf32 && R >= opts().v[OPT_X3_MIN_ROWS];
  const bool fused_wn = wn_tf32 || wn_x3;            // tensor-core path: gate and res/skip update live in the conv epilogues
  if (wn_tf32 && opts().v[OPT_WN_FUSED]) {                       // one kernel per layer on planar fp32 h / skip (umma_wn.cu)
    float* hb[2] = {b.a, b.a + (size_t)R * H};       // b.a and b.rs are [R][2H] scratch of the unfused path
    float* skip_pl = b.rs;
    VS_TRY(rows_to_planar4(h, hb[0], R, H, st));
    for (int l = 0; l < L; ++l) {
      UmmaWn u;
      u.h_in = hb[l & 1]; u.h_out = hb[(l + 1) & 1]; u.skip = skip_pl;
      u.w_in = w.t_in_gate[l]; u.b_in = w.in_gate_b[l];
      u.cond = w.cond_tab_gate + (size_t)2 * H * l; u.cond_ld = 2 * H * L; u.cond_idx = rows.sid;
      u.w_rs = w.t_rs[l]; u.b_rs = w.rs_b[l]; u.row_utt = rows.row_utt; u.R = R; u.first = (l == 0); u.last = (l == L - 1);
      VS_TRY(umma_wn_layer(u, st));
    }
    return planar4_to_rows(skip_pl, skip, R, H, st);
  }
  for (int l = 0; l < L; ++l) {
    const int rsC = (l < L - 1) ? 2 * H : H;
    if (fused_wn) {
      UmmaTf32 u;                                                          // acts = tanh . sigmoid (in_layer(h) + g_l)
      u.in = h; u.in_ld = H; u.w = wn_tf32 ? w.t_in_gate[l] : w.x_in_gate[l]; u.split3 = wn_x3 ? 1 : 0; u.bias = w.in_gate_b[l];
      u.ubias = w.cond_tab_gate + (size_t)2 * H * l; u.ubias_ld = 2 * H * L; u.ubias_idx = rows.sid;
      u.out = b.acts; u.out_ld = H; u.row_utt = rows.row_utt; u.R = R; u.Cin = H; u.N = 2 * H; u.taps = 5; u.pad_l = 2;
      u.epi = 1;
      VS_TRY(umma_tf32(u, st));
      u = UmmaTf32();                                                      // h += rs[:, :H]; skip (+)= rs[:, H:]
      u.in = b.acts; u.in_ld = H; u.w = wn_tf32 ? w.t_rs[l] : w.x_rs[l]; u.split3 = wn_x3 ? 1 : 0; u.bias = w.rs_b[l];
      u.out = h; u.out_ld = H; u.out2 = skip; u.out2_ld = H;
      u.nb_split = (l < L - 1) ? 1 : 0; u.accumulate2 = (l > 0); u.row_utt = rows.row_utt; u.R = R; u.Cin = H; u.N = rsC;
      u.epi = 2;
      VS_TRY(umma_tf32(u, st));
      continue;
    }
    ConvF32 c;
    c.R = R;
    c.in = h; c.in_ld = H; c.Cin = H; c.w = w.in_w[l]; c.bias = w.in_b[l]; c.out = b.a; c.out_ld = 2 * H; c.Cout = 2 * H;
    c.k = 5; c.pad_l = 2;
    VS_TRY(conv_rows(c, w.t_in[l], w.x_in[l], st));
    VS_TRY(wn_gate(b.a, w.cond_tab, 2 * H * L, 2 * H * l, rows, b.acts, st));
    c = ConvF32(); c.R = R;
    c.in = b.acts; c.in_ld = H; c.Cin = H; c.w = w.rs_w[l]; c.bias = w.rs_b[l]; c.out = b.rs; c.out_ld = rsC; c.Cout = rsC;
    VS_TRY(conv_rows(c, w.t_rs[l], w.x_rs[l], st));
    VS_TRY(wn_update(b.rs, rsC, l == L - 1, l == 0, rows, h, skip, st));
  }
  return VS_OK;
}

// ResidualCouplingBlock (models.py:177-209), mean-only layers (modules.py:324-343), in place on z.
// reverse: Flip,RCL3,Flip,RCL2,Flip,RCL1,Flip,RCL0 with x1 -= m;  forward: RCL0,Flip,...,RCL3,Flip with x1 += m.
// The Flips are folded into packed weights: layer f runs on the physically un-flipped tensor with `flipped = f odd`
// in BOTH directions (packing.py).
static int flow_run(const VsModel* m, const VsRows* rows, float* z, bool reverse, Workspace& W, cudaStream_t st) {
  const int R = rows->n_rows, H = kHidden, F = m->cfg.n_flows;
  float* h = W.take<float>((int64_t)R * H);
  WnBufs b;
  b.a = W.take<float>((int64_t)R * 2 * H);
  b.acts = W.take<float>((int64_t)R * H);
  b.rs = W.take<float>((int64_t)R * 2 * H);
  float* skip = W.take<float>((int64_t)R * H);
  float* mm = W.take<float>((int64_t)R * (H / 2));
  if (!W.ok) { set_error("flow: workspace too small"); return VS_ERR_WORKSPACE; }
  for (int i = 0; i < F; ++i) {
    const int f = reverse ? F - 1 - i : i;
    const FlowW& w = m->flows[f];
    const bool flipped = (f & 1) != 0;
    const int in_off = flipped ? H / 2 : 0, upd_off = flipped ? 0 : H / 2;
    if (w.c16_w && opts().v[OPT_COUPLING_FUSED] && R >= opts().v[OPT_COUPLING_MIN_ROWS] && (R >= opts().v[OPT_TF32_MIN_ROWS] || opts().v[OPT_COUPLING_MIN_ROWS] < 4096)) {
      // the whole coupling layer as ONE kernel, residual stream and skip sum in fp32 in TMEM, fp16 operands (the 11-bit significand
      // of the TF32 regime it replaced for large calls; small calls take it too by default, option coupling_min_rows)  modules.py:324-343, 148-176
      UmmaCoupling u;
      u.z = z; u.w = w.c16_w; u.bias = w.c16_b; u.row_utt = rows->row_utt; u.sid = rows->sid; u.R = R;
      u.in_off = in_off; u.upd_off = upd_off; u.sign = reverse ? -1.f : 1.f;
      VS_TRY(umma_coupling(u, st));
      continue;
    }
    ConvF32 c;
    c.R = R; c.row_utt = rows->row_utt;
    c.in = z + in_off; c.in_ld = H; c.Cin = H / 2; c.w = w.pre_w; c.bias = w.pre_b; c.out = h; c.out_ld = H; c.Cout = H;
    VS_TRY(conv_rows(c, w.t_pre, w.x_pre, st));                            // h = pre(x0) * mask  (modules.py:326)
    VS_TRY(wn_forward(w.wn, *rows, h, skip, b, st));
    c = ConvF32(); c.R = R;
    c.in = skip; c.in_ld = H; c.Cin = H; c.w = w.post_w; c.bias = w.post_b; c.out = mm; c.out_ld = H / 2; c.Cout = H / 2;
    VS_TRY(conv_rows(c, w.t_post, w.x_post, st));                          // m = post(h) (modules.py:328)
    VS_TRY(coupling_update(z, upd_off, mm, reverse ? -1.f : 1.f, *rows, st));   // x1 = (x1 -/+ m) * mask (modules.py:336,341)
  }
  return VS_OK;
}

int vs_flow_reverse(const VsModel* m, const VsRows* rows, float* z, void* ws, int64_t ws_bytes, void* stream) {
  VS_ENTER(m, rows, "vs_flow_reverse");
  return flow_run(m, rows, z, true, W, st);
}

int vs_flow_forward(const VsModel* m, const VsRows* rows, float* z, void* ws, int64_t ws_bytes, void* stream) {
  VS_ENTER(m, rows, "vs_flow_forward");
  return flow_run(m, rows, z, false, W, st);
}

int vs_posterior_encode(const VsModel* m, const VsRows* rows, const float* spec, const float* noise, uint64_t noise_seed,
                        float* z, float* m_q, float* logs_q, void* ws, int64_t ws_bytes, void* stream) {
  VS_ENTER(m, rows, "vs_posterior_encode");
  if (!m->enc_q.present) { set_error("vs_posterior_encode: enc_q.* weights were not registered"); return VS_ERR_MISSING; }
  const PosteriorW& q = m->enc_q;
  const int R = rows->n_rows, H = kHidden;
  float* h = W.take<float>((int64_t)R * H);
  WnBufs b;
  b.a = W.take<float>((int64_t)R * 2 * H);
  b.acts = W.take<float>((int64_t)R * H);
  b.rs = W.take<float>((int64_t)R * 2 * H);
  float* skip = W.take<float>((int64_t)R * H);
  float* stats = W.take<float>((int64_t)R * 2 * H);
  if (!W.ok) { set_error("vs_posterior_encode: workspace too small"); return VS_ERR_WORKSPACE; }
  ConvF32 c;
  c.R = R; c.row_utt = rows->row_utt; c.in = spec; c.in_ld = q.c_in; c.Cin = q.c_in; c.w = q.pre_w; c.bias = q.pre_b;
  c.out = h; c.out_ld = H; c.Cout = H;
  VS_TRY(conv_rows(c, q.t_pre, q.x_pre, st));                              // x = pre(x) * x_mask  (models.py:235)
  VS_TRY(wn_forward(q.wn, *rows, h, skip, b, st));                         // enc (16-layer WN)  (models.py:236)
  c = ConvF32(); c.R = R; c.row_utt = rows->row_utt; c.in = skip; c.in_ld = H; c.Cin = H; c.w = q.proj_w; c.bias = q.proj_b;
  c.out = stats; c.out_ld = 2 * H; c.Cout = 2 * H;
  VS_TRY(conv_rows(c, q.t_proj, q.x_proj, st));                            // stats = proj(x) * x_mask  (models.py:237)
  return prior_sample(stats, noise, noise_seed, 1.f, *rows, m_q, logs_q, z, st);       // z = (m + eps*exp(logs)) * x_mask  (:239)
}

int vs_hifigan_decode(const VsModel* m, const VsRows* rows, const float* z, int32_t max_len, float* wave_out,
                      int32_t precision, void* ws, int64_t ws_bytes, void* stream) {
  VS_ENTER(m, rows, "vs_hifigan_decode");
  VS_REQUIRE(precision == 0 || precision == 1, "vs_hifigan_decode: precision must be 0 (f16 tcgen05) or 1 (fp32 check)");
  if (precision == 1) return decode_f32(m->dec, *rows, z, max_len, wave_out, W, st);
  return decode_f16(m->dec, *rows, z, max_len, wave_out, W, st);
}

int vs_randn(float* out, int64_t n, uint64_t seed, void* stream) {
  return randn_fill(out, n, seed, static_cast<cudaStream_t>(stream));
}

int vs_unpack_rows(const VsRows* rows, const float* x, int32_t C, int32_t rows_mul, int32_t t_max, float* out,
                   void* stream) {
  VS_TRY(check_rows(rows, "vs_unpack_rows"));
  return unpack_rows(*rows, x, C, rows_mul, t_max, out, static_cast<cudaStream_t>(stream));
}

int vs_wave_pcm16(const float* wave, int32_t n_utt, int32_t t_max, const int32_t* n_samples, int32_t decimate,
                  const float* fir, int32_t n_taps, int16_t* out, int32_t t_out, void* stream) {
  return pcm16(wave, n_utt, t_max, n_samples, decimate, fir, n_taps, out, t_out, static_cast<cudaStream_t>(stream));
}

int vs_op_conv1d_f32(const float* in, int32_t in_ld, const float* w, const float* bias, float* out, int32_t out_ld,
                     int32_t n_rows, int32_t c_in, int32_t c_out, int32_t k, int32_t dil, int32_t pad_l, float in_slope,
                     int32_t act, const int32_t* row_utt, int32_t row_div, void* stream) {
  ConvF32 c;
  c.in = in; c.in_ld = in_ld; c.w = w; c.bias = bias; c.out = out; c.out_ld = out_ld; c.R = n_rows; c.Cin = c_in;
  c.Cout = c_out; c.k = k; c.dil = dil; c.pad_l = pad_l; c.in_slope = in_slope; c.act = act; c.row_utt = row_utt;
  c.row_div = row_div > 0 ? row_div : 1;
  return conv1d_f32(c, static_cast<cudaStream_t>(stream));
}

int vs_op_layernorm(const float* a, const float* b, const float* gamma, const float* beta, float* out, int32_t n_rows,
                    int32_t C, const int32_t* row_utt, void* stream) {
  return layernorm_rows(a, b, gamma, beta, out, n_rows, C, row_utt, static_cast<cudaStream_t>(stream));
}

int vs_op_rel_attention(const VsRows* rows, const float* qkv, const float* emb_rel_k, const float* emb_rel_v,
                        float* out, void* ws, int64_t ws_bytes, void* stream) {
  VS_TRY(check_rows(rows, "vs_op_rel_attention"));
  vs::Workspace w(ws, ws_bytes);
  return rel_attention(*rows, qkv, emb_rel_k, emb_rel_v, out, static_cast<cudaStream_t>(stream), ws ? &w : nullptr);
}

int vs_op_conv1d_tf32(const float* in, int32_t in_ld, const float* w_packed, const float* bias, float* out, int32_t out_ld,
                      int32_t n_rows, int32_t c_in, int32_t c_out, int32_t taps, int32_t dil, int32_t pad_l, int32_t act,
                      int32_t split3, const int32_t* row_utt, void* stream) {
  UmmaTf32 u;
  u.in = in; u.in_ld = in_ld; u.w = w_packed; u.bias = bias; u.out = out; u.out_ld = out_ld; u.row_utt = row_utt;
  u.R = n_rows; u.Cin = c_in; u.N = c_out; u.taps = taps; u.dil = dil; u.pad_l = pad_l; u.act = act; u.split3 = split3;
  return umma_tf32(u, static_cast<cudaStream_t>(stream));
}

// op-level hook for the fp16 hi/lo conv (csrc/umma_split.cu): fp32 rows in, fp32 rows out; the split / partial-sum kernels around it
// are the ones the model uses
int vs_op_conv1d_split(const float* in, int32_t in_ld, const void* w_packed, const float* bias, float* out, int32_t out_ld,
                       int32_t n_rows, int32_t c_in, int32_t c_out, int32_t taps, int32_t dil, int32_t pad_l, int32_t act,
                       const int32_t* row_utt, void* ws, int64_t ws_bytes, void* stream) {
  VS_REQUIRE(in && w_packed && out && ws && n_rows > 0 && out_ld == c_out, "vs_op_conv1d_split: bad arguments (out must be dense)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  vs::Workspace w(ws, ws_bytes);
  const int slices = c_in <= 192 ? 1 : c_in / 192;
  __half* hi = w.take<__half>((int64_t)n_rows * c_in);
  __half* lo = w.take<__half>((int64_t)n_rows * c_in);
  float* part = slices > 1 ? w.take<float>((int64_t)slices * n_rows * c_out) : nullptr;
  if (!w.ok) { vs::set_error("vs_op_conv1d_split: workspace too small"); return VS_ERR_WORKSPACE; }
  VS_TRY(vs::rows_to_split(in, in_ld, 0, 1, nullptr, hi, lo, n_rows, c_in, st));
  vs::UmmaSplit u;
  u.in_hi = hi; u.in_lo = lo; u.w = static_cast<const __half*>(w_packed); u.bias = bias; u.row_utt = row_utt;
  u.out32 = slices > 1 ? part : out; u.out32_ld = out_ld; u.out32_slice = (int64_t)n_rows * c_out;
  u.R = n_rows; u.Cin = c_in; u.N = c_out; u.taps = taps; u.dil = dil; u.pad_l = pad_l; u.act = act; u.k_slices = slices;
  VS_TRY(vs::umma_split(u, st));
  if (slices > 1) VS_TRY(vs::sum_partials(part, (int64_t)n_rows * c_out, slices, out, (int64_t)n_rows * c_out, st));
  return VS_OK;
}

// 8(f) rank 4: spectrogram_torch / mel_spectrogram_torch (reference mel_processing.py:50-112) as frame rows -> 3xTF32 DFT
// GEMM (tcgen05) -> magnitude -> 3xTF32 mel GEMM -> log.  rows: n_frames[b] + 3 rows per utterance (vispeech_b200/mel.py).
int vs_mel_spectrogram(const VsRows* rows, const float* wave, int32_t t_max, const int32_t* n_samples, int32_t hop,
                       int32_t n_bins, int32_t n_mels, const float* dft_packed, const float* mel_packed, int32_t frames_max,
                       float* spec_out, float* mel_out, void* ws, int64_t ws_bytes, void* stream) {
  VS_TRY(check_rows(rows, "vs_mel_spectrogram"));
  VS_REQUIRE(wave && n_samples && dft_packed && (spec_out || mel_out), "vs_mel_spectrogram: null pointer");
  VS_REQUIRE(hop > 0 && hop % 4 == 0 && n_bins == 2 * hop + 1 && n_mels > 0 && n_mels <= 96 && frames_max > 0,
             "vs_mel_spectrogram: needs n_fft = 4 * hop, n_bins = n_fft / 2 + 1, n_mels <= 96");
  VS_REQUIRE(!mel_out || mel_packed, "vs_mel_spectrogram: mel basis missing");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int R = rows->n_rows;
  const int ld_x = (hop + 47) / 48 * 48;                  // K of the DFT GEMM, a multiple of the 3xTF32 stage (48)
  const int half = (n_bins + 191) / 192 * 192;            // re | im column blocks, each a multiple of the 192-column n-block
  const int ld_mag = (n_bins + 47) / 48 * 48;
  vs::Workspace w(ws, ws_bytes);
  float* x = w.take<float>((int64_t)R * ld_x);
  float* dft = w.take<float>((int64_t)R * 2 * half);
  float* mag = w.take<float>((int64_t)R * ld_mag);
  float* melv = w.take<float>((int64_t)R * 96);
  if (!w.ok) { vs::set_error("vs_mel_spectrogram: workspace too small (%lld bytes given)", (long long)ws_bytes); return VS_ERR_WORKSPACE; }
  VS_TRY(vs::mel_frame_rows(*rows, wave, t_max, n_samples, hop, ld_x, (4 * hop - hop) / 2, x, st));
  vs::UmmaTf32 u;
  u.in = x; u.in_ld = ld_x; u.w = dft_packed; u.out = dft; u.out_ld = 2 * half; u.R = R; u.Cin = ld_x; u.N = 2 * half;
  u.taps = 4; u.dil = 1; u.pad_l = 0; u.split3 = 1;
  VS_TRY(vs::umma_tf32(u, st));
  VS_TRY(vs::mel_magnitude(dft, 2 * half, half, n_bins, ld_mag, R, mag, st));
  if (spec_out) VS_TRY(vs::mel_unpack(*rows, mag, ld_mag, n_bins, frames_max, 0, spec_out, st));
  if (mel_out) {
    vs::UmmaTf32 m;
    m.in = mag; m.in_ld = ld_mag; m.w = mel_packed; m.out = melv; m.out_ld = 96; m.R = R; m.Cin = ld_mag; m.N = 96;
    m.taps = 1; m.split3 = 1;
    VS_TRY(vs::umma_tf32(m, st));
    VS_TRY(vs::mel_unpack(*rows, melv, 96, n_mels, frames_max, 1, mel_out, st));
  }
  return VS_OK;
}

// op-level hook for the one-kernel WN layer (csrc/umma_wn.cu): row-major in / out, planar inside, like wn_forward uses it
int vs_op_wn_layer(const float* h_in, const float* w_in_packed, const float* b_in, const float* cond, int32_t cond_ld,
                   const int32_t* cond_idx, const float* w_rs_packed, const float* b_rs, const int32_t* row_utt,
                   int32_t n_rows, int32_t first, int32_t last, float* h_out, float* skip, void* ws, int64_t ws_bytes,
                   void* stream) {
  VS_REQUIRE(h_in && skip && ws && n_rows > 0, "vs_op_wn_layer: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  vs::Workspace w(ws, ws_bytes);
  const int64_t n = (int64_t)n_rows * vs::kHidden;
  float *hp = w.take<float>(n), *ho = w.take<float>(n), *sp = w.take<float>(n);
  if (!w.ok) { vs::set_error("vs_op_wn_layer: workspace too small"); return VS_ERR_WORKSPACE; }
  VS_TRY(vs::rows_to_planar4(h_in, hp, n_rows, vs::kHidden, st));
  if (!first) VS_TRY(vs::rows_to_planar4(skip, sp, n_rows, vs::kHidden, st));
  vs::UmmaWn u;
  u.h_in = hp; u.h_out = ho; u.skip = sp; u.w_in = w_in_packed; u.b_in = b_in; u.cond = cond; u.cond_ld = cond_ld;
  u.cond_idx = cond_idx; u.w_rs = w_rs_packed; u.b_rs = b_rs; u.row_utt = row_utt; u.R = n_rows; u.first = first; u.last = last;
  VS_TRY(vs::umma_wn_layer(u, st));
  if (!last) { VS_REQUIRE(h_out, "vs_op_wn_layer: h_out missing"); VS_TRY(vs::planar4_to_rows(ho, h_out, n_rows, vs::kHidden, st)); }
  return vs::planar4_to_rows(sp, skip, n_rows, vs::kHidden, st);
}

int vs_op_conv1d_umma(const void* in_planar, const void* w_packed, const float* bias, const void* res_planar,
                      void* out_raw, void* out_act, int32_t n_rows, int32_t c_in, int32_t n_cols, int32_t taps,
                      int32_t dil, int32_t pad_l, int32_t up, float act_slope, float act_scale,
                      const int32_t* row_utt, int32_t row_div, void* stream) {
  UmmaConv c;
  c.in = static_cast<const __half*>(in_planar); c.w = static_cast<const __half*>(w_packed);
  c.bias = bias; c.res = static_cast<const __half*>(res_planar);
  c.out_raw = static_cast<__half*>(out_raw); c.out_act = static_cast<__half*>(out_act);
  c.R = n_rows; c.Cin = c_in; c.N = n_cols; c.taps = taps; c.dil = dil; c.pad_l = pad_l; c.up = up;
  c.act_slope = act_slope; c.act_scale = act_scale; c.row_utt = row_utt; c.row_div = row_div > 0 ? row_div : 1;
  return umma_conv1d(c, static_cast<cudaStream_t>(stream));
}

int vs_op_conv1d_umma2(const void* in_planar, const void* w_packed, const float* bias, const void* res_planar, const void* res2_planar,
                       float res_inv_slope, void* out_raw, void* out_act, int32_t n_rows, int32_t c_in, int32_t n_cols, int32_t taps,
                       int32_t dil, int32_t pad_l, int32_t up, float act_slope, float act_scale,
                       const int32_t* row_utt, int32_t row_div, void* stream) {
  UmmaConv c;
  c.in = static_cast<const __half*>(in_planar); c.w = static_cast<const __half*>(w_packed);
  c.bias = bias; c.res = static_cast<const __half*>(res_planar); c.res2 = static_cast<const __half*>(res2_planar);
  c.res_inv_slope = res_inv_slope;
  c.out_raw = static_cast<__half*>(out_raw); c.out_act = static_cast<__half*>(out_act);
  c.R = n_rows; c.Cin = c_in; c.N = n_cols; c.taps = taps; c.dil = dil; c.pad_l = pad_l; c.up = up;
  c.act_slope = act_slope; c.act_scale = act_scale; c.row_utt = row_utt; c.row_div = row_div > 0 ? row_div : 1;
  return umma_conv1d(c, static_cast<cudaStream_t>(stream));
}

int vs_op_respair(const void* x_planar, const void* w1_packed, const void* w2_packed, const float* b1, const float* b2,
                  const void* res2_planar, void* out_raw, void* out_act, int32_t n_rows, int32_t channels, int32_t taps,
                  int32_t dil, float act_slope, float act_scale, const int32_t* row_utt, int32_t row_div, void* stream) {
  UmmaPair c;
  c.x = static_cast<const __half*>(x_planar); c.w1 = static_cast<const __half*>(w1_packed);
  c.w2 = static_cast<const __half*>(w2_packed); c.b1 = b1; c.b2 = b2;
  c.res2 = static_cast<const __half*>(res2_planar); c.out_raw = static_cast<__half*>(out_raw);
  c.out_act = static_cast<__half*>(out_act); c.R = n_rows; c.C = channels; c.taps = taps; c.dil = dil;
  c.act_slope = act_slope; c.act_scale = act_scale; c.row_utt = row_utt; c.row_div = row_div > 0 ? row_div : 1;
  return umma_respair(c, static_cast<cudaStream_t>(stream));
}

int vs_op_resblock64(const void* a_planar, const void* const* w_packed /*[6]*/, const float* const* b_host /*[6]*/,
                     const int32_t* row_utt, int32_t row_div, int32_t n_rows, void* out_raw, void* stream) {
  VS_REQUIRE(w_packed && b_host, "vs_op_resblock64: null pointer");
  UmmaResBlock f;
  f.a = static_cast<const __half*>(a_planar); f.out_raw = static_cast<__half*>(out_raw);
  for (int m = 0; m < 3; ++m) {
    f.w[m][0] = static_cast<const __half*>(w_packed[2 * m]); f.w[m][1] = static_cast<const __half*>(w_packed[2 * m + 1]);
    f.b1_host[m] = b_host[2 * m]; f.b2_host[m] = b_host[2 * m + 1];
  }
  f.row_utt = row_utt; f.row_div = row_div > 0 ? row_div : 1; f.R = n_rows;
  return umma_resblock(f, static_cast<cudaStream_t>(stream));
}

int vs_op_mrf32(const void* x_hi, const void* x_lo, const void* const* w_packed /*[18]*/, const float* const* b_host /*[18]*/,
                const float* post_w_host, const int32_t* row_utt, int32_t row_div, int32_t n_rows, float* wave, void* stream) {
  VS_REQUIRE(w_packed && b_host, "vs_op_mrf32: null pointer");
  UmmaMrf f;
  f.x_hi = static_cast<const __half*>(x_hi); f.x_lo = static_cast<const __half*>(x_lo);
  for (int j = 0; j < 3; ++j)
    for (int m = 0; m < 3; ++m) {
      const int q = (j * 3 + m) * 2;
      f.w[j][m][0] = static_cast<const __half*>(w_packed[q]); f.w[j][m][1] = static_cast<const __half*>(w_packed[q + 1]);
      f.b1_host[j][m] = b_host[q]; f.b2_host[j][m] = b_host[q + 1];
    }
  f.post_w_host = post_w_host; f.row_utt = row_utt; f.row_div = row_div > 0 ? row_div : 1; f.R = n_rows; f.wave = wave;
  return umma_mrf(f, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
