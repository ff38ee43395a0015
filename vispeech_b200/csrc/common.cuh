// Shared helpers for the vispeech_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/vispeech_b200.h"

namespace vs {

void set_error(const char* fmt, ...);

#define VS_CUDA_CHECK(expr)                                                                     \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      vs::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));       \
      return VS_ERR_CUDA;                                                                       \
    }                                                                                           \
  } while (0)

// every kernel launch site ends with this: counts launches (vs_launch_count) and surfaces launch errors
extern unsigned long long g_launch_count;
#define VS_LAUNCH_CHECK()                      \
  do {                                         \
    ++vs::g_launch_count;                      \
    VS_CUDA_CHECK(cudaGetLastError());         \
  } while (0)

#define VS_REQUIRE(cond, ...)                                                                   \
  do {                                                                                          \
    if (!(cond)) {                                                                              \
      vs::set_error(__VA_ARGS__);                                                               \
      return VS_ERR_INVALID;                                                                    \
    }                                                                                           \
  } while (0)

#define VS_TRY(expr)                                                                            \
  do {                                                                                          \
    int _s = (expr);                                                                            \
    if (_s != VS_OK) return _s;                                                                 \
  } while (0)

constexpr int kHidden = 192;
constexpr int kFilter = 768;
constexpr int kHeads = 2;
constexpr int kHeadDim = 96;
constexpr int kWindow = 4;
constexpr int kRel = 2 * kWindow + 1;
constexpr int kHop = 512;

// Generic fp32 conv1d over ragged rows (implicit GEMM on CUDA cores).  See ops_simt.cu.
struct ConvF32 {
  const float* in = nullptr; int in_ld = 0;         // [R][in_ld], channel offset folded into the pointer
  const float* w = nullptr;                          // [k][Cin][Cout]
  const float* bias = nullptr;                       // [Cout] or null
  float* out = nullptr; int out_ld = 0;
  const float* res = nullptr; int res_ld = 0;        // optional residual, indexed like out
  const float* ubias = nullptr; int ubias_ld = 0;    // optional per-speaker bias table [n_spk][ubias_ld]
  const int32_t* ubias_idx = nullptr;                // [n_utt] row of ubias per utterance (sid)
  const int32_t* row_utt = nullptr; int row_div = 1; // validity of OUTPUT row r: row_utt[out_row / row_div] >= 0
  int R = 0, Cin = 0, Cout = 0, k = 1, dil = 1, pad_l = 0;
  int out_row_mul = 1, out_row_off = 0;              // output row = r*mul + off (polyphase ConvTranspose1d)
  int R_out = 0;                                     // rows in out (bounds), 0 -> R
  float in_slope = 1.f;                              // leaky-relu slope applied to the input (1 = identity)
  float out_scale = 1.f;
  int act = 0;                                       // 0 none, 1 relu, 2 tanh
  int accumulate = 0;                                // out += y
};
int conv1d_f32(const ConvF32& a, cudaStream_t st);

int layernorm_rows(const float* a, const float* b, const float* gamma, const float* beta, float* out, int R, int C,
                   const int32_t* row_utt, cudaStream_t st);
struct Workspace;
// `ws`: scratch for the tcgen05 kernel (attention_umma.cu, attention_umma_ws_floats(n_rows) floats); without it the
// frame-level path falls back to the mma.sync kernel
// LN(a + sum of n_b partials of b) with an optional second output: the row as planar fp16 hi / lo (umma_split.cu's operand);
// a may be null; pre_relu: LN(relu(a + sum b)) (a conv whose ReLU has to wait for the sum of its K-slice partials)
int layernorm_rows_ex(const float* a, const float* b, int n_b, int64_t b_stride, const float* gamma, const float* beta, float* out,
                      __half* out_hi, __half* out_lo, int R, int C, const int32_t* row_utt, cudaStream_t st, int pre_relu = 0);
// `out_hi` / `out_lo` (optional, with `wrote_planar`): where the tcgen05 path may write the result as planar fp16 hi / lo INSTEAD
// of fp32 rows (the O conv's operand); *wrote_planar tells the caller which form it got
int rel_attention(const VsRows& rows, const float* qkv, const float* ek, const float* ev, float* out, cudaStream_t st,
                  Workspace* ws = nullptr, __half* out_hi = nullptr, __half* out_lo = nullptr, bool* wrote_planar = nullptr,
                  bool gaps_dont_care = false);   // true: the caller masks gap rows itself (no memset of `out` on the CUDA-core path)
int rel_attention_umma(const VsRows& rows, const float* qkv, const float* ek, const float* ev, float* out, Workspace& ws,
                       cudaStream_t st, __half* out_hi = nullptr, __half* out_lo = nullptr);
int64_t attention_umma_ws_floats(int n_rows);
bool rel_attention_umma_fits(const VsRows& rows, const Workspace& ws);   // enough scratch left for this batch's tiles?
// tensor-core (3xTF32 mma.sync) form of the same op, attention_mma.cu; rel_attention() dispatches on opts() "attention_mma"
int rel_attention_mma(const VsRows& rows, const float* qkv, const float* ek, const float* ev, float* out, cudaStream_t st);
int row_dot(const float* x, int ld, const float* w, const float* bias, float* out, int R, int C,
            const int32_t* row_utt, cudaStream_t st);

// Per-device launch configuration.  cudaFuncSetAttribute applies to the device that is current when it is called and a
// process may hold models on several devices (and call from several threads: serving.py's worker), so the "done once"
// state is kept per (kernel, device) behind a mutex instead of in function-local statics.
// Programmatic dependent launch: a kernel launched with launch_pdl() may be scheduled while its predecessor in the stream is still
// draining - launch latency, CTA scheduling and the kernel's own prologue (barrier init, TMEM allocation) overlap the predecessor's
// tail.  Protocol, in EVERY kernel launched this way: pdl_trigger() first thing (lets the next kernel in), pdl_wait() before the
// first access to global memory another kernel writes or reads (it returns once every preceding grid has completed and flushed).
// Option "pdl" is a bit mask of kernel groups launched this way (0 = everything fully serialised): 1 the three-term conv, LayerNorm and
// rows_to_split, 32 the small element-wise kernels of the latent stages (default 33: text encoder 0.54 -> 0.47 ms, variance adapter 0.99 ->
// 0.88 at the C2 size, where ~100 kernels of 5 - 20 us on 20 - 90 CTAs run back to back), 16 the CUDA-core attention and row_dot (measured:
// cancels that gain), 2 / 64 / 128 the frame-level attention kernel, its tile re-layout and its band fix-up (64 alone: frame prior 2.19 ->
// 2.49 ms; the other two +-0), 4 the decoder (+-0 at batch size, on for small calls), 8 the flow (+-0).
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
bool pdl_enabled(int group);
struct PdlExtra {                 // for the lifetime of the object (this thread): also launch these groups' kernels that way
  explicit PdlExtra(int groups);
  ~PdlExtra();
  int saved;
};
template <int GROUP = 1, typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled(GROUP) ? 1 : 0;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif
int device_sm_count(int* n_sm);                            // SM count of the CURRENT device
int ensure_dynamic_smem(const void* kernel, int bytes);    // opt in to `bytes` of dynamic shared memory, once per device

// Runtime options (include/vispeech_b200.h lists them).  Two layers: process-wide defaults (vs_set_option) and per-model
// overrides (vs_model_set_option).  Every C-ABI entry point that takes a model opens an OptionScope, which freezes
// "defaults overlaid with that model's overrides" into a thread-local snapshot for the duration of the call: kernels'
// host code reads opts() and never a mutable global, so two models (or two threads) cannot see each other's settings.
enum Opt { OPT_TF32_MIN_ROWS, OPT_X3_MIN_ROWS, OPT_TF32_PRIOR, OPT_WN_FUSED, OPT_ATTENTION_MMA, OPT_TF32_CLUSTER,
           OPT_MRF_FUSED, OPT_DECODER_STREAMS, OPT_RESPAIR_GRID_DIV, OPT_FUSED_RESPAIR, OPT_TIMING_BUFFER, OPT_SPLIT16, OPT_RESBLOCK_FUSED, OPT_PAIR_CONV, OPT_PAIR_FUSED, OPT_COUPLING_FUSED, OPT_PDL, OPT_COUPLING_MIN_ROWS, OPT_TAP_PAIRS, OPT_CONV_SPREAD, OPT_ATTENTION_SMALL, OPT_COUNT };
constexpr int64_t kOptUnset = INT64_MIN;
struct Options { int64_t v[OPT_COUNT]; };
const Options& opts();                                     // the executing call's snapshot (outside a scope: the defaults)
int option_set(Options* o /*null = process defaults*/, const char* name, int64_t value);
struct OptionScope {
  explicit OptionScope(const Options* overrides);
  ~OptionScope();
};

// bump allocator over the caller's workspace
struct Workspace {
  char* base; int64_t size; int64_t off = 0; bool ok = true;
  Workspace(void* p, int64_t n) : base(static_cast<char*>(p)), size(n) {}
  template <typename T> T* take(int64_t n) {
    int64_t bytes = (n * int64_t(sizeof(T)) + 255) & ~int64_t(255);
    if (off + bytes > size) { ok = false; return nullptr; }
    T* r = reinterpret_cast<T*>(base + off);
    off += bytes;
    return r;
  }
};

__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }

}  // namespace vs
