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
int rel_attention(const VsRows& rows, const float* qkv, const float* ek, const float* ev, float* out, cudaStream_t st);
// tensor-core (3xTF32 mma.sync) form of the same op, attention_mma.cu; rel_attention() dispatches on g_attention_mma
int rel_attention_mma(const VsRows& rows, const float* qkv, const float* ek, const float* ev, float* out, cudaStream_t st);
extern int g_attention_mma;                        // vs_set_option("attention_mma", 0..3), default 1 = auto
int row_dot(const float* x, int ld, const float* w, const float* bias, float* out, int R, int C,
            const int32_t* row_utt, cudaStream_t st);

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
