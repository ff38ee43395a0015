// TF32 tcgen05 conv over fp32 row-major ragged rows.  See umma_tf32.cu.
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"

namespace vs {

// Weights: fp32 words already rounded to TF32, slabs [NB][Cin/96][taps][3][8 planes][Nblk][4]:
// element (nb, ka, t, j, p, n, e) = W[t][96*ka + 32*j + 4*p + e][nb*Nblk + n]   (packing.py pack_tf32)
struct UmmaTf32 {
  const float* in = nullptr; int in_ld = 0;          // [R][in_ld], channel offset folded into the pointer
  const float* w = nullptr;
  const float* bias = nullptr;                       // [N] or null
  const float* ubias = nullptr; int ubias_ld = 0;    // optional per-speaker bias rows (column offset folded in)
  const int32_t* ubias_idx = nullptr;                // [n_utt] -> row of ubias (sid)
  float* out = nullptr; int out_ld = 0;              // [R][out_ld]
  const int32_t* row_utt = nullptr;                  // null = every row valid; invalid rows are written as zeros
  int R = 0, Cin = 0, N = 0, taps = 1, dil = 1, pad_l = 0;
  int act = 0;                                       // 0 none, 1 relu
  // Epilogue variants (both used by the flow's WN layers, modules.py:148-176):
  //  epi = 1  WN gate: weight columns are packed in blocks [16 tanh | 16 sigmoid] (packing.py gate_columns), the epilogue
  //           writes acts = tanh(a_t) * sigmoid(a_s) to out[r][N/2]              (commons.py:100-107)
  //  epi = 2  res/skip update: n-blocks < nb_split are ADDED into out (h += res), the others go to out2 (skip), added
  //           when accumulate2 else assigned; invalid rows are left untouched (out2: zeroed when assigned)
  // split3 = 1: error-compensated "3xTF32": a = a_hi + a_lo, w = w_hi + w_lo (each TF32), D += a_hi w_hi + a_lo w_hi +
  // a_hi w_lo (the 2^-22 term a_lo w_lo is dropped): fp32-level accuracy on the tensor cores at 3x the MMA count.
  // Weights then come in [hi | lo] slab pairs of 16 channels (packing.py pack_tf32(split3=True)).
  int split3 = 0;
  int epi = 0;
  float* out2 = nullptr; int out2_ld = 0; int nb_split = 0; int accumulate2 = 0;
};
int umma_tf32(const UmmaTf32& c, cudaStream_t st);

// One whole WN layer (in_layer k5 -> gate -> res_skip 1x1 -> h / skip update) in one kernel, plain TF32.  See umma_wn.cu.
struct UmmaWn {
  // h and skip are PLANAR fp32: [192/4][R][4] (rows_to_planar4 / planar4_to_rows convert at the ends of the WN stack)
  const float* h_in = nullptr;
  float* h_out = nullptr;                            // != h_in (neighbouring tiles read h_in's halo rows); unused if last
  float* skip = nullptr;                             // assigned when first, else accumulated
  const float* w_in = nullptr;                       // in_layer weights, gate-interleaved columns, pack_tf32 (plain) slabs
  const float* b_in = nullptr;                       // [384], gate-interleaved
  const float* cond = nullptr; int cond_ld = 0;      // per-speaker cond rows (gate-interleaved, this layer's 384 columns)
  const int32_t* cond_idx = nullptr;                 // [n_utt] -> row of cond (sid)
  const float* w_rs = nullptr;                       // res_skip weights, pack_tf32 (plain) slabs: 384 columns, 192 when last
  const float* b_rs = nullptr;
  const int32_t* row_utt = nullptr;
  int R = 0, first = 0, last = 0;
};
int umma_wn_layer(const UmmaWn& c, cudaStream_t st);
// One whole mean-only coupling layer of the flow (pre, 4-layer WN, post, x1 update) in one kernel (umma_coupling.cu), fp16 operands.
struct UmmaCoupling {
  float* z = nullptr;                                // [R][192] fp32 rows; x0 = columns [in_off, +96) is read, x1 = [upd_off, +96) updated in place
  const __half* w = nullptr;                         // packing.py pack_coupling: 92 slabs of 36,864 bytes
  const float* bias = nullptr;                       // packing.py pack_coupling_bias: [4][192] h biases | [96] m bias | [n_spk][4][384]
  const int32_t* row_utt = nullptr;                  // [R] utterance of a row, < 0: gap row
  const int32_t* sid = nullptr;                      // [n_utt] speaker of an utterance
  int R = 0, in_off = 0, upd_off = 96;
  float sign = -1.f;                                 // reverse: x1 -= m; forward: x1 += m
};
int umma_coupling(const UmmaCoupling& c, cudaStream_t st);
int rows_to_planar4(const float* in, float* out, int R, int C, cudaStream_t st);    // [R][C] -> [C/4][R][4]
int planar4_to_rows(const float* in, float* out, int R, int C, cudaStream_t st);    // [C/4][R][4] -> [R][C]

}  // namespace vs
