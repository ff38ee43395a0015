// Windowed relative-position self-attention (reference attentions.py:148-179) on the tensor cores.
//
// Same contract as the CUDA-core kernel in ops_simt.cu (banded form of the relative terms, streaming softmax, keys =
// the utterance's own rows), but QK^T and PV run as warp-level m16n8k8 TF32 MMAs with the 3xTF32 error compensation
// (a = a_hi + a_lo, b = b_hi + b_lo, D += a_lo b_hi + a_hi b_lo + a_hi b_hi), which keeps the result at fp32 level
// (the latents downstream are held to 1e-2 after the prior sampling amplifies every error upstream).
// Why mma.sync and not tcgen05 here: the softmax sits between the two contractions and P has to be split hi/lo in
// registers before it is an operand again - with the accumulators in registers (flash-attention-2 form) that is a
// register permutation, with TMEM accumulators it is a TMEM -> RF -> smem round trip per key tile for a kernel that is
// 5 % of the step.
//
// One CTA = 64 queries of one (utterance, head), 4 warps x 16 query rows, two CTAs per SM.  Key/value tiles of 32 rows
// arrive through a cp.async double buffer as raw fp32 and are split hi/lo when a fragment is read.
// Measured (tools/attention_timing.py, 64 x 431 frames, 2 heads): 0.319 ms vs 0.446 ms for the CUDA-core kernel; the legacy
// HMMA.1688.TF32 path issues one MMA per ~13 clk per SM sub-partition on sm_100a, so the three-term form is MMA-bound at
// ~0.16 ms and plain TF32 (0.237 ms, error 6e-4) is not accurate enough for the prior.  Short sequences (phoneme level)
// stay on the CUDA-core kernel, which is faster below ~128 rows.  Fragment <-> memory maps (g = lane/4, t = lane%4):
//   A (16x8): a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4);  B (8x8): b0 (k=t, n=g) b1 (k=t+4, n=g);
//   C (16x8): c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1).
// P (a C fragment of S) becomes the A fragment of PV without any shuffle by renaming the contraction index:
// k-slot t <-> key 2t, k-slot t+4 <-> key 2t+1, applied to V's rows as well.
#include "common.cuh"

namespace vs {
namespace {

constexpr int TQ = 64, TK = 32, D = kHeadDim;
constexpr int LD = 100;              // row pitch (floats): 100 mod 32 = 4 -> every fragment load below is conflict-free
constexpr int NKS = D / 8;           // 12 k-steps over the head dim (QK^T) / 12 n-tiles over the head dim (PV)
constexpr int NNT = TK / 8;          // 4 n-tiles over the keys (QK^T) / 4 k-steps over the keys (PV)

__device__ __forceinline__ uint32_t tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = tf32_rna(x);
  lo = __float_as_uint(x - __uint_as_float(hi));     // exact difference; the MMA reads its top 19 bits (error 2^-21 |x|)
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct Smem {
  float kv[2][2][TK * LD];   // [stage][K | V] raw fp32 tiles (cp.async double buffer); the Q tile is staged in kv[1] first
  float ev[kRel * D];
  float relq[TQ * kRel];     // (q / sqrt(d)) . Ek[w]
  float pband[TQ * kRel];    // p[i, i+w-4] of the current key tile (relative values, attentions.py:174-177)
};
static_assert(2 * TK * LD >= TQ * LD, "Q staging must fit in one stage");

__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  const int n = valid ? 16 : 0;                      // src-size 0 -> the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <bool X3>
__global__ void __launch_bounds__(128, 2) rel_attention_mma_kernel(VsRows rows, const float* __restrict__ qkv,
                                                                   const float* __restrict__ ek,
                                                                   const float* __restrict__ ev,
                                                                   float* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>(smem_raw);
  const int b = blockIdx.z, h = blockIdx.y;
  const int T = rows.utt_len[b], start = rows.utt_start[b];
  const int q0 = blockIdx.x * TQ;
  if (q0 >= T) return;
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32, g = lane / 4, t = lane % 4;
  const int ld = 3 * kHidden;
  const float scale = rsqrtf((float)D);
  float* Qs = &S.kv[1][0][0];
  const float* kbase = qkv + (size_t)start * ld + h * D + kHidden;

  auto prefetch = [&](int k0, int stage) {            // K and V rows [k0, k0+TK) -> smem, zero-filled past T
    for (int i = tid; i < TK * (D / 4); i += 128) {
      const int r = i / (D / 4), d4 = i % (D / 4);
      const bool ok = k0 + r < T;
      const float* p = kbase + (size_t)(ok ? k0 + r : 0) * ld + 4 * d4;
      cp_async16(&S.kv[stage][0][r * LD + 4 * d4], p, ok);
      cp_async16(&S.kv[stage][1][r * LD + 4 * d4], p + kHidden, ok);
    }
    cp_async_commit();
  };
  prefetch(0, 0);

  for (int i = tid; i < TQ * (D / 4); i += 128) {
    const int r = i / (D / 4), d4 = i % (D / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + r < T) v = *reinterpret_cast<const float4*>(qkv + (size_t)(start + q0 + r) * ld + h * D + 4 * d4);
    v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    *reinterpret_cast<float4*>(Qs + r * LD + 4 * d4) = v;
  }
  for (int i = tid; i < kRel * D; i += 128) S.ev[i] = ev[i];
  __syncthreads();
  for (int i = tid; i < TQ * kRel; i += 128) {
    const int r = i / kRel, w = i % kRel;
    float s = 0.f;
#pragma unroll 8
    for (int d = 0; d < D; ++d) s = fmaf(Qs[r * LD + d], __ldg(ek + w * D + d), s);
    S.relq[i] = s;
  }
  // this warp's 16 query rows as A fragments (fp32; split hi/lo at use)
  const int r0 = warp * 16;
  float qa[NKS][4];
#pragma unroll
  for (int ks = 0; ks < NKS; ++ks) {
    qa[ks][0] = Qs[(r0 + g) * LD + 8 * ks + t];
    qa[ks][1] = Qs[(r0 + g + 8) * LD + 8 * ks + t];
    qa[ks][2] = Qs[(r0 + g) * LD + 8 * ks + t + 4];
    qa[ks][3] = Qs[(r0 + g + 8) * LD + 8 * ks + t + 4];
  }
  float o[NKS][4];
#pragma unroll
  for (int n = 0; n < NKS; ++n) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f; }
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};     // rows g and g+8
  const int qi0 = q0 + r0 + g, qi1 = qi0 + 8;
  const int wq_lo = q0 + r0, wq_hi = wq_lo + 15;                      // this warp's query range

  for (int k0 = 0, it = 0; k0 < T; k0 += TK, ++it) {
    cp_async_wait_all();
    __syncthreads();                    // tile `it` has landed; everyone is done with tile it-1 (and with the Q staging)
    if (k0 + TK < T) prefetch(k0 + TK, (it + 1) & 1);
    const float* Kt = &S.kv[it & 1][0][0];
    const float* Vt = &S.kv[it & 1][1][0];

    // ---- S = (Q / sqrt(d)) K^T for 16 rows x 32 keys; the three 3xTF32 terms go term-major so that consecutive
    //      MMAs never chain on one accumulator
    float s[NNT][4];
#pragma unroll
    for (int n = 0; n < NNT; ++n) { s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < NKS; ++ks) {
      uint32_t ah[4], al[4], bh[NNT][2], bl[NNT][2];
#pragma unroll
      for (int j = 0; j < 4; ++j) split_tf32(qa[ks][j], ah[j], al[j]);
#pragma unroll
      for (int n = 0; n < NNT; ++n) {
        const int off = (8 * n + g) * LD + 8 * ks + t;
        split_tf32(Kt[off], bh[n][0], bl[n][0]);
        split_tf32(Kt[off + 4], bh[n][1], bl[n][1]);
      }
      if (X3) {
#pragma unroll
        for (int n = 0; n < NNT; ++n) mma_tf32(s[n], al, bh[n][0], bh[n][1]);
#pragma unroll
        for (int n = 0; n < NNT; ++n) mma_tf32(s[n], ah, bl[n][0], bl[n][1]);
      }
#pragma unroll
      for (int n = 0; n < NNT; ++n) mma_tf32(s[n], ah, bh[n][0], bh[n][1]);
    }
    // ---- relative keys on the band, key mask, streaming softmax
    const bool band = (k0 <= wq_hi + kWindow) && (k0 + TK > wq_lo - kWindow);      // warp-uniform
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int n = 0; n < NNT; ++n)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int kj = k0 + 8 * n + 2 * t + (c & 1);
        const int qi = (c & 2) ? qi1 : qi0;
        float v = s[n][c];
        if (band) {
          const int dd = kj - qi;
          if (dd >= -kWindow && dd <= kWindow) v += S.relq[(qi - q0) * kRel + dd + kWindow];
        }
        if (kj >= T) v = -INFINITY;
        s[n][c] = v;
        if (c & 2) mx1 = fmaxf(mx1, v); else mx0 = fmaxf(mx0, v);
      }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m_run[0], mx0), mn1 = fmaxf(m_run[1], mx1);       // finite: key k0 < T is always valid
    const float al0 = (m_run[0] == -INFINITY) ? 0.f : __expf(m_run[0] - mn0);
    const float al1 = (m_run[1] == -INFINITY) ? 0.f : __expf(m_run[1] - mn1);
    m_run[0] = mn0; m_run[1] = mn1;
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int n = 0; n < NNT; ++n) {
      s[n][0] = __expf(s[n][0] - mn0); s[n][1] = __expf(s[n][1] - mn0);
      s[n][2] = __expf(s[n][2] - mn1); s[n][3] = __expf(s[n][3] - mn1);
      sum0 += s[n][0] + s[n][1]; sum1 += s[n][2] + s[n][3];
    }
    l_run[0] = l_run[0] * al0 + sum0; l_run[1] = l_run[1] * al1 + sum1;       // per-lane partial sums; reduced at the end
#pragma unroll
    for (int n = 0; n < NKS; ++n) { o[n][0] *= al0; o[n][1] *= al0; o[n][2] *= al1; o[n][3] *= al1; }

    // ---- relative values on the band: o[i] += sum_w p[i, i+w-4] Ev[w]   (fp32, exact)
    if (band) {
      float* pb = &S.pband[r0 * kRel];
      for (int i = lane; i < 16 * kRel; i += 32) pb[i] = 0.f;
      __syncwarp();
#pragma unroll
      for (int n = 0; n < NNT; ++n)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int kj = k0 + 8 * n + 2 * t + (c & 1);
          const int qi = (c & 2) ? qi1 : qi0;
          const int dd = kj - qi;
          if (dd >= -kWindow && dd <= kWindow) S.pband[(qi - q0) * kRel + dd + kWindow] = s[n][c];
        }
      __syncwarp();
      for (int w = 0; w < kRel; ++w) {
        const float p0 = pb[g * kRel + w], p1 = pb[(g + 8) * kRel + w];
#pragma unroll
        for (int n = 0; n < NKS; ++n) {
          const float2 e = *reinterpret_cast<const float2*>(&S.ev[w * D + 8 * n + 2 * t]);
          o[n][0] = fmaf(p0, e.x, o[n][0]); o[n][1] = fmaf(p0, e.y, o[n][1]);
          o[n][2] = fmaf(p1, e.x, o[n][2]); o[n][3] = fmaf(p1, e.y, o[n][3]);
        }
      }
      __syncwarp();
    }

    // ---- O += P V   (k-slot t <-> key 2t, k-slot t+4 <-> key 2t+1); 12 independent accumulators per term
#pragma unroll
    for (int kt = 0; kt < NNT; ++kt) {
      uint32_t ah[4], al[4];
      split_tf32(s[kt][0], ah[0], al[0]);      // (g,   key 2t)
      split_tf32(s[kt][2], ah[1], al[1]);      // (g+8, key 2t)
      split_tf32(s[kt][1], ah[2], al[2]);      // (g,   key 2t+1)
      split_tf32(s[kt][3], ah[3], al[3]);      // (g+8, key 2t+1)
#pragma unroll
      for (int n4 = 0; n4 < NKS; n4 += 4) {
        uint32_t bh[4][2], bl[4][2];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int off = (8 * kt + 2 * t) * LD + 8 * (n4 + j) + g;
          split_tf32(Vt[off], bh[j][0], bl[j][0]);
          split_tf32(Vt[off + LD], bh[j][1], bl[j][1]);
        }
        if (X3) {
#pragma unroll
          for (int j = 0; j < 4; ++j) mma_tf32(o[n4 + j], al, bh[j][0], bh[j][1]);
#pragma unroll
          for (int j = 0; j < 4; ++j) mma_tf32(o[n4 + j], ah, bl[j][0], bl[j][1]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) mma_tf32(o[n4 + j], ah, bh[j][0], bh[j][1]);
      }
    }
  }
  float l0 = l_run[0], l1 = l_run[1];
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float inv0 = 1.f / l0, inv1 = 1.f / l1;
  if (qi0 < T) {
    float* dst = out + (size_t)(start + qi0) * kHidden + h * D + 2 * t;
#pragma unroll
    for (int n = 0; n < NKS; ++n) *reinterpret_cast<float2*>(dst + 8 * n) = make_float2(o[n][0] * inv0, o[n][1] * inv0);
  }
  if (qi1 < T) {
    float* dst = out + (size_t)(start + qi1) * kHidden + h * D + 2 * t;
#pragma unroll
    for (int n = 0; n < NKS; ++n) *reinterpret_cast<float2*>(dst + 8 * n) = make_float2(o[n][2] * inv1, o[n][3] * inv1);
  }
}

}  // namespace

int rel_attention_mma(const VsRows& rows, const float* qkv, const float* ek, const float* ev, float* out,
                      cudaStream_t st) {
  const size_t smem = sizeof(Smem);
  VS_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(rel_attention_mma_kernel<true>), (int)smem));
  VS_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(rel_attention_mma_kernel<false>), (int)smem));
  VS_CUDA_CHECK(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)rows.n_rows * kHidden, st));   // gap rows stay zero
  VS_REQUIRE(rows.max_len > 0 && rows.max_len <= rows.n_rows, "rel_attention: bad max_len %d", rows.max_len);
  dim3 grid((rows.max_len + TQ - 1) / TQ, kHeads, rows.n_utt);
  if (opts().v[OPT_ATTENTION_MMA] == 3) rel_attention_mma_kernel<false><<<grid, 128, smem, st>>>(rows, qkv, ek, ev, out);   // plain TF32: A/B only
  else rel_attention_mma_kernel<true><<<grid, 128, smem, st>>>(rows, qkv, ek, ev, out);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

}  // namespace vs
