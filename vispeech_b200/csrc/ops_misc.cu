// Small memory-bound kernels of the path: embedding, speaker-conditioning adds, variance-adapter
// formulas and prenets, the length regulator (prefix sum + search + gather), WN gate / residual
// updates, coupling update, prior sampling, ragged-rows -> [B][C][T] unpack.
#include "ops_misc.cuh"

namespace vs {

// ---- TextEncoder embedding: x = emb[id] * sqrt(H) (models.py:169); gaps -> 0 -----------------------
__global__ void embed_rows_kernel(const int32_t* __restrict__ ids, const float* __restrict__ emb, float* __restrict__ x,
                                  int R, int n_vocab, float scale) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * kHidden) return;
  const int r = i / kHidden, c = i % kHidden;
  const int id = ids[r];
  x[i] = (id >= 0 && id < n_vocab) ? emb[(size_t)id * kHidden + c] * scale : 0.f;
}
int embed_rows(const int32_t* ids, const float* emb, float* x, int R, int n_vocab, cudaStream_t st) {
  VS_CUDA_CHECK(launch_pdl<32>(embed_rows_kernel, dim3((R * kHidden + 255) / 256), dim3(256), 0, st, ids, emb, x, R, n_vocab, sqrtf((float)kHidden)));
  VS_LAUNCH_CHECK();
  return VS_OK;
}

// ---- out = valid ? x + tab[sid[utt]] : 0   (x + cond(g): models.py:123, :509, frame_prior_network.py:121) ----
__global__ void add_speaker_rows_kernel(const float* __restrict__ x, const float* __restrict__ tab,
                                        const int32_t* __restrict__ row_utt, const int32_t* __restrict__ sid,
                                        float* __restrict__ out, int R, int C) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * C) return;
  const int r = i / C, c = i % C;
  const int u = row_utt[r];
  out[i] = (u >= 0) ? x[i] + tab[(size_t)sid[u] * C + c] : 0.f;
}
int add_speaker_rows(const float* x, const float* tab, const VsRows& rows, float* out, int C, cudaStream_t st) {
  VS_CUDA_CHECK(launch_pdl<32>(add_speaker_rows_kernel, dim3((rows.n_rows * C + 255) / 256), dim3(256), 0, st, x, tab, rows.row_utt, rows.sid, out,
                                                                        rows.n_rows, C));
  VS_LAUNCH_CHECK();
  return VS_OK;
}

// ---- variance adapter glue (models.py:681-708) ------------------------------------------------------
// duration: mode 0: ceil((exp(logw) - 1) * scale) (x_mask == 1 on valid rows); mode 2: control verbatim.
__global__ void duration_kernel(const float* __restrict__ logw, const double* __restrict__ ctrl, int mode, float scale,
                                const int32_t* __restrict__ row_utt, double* __restrict__ dur, int R) {
  pdl_trigger();
  pdl_wait();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  if (row_utt[r] < 0) { dur[r] = 0.0; return; }
  if (mode == 2) dur[r] = ctrl[r];
  else dur[r] = (double)ceilf((expf(logw[r]) - 1.f) * scale);
}
// pitch: lf0 (mode 0: pred*scale; mode 2: 2595*log10(1+hz/700)/500), F0 = (10^(lf0*500/2590)-1)*700 (sic, Q2)
__global__ void pitch_kernel(const float* __restrict__ pred, const float* __restrict__ ctrl, int mode, float scale,
                             const int32_t* __restrict__ row_utt, float* __restrict__ lf0, float* __restrict__ f0, int R) {
  pdl_trigger();
  pdl_wait();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  if (row_utt[r] < 0) { lf0[r] = 0.f; f0[r] = 0.f; return; }
  float v;
  if (mode == 2) v = (2595.f * log10f(1.f + ctrl[r] / 700.f)) / 500.f;
  else v = pred[r] * scale;
  lf0[r] = v;
  f0[r] = (powf(10.f, v * 500.f / 2590.f) - 1.f) * 700.f;
}
// energy: norm (mode 0: (((pred*36+60)*scale)-60)/36; mode 2: (raw-60)/36), energy = norm*36+60
__global__ void energy_kernel(const float* __restrict__ pred, const float* __restrict__ ctrl, int mode, float scale,
                              const int32_t* __restrict__ row_utt, float* __restrict__ norm, float* __restrict__ energy,
                              int R) {
  pdl_trigger();
  pdl_wait();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  if (row_utt[r] < 0) { norm[r] = 0.f; energy[r] = 0.f; return; }
  float v;
  if (mode == 2) v = (ctrl[r] - 60.f) / 36.f;
  else v = (((pred[r] * 36.f + 60.f) * scale) - 60.f) / 36.f;
  norm[r] = v;
  energy[r] = v * 36.f + 60.f;
}
int duration_rows(const float* logw, const double* ctrl, int mode, float scale, const VsRows& rows, double* dur,
                  cudaStream_t st) {
  VS_CUDA_CHECK(launch_pdl<32>(duration_kernel, dim3((rows.n_rows + 255) / 256), dim3(256), 0, st, logw, ctrl, mode, scale, rows.row_utt, dur, rows.n_rows));
  VS_LAUNCH_CHECK();
  return VS_OK;
}
int pitch_rows(const float* pred, const float* ctrl, int mode, float scale, const VsRows& rows, float* lf0, float* f0,
               cudaStream_t st) {
  VS_CUDA_CHECK(launch_pdl<32>(pitch_kernel, dim3((rows.n_rows + 255) / 256), dim3(256), 0, st, pred, ctrl, mode, scale, rows.row_utt, lf0, f0, rows.n_rows));
  VS_LAUNCH_CHECK();
  return VS_OK;
}
int energy_rows(const float* pred, const float* ctrl, int mode, float scale, const VsRows& rows, float* norm,
                float* energy, cudaStream_t st) {
  VS_CUDA_CHECK(launch_pdl<32>(energy_kernel, dim3((rows.n_rows + 255) / 256), dim3(256), 0, st, pred, ctrl, mode, scale, rows.row_utt, norm, energy,
                                                          rows.n_rows));
  VS_LAUNCH_CHECK();
  return VS_OK;
}

// x[r][c] += b[c] + w[c][0]*v[r-1] + w[c][1]*v[r] + w[c][2]*v[r+1]   (Conv1d(1,192,3,padding=1): models.py:612-613,697,707)
// v is zero on gap rows, which is exactly the zero padding of a batch-1 call.
__global__ void prenet_add_kernel(float* __restrict__ x, const float* __restrict__ v, const float* __restrict__ w,
                                  const float* __restrict__ b, const int32_t* __restrict__ row_utt, int R) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * kHidden) return;
  const int r = i / kHidden, c = i % kHidden;
  if (row_utt[r] < 0) return;
  const float vm = r > 0 ? v[r - 1] : 0.f, v0 = v[r], vp = r + 1 < R ? v[r + 1] : 0.f;
  x[i] += b[c] + w[c * 3 + 0] * vm + w[c * 3 + 1] * v0 + w[c * 3 + 2] * vp;
}
int prenet_add(float* x, const float* v, const float* w, const float* b, const VsRows& rows, cudaStream_t st) {
  VS_CUDA_CHECK(launch_pdl<32>(prenet_add_kernel, dim3((rows.n_rows * kHidden + 255) / 256), dim3(256), 0, st, x, v, w, b, rows.row_utt, rows.n_rows));
  VS_LAUNCH_CHECK();
  return VS_OK;
}

// ---- length regulator (models.py:398-427) -----------------------------------------------------------
// count: one CTA per utterance; n_i = max(int(d_i), 0) with int() = truncation toward zero of the
// Python float that .item() returned (models.py:421-423); inclusive scan in phoneme order.
__global__ void lr_count_kernel(VsRows rows, const double* __restrict__ dur, int32_t* __restrict__ cum,
                                int32_t* __restrict__ frames) {
  pdl_trigger();
  pdl_wait();
  __shared__ int32_t warp_tot[32];
  __shared__ int32_t carry_s;
  const int b = blockIdx.x, T = rows.utt_len[b], start = rows.utt_start[b];
  const int tid = threadIdx.x, lane = tid % 32, warp = tid / 32;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < T; base += blockDim.x) {
    const int i = base + tid;
    int32_t n = 0;
    if (i < T) {
      double d = trunc(dur[start + i]);
      if (!(d > 0.0)) d = 0.0;                 // also maps NaN to 0
      if (d > 1.0e6) d = 1.0e6;                // guard: one phoneme never spans > 1e6 frames
      n = (int32_t)d;
    }
    int32_t v = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int32_t t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
    if (lane == 31) warp_tot[warp] = v;
    __syncthreads();
    if (warp == 0) {
      int32_t w = lane < (int)(blockDim.x / 32) ? warp_tot[lane] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int32_t t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
      warp_tot[lane] = w;
    }
    __syncthreads();
    const int32_t before = carry_s + (warp > 0 ? warp_tot[warp - 1] : 0);
    if (i < T) cum[start + i] = before + v;
    __syncthreads();
    if (tid == 0) carry_s += warp_tot[blockDim.x / 32 - 1];
    __syncthreads();
  }
  if (tid == 0) frames[b] = carry_s;
}
int lr_count(const VsRows& rows, const double* dur, int32_t* cum, int32_t* frames, cudaStream_t st) {
  VS_CUDA_CHECK(cudaMemsetAsync(cum, 0, sizeof(int32_t) * rows.n_rows, st));
  VS_CUDA_CHECK(launch_pdl<32>(lr_count_kernel, dim3(rows.n_utt), dim3(256), 0, st, rows, dur, cum, frames));
  VS_LAUNCH_CHECK();
  return VS_OK;
}

// gather: frame t of utterance b copies phoneme idx = #{i : cum[i] <= t} (== searchsorted(cum, t, right=True)),
// which is the index the reference's expand()+cat() places at frame t.  One warp per frame row.
__global__ void lr_gather_kernel(VsRows rp, VsRows rf, const float* __restrict__ xp, const int32_t* __restrict__ cum,
                                 float* __restrict__ xf, int32_t* __restrict__ lr_index) {
  pdl_trigger();
  pdl_wait();
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x % 32;
  if (r >= rf.n_rows) return;
  const int b = rf.row_utt[r];
  float* dst = xf + (size_t)r * kHidden;
  if (b < 0) {
    for (int c = lane; c < kHidden; c += 32) dst[c] = 0.f;
    if (lane == 0) lr_index[r] = -1;
    return;
  }
  const int t = r - rf.utt_start[b];
  const int32_t* cb = cum + rp.utt_start[b];
  int lo = 0, hi = rp.utt_len[b];          // first i with cum[i] > t
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (cb[mid] <= t) lo = mid + 1; else hi = mid;
  }
  const float* src = xp + (size_t)(rp.utt_start[b] + lo) * kHidden;
  for (int c = lane; c < kHidden; c += 32) dst[c] = src[c];
  if (lane == 0) lr_index[r] = lo;
}
int lr_gather(const VsRows& rp, const VsRows& rf, const float* xp, const int32_t* cum, float* xf, int32_t* lr_index,
              cudaStream_t st) {
  VS_CUDA_CHECK(launch_pdl<32>(lr_gather_kernel, dim3((rf.n_rows + 7) / 8), dim3(256), 0, st, rp, rf, xp, cum, xf, lr_index));
  VS_LAUNCH_CHECK();
  return VS_OK;
}

// ---- prior sample: z_p = m_p + eps * exp(logs_p) * noise_scale (models.py:718); stats = [m_p | logs_p] ----
// eps ~ N(0, 1) is the one RNG draw of the path (`torch.randn_like`, models.py:718).  Injected by the caller for parity
// runs; otherwise generated here, in the kernel that consumes it: counter-based Philox4x32-10 (Salmon et al., SC'11;
// key = the call's 64-bit seed, counter = element index) -> two uniforms -> Box-Muller.  No eps tensor ever exists in HBM
// and the result does not depend on the launch geometry.
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
  c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}
__device__ __forceinline__ float philox_normal(uint64_t seed, uint64_t index) {
  uint32_t c[4] = {(uint32_t)index, (uint32_t)(index >> 32), 0u, 0u};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  const float u1 = ((float)(c[0] >> 8) + 0.5f) * (1.f / 16777216.f);      // (0, 1): log() stays finite
  const float u2 = ((float)(c[1] >> 8) + 0.5f) * (1.f / 16777216.f);
  return sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
}
// eps as a tensor: the same stream as the in-kernel draw (element i of a call with this seed), for callers that must keep the
// seed out of the kernel arguments (CUDA-graph replays: the graph reads eps from a static buffer refilled before each replay)
__global__ void randn_kernel(float* __restrict__ out, int64_t n, uint64_t seed) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = philox_normal(seed, (uint64_t)i);
}
int randn_fill(float* out, int64_t n, uint64_t seed, cudaStream_t st) {
  VS_REQUIRE(out && n > 0, "randn_fill: bad arguments");
  randn_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(out, n, seed);
  VS_LAUNCH_CHECK();
  return VS_OK;
}
__global__ void prior_sample_kernel(const float* __restrict__ stats, const float* __restrict__ noise, uint64_t seed, float ns,
                                    const int32_t* __restrict__ row_utt, float* __restrict__ m_p,
                                    float* __restrict__ logs_p, float* __restrict__ z_p, int R) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * kHidden) return;
  const int r = i / kHidden, c = i % kHidden;
  if (row_utt[r] < 0) { m_p[i] = 0.f; logs_p[i] = 0.f; z_p[i] = 0.f; return; }
  const float m = stats[(size_t)r * 2 * kHidden + c], ls = stats[(size_t)r * 2 * kHidden + kHidden + c];
  m_p[i] = m;
  logs_p[i] = ls;
  const float eps = noise ? noise[i] : philox_normal(seed, (uint64_t)i);
  z_p[i] = m + eps * expf(ls) * ns;
}
int prior_sample(const float* stats, const float* noise, uint64_t seed, float ns, const VsRows& rows, float* m_p,
                 float* logs_p, float* z_p, cudaStream_t st) {
  VS_CUDA_CHECK(launch_pdl<32>(prior_sample_kernel, dim3((rows.n_rows * kHidden + 255) / 256), dim3(256), 0, st, stats, noise, seed, ns, rows.row_utt, m_p,
                                                                          logs_p, z_p, rows.n_rows));
  VS_LAUNCH_CHECK();
  return VS_OK;
}

// ---- WN pieces (modules.py:148-176) -----------------------------------------------------------------
// acts = tanh(a[:192] + g[:192]) * sigmoid(a[192:] + g[192:])   (commons.py:100-107); g = cond row of the speaker
__global__ void wn_gate_kernel(const float* __restrict__ a, const float* __restrict__ cond, int cond_ld, int cond_off,
                               const int32_t* __restrict__ row_utt, const int32_t* __restrict__ sid,
                               float* __restrict__ acts, int R) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * kHidden) return;
  const int r = i / kHidden, c = i % kHidden;
  const int u = row_utt[r];
  if (u < 0) { acts[i] = 0.f; return; }
  const float* g = cond + (size_t)sid[u] * cond_ld + cond_off;
  const float t = a[(size_t)r * 2 * kHidden + c] + g[c];
  const float s = a[(size_t)r * 2 * kHidden + kHidden + c] + g[kHidden + c];
  acts[i] = tanhf(t) * (1.f / (1.f + expf(-s)));
}
int wn_gate(const float* a, const float* cond, int cond_ld, int cond_off, const VsRows& rows, float* acts,
            cudaStream_t st) {
  wn_gate_kernel<<<(rows.n_rows * kHidden + 255) / 256, 256, 0, st>>>(a, cond, cond_ld, cond_off, rows.row_utt, rows.sid,
                                                                     acts, rows.n_rows);
  VS_LAUNCH_CHECK();
  return VS_OK;
}
// h += rs[:, :192]; skip (+)= rs[:, 192:]   (or skip += rs when last)
__global__ void wn_update_kernel(const float* __restrict__ rs, int rs_ld, int last, int first,
                                 const int32_t* __restrict__ row_utt, float* __restrict__ h, float* __restrict__ skip,
                                 int R) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * kHidden) return;
  const int r = i / kHidden, c = i % kHidden;
  if (row_utt[r] < 0) { if (first) skip[i] = 0.f; return; }
  const float* row = rs + (size_t)r * rs_ld;
  float sk;
  if (last) sk = row[c];
  else { h[i] += row[c]; sk = row[kHidden + c]; }
  skip[i] = first ? sk : skip[i] + sk;
}
int wn_update(const float* rs, int rs_ld, int last, int first, const VsRows& rows, float* h, float* skip,
              cudaStream_t st) {
  wn_update_kernel<<<(rows.n_rows * kHidden + 255) / 256, 256, 0, st>>>(rs, rs_ld, last, first, rows.row_utt, h, skip,
                                                                       rows.n_rows);
  VS_LAUNCH_CHECK();
  return VS_OK;
}
// coupling with mean_only (modules.py:334-343): x1 = x1 + sign * m on valid rows (sign -1: reverse, +1: forward)
__global__ void coupling_update_kernel(float* __restrict__ z, int z_off, const float* __restrict__ m, float sign,
                                       const int32_t* __restrict__ row_utt, int R) {
  const int half = kHidden / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * half) return;
  const int r = i / half, c = i % half;
  if (row_utt[r] < 0) return;
  z[(size_t)r * kHidden + z_off + c] += sign * m[i];
}
int coupling_update(float* z, int z_off, const float* m, float sign, const VsRows& rows, cudaStream_t st) {
  coupling_update_kernel<<<(rows.n_rows * (kHidden / 2) + 255) / 256, 256, 0, st>>>(z, z_off, m, sign, rows.row_utt,
                                                                                   rows.n_rows);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

// ---- copy with validity mask limited to the first max_len frames of each utterance (models.py:720) ----
__global__ void mask_frames_kernel(VsRows rows, int max_len, int32_t* __restrict__ row_utt_out) {
  pdl_trigger();
  pdl_wait();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows.n_rows) return;
  int u = rows.row_utt[r];
  if (u >= 0 && max_len >= 0 && r - rows.utt_start[u] >= max_len) u = -1;
  row_utt_out[r] = u;
}
int mask_frames(const VsRows& rows, int max_len, int32_t* row_utt_out, cudaStream_t st) {
  VS_CUDA_CHECK(launch_pdl<32>(mask_frames_kernel, dim3((rows.n_rows + 255) / 256), dim3(256), 0, st, rows, max_len, row_utt_out));
  VS_LAUNCH_CHECK();
  return VS_OK;
}
__global__ void masked_copy_kernel(const float* __restrict__ x, const int32_t* __restrict__ row_utt,
                                   float* __restrict__ out, int R, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * C) return;
  out[i] = row_utt[i / C] >= 0 ? x[i] : 0.f;
}
int masked_copy(const float* x, const int32_t* row_utt, float* out, int R, int C, cudaStream_t st) {
  masked_copy_kernel<<<(R * C + 255) / 256, 256, 0, st>>>(x, row_utt, out, R, C);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

// ---- ragged rows -> [n_utt][C][t_max] (the reference's output layout), tiled transpose ----------------
__global__ void unpack_rows_kernel(VsRows rows, const float* __restrict__ x, int C, int mul, int t_max,
                                   float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int len = rows.utt_len[b] * mul;
  const size_t start = (size_t)rows.utt_start[b] * mul;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int t = t0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (t < len && t < t_max && c < C) ? x[(start + t) * C + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, t = t0 + threadIdx.x;
    if (c < C && t < t_max) out[((size_t)b * C + c) * t_max + t] = tile[threadIdx.x][i];
  }
}
// C == 1 (the waveform): no transpose, a masked contiguous copy per utterance, 4 samples per thread
__global__ void unpack_wave_kernel(VsRows rows, const float* __restrict__ x, int mul, int t_max, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y;
  const int len = rows.utt_len[b] * mul;
  const size_t start = (size_t)rows.utt_start[b] * mul;
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (t >= t_max) return;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (t + 3 < len) v = *reinterpret_cast<const float4*>(x + start + t);
  else {
    if (t + 0 < len) v.x = x[start + t + 0];
    if (t + 1 < len) v.y = x[start + t + 1];
    if (t + 2 < len) v.z = x[start + t + 2];
  }
  float* o = out + (size_t)b * t_max + t;
  if (t + 3 < t_max) *reinterpret_cast<float4*>(o) = v;
  else {
    o[0] = v.x;
    if (t + 1 < t_max) o[1] = v.y;
    if (t + 2 < t_max) o[2] = v.z;
  }
}

int unpack_rows(const VsRows& rows, const float* x, int C, int mul, int t_max, float* out, cudaStream_t st) {
  VS_REQUIRE(t_max > 0 && C > 0, "unpack_rows: empty output");
  if (C == 1 && mul % 4 == 0 && t_max % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    dim3 grid((t_max / 4 + 255) / 256, rows.n_utt);
    VS_CUDA_CHECK(launch_pdl<32>(unpack_wave_kernel, dim3(grid), dim3(256), 0, st, rows, x, mul, t_max, out));
    VS_LAUNCH_CHECK();
    return VS_OK;
  }
  dim3 grid((t_max + 31) / 32, (C + 31) / 32, rows.n_utt);
  VS_CUDA_CHECK(launch_pdl<32>(unpack_rows_kernel, dim3(grid), dim3(dim3(32, 8)), 0, st, rows, x, C, mul, t_max, out));
  VS_LAUNCH_CHECK();
  return VS_OK;
}

// ---- 8(f) spectrogram / log-mel on the GPU (reference mel_processing.py:50-112) --------------------------------------
// The STFT (n_fft = win = 4 hops, hann, center=False after a reflect pad of (n_fft-hop)/2) is a 4-tap conv over rows of
// one hop each with the windowed DFT basis as weights, i.e. a GEMM for umma_tf32.cu in 3xTF32.  These three kernels are
// the memory-bound glue around it.
// rows: utterance b owns n_frames[b] + 3 rows; row j holds padded samples [j*hop, (j+1)*hop) in columns [0, hop)
__global__ void mel_frame_rows_kernel(VsRows rows, const float* __restrict__ wave, int t_max,
                                      const int32_t* __restrict__ n_samples, int hop, int ld, int pad,
                                      float* __restrict__ out) {
  const int r = blockIdx.x;
  const int b = rows.row_utt[r];
  float* o = out + (size_t)r * ld;
  if (b < 0) {
    for (int c = threadIdx.x; c < ld; c += blockDim.x) o[c] = 0.f;
    return;
  }
  const int j = r - rows.utt_start[b], L = n_samples[b];
  const float* w = wave + (size_t)b * t_max;
  for (int c = threadIdx.x; c < ld; c += blockDim.x) {
    float v = 0.f;
    if (c < hop) {
      int sidx = j * hop + c - pad;                      // torch 'reflect': no edge repeat
      if (sidx < 0) sidx = -sidx;
      if (sidx >= L) sidx = 2 * (L - 1) - sidx;
      if (sidx >= 0 && sidx < L) v = w[sidx];
    }
    o[c] = v;
  }
}
// mag[r][f] = sqrt(re^2 + im^2 + 1e-6) (mel_processing.py:69,107); columns [n_bins, ld_out) are zero padding for the mel GEMM
__global__ void mel_magnitude_kernel(const float* __restrict__ dft, int ld_in, int im_off, int n_bins, int ld_out, int R,
                                     float* __restrict__ mag) {
  const int r = blockIdx.y, f = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R || f >= ld_out) return;
  float v = 0.f;
  if (f < n_bins) {
    const float re = dft[(size_t)r * ld_in + f], im = dft[(size_t)r * ld_in + im_off + f];
    v = sqrtf(re * re + im * im + 1e-6f);
  }
  mag[(size_t)r * ld_out + f] = v;
}
// out[b][c][j] = x[row(b, j)][c] (optionally log(max(x, 1e-5)): spectral_normalize_torch) for j < n_frames[b], else 0
__global__ void mel_unpack_kernel(VsRows rows, const float* __restrict__ x, int ld, int C, int t_max, int take_log,
                                  float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int len = rows.utt_len[b] - 3;
  const size_t start = (size_t)rows.utt_start[b];
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int t = t0 + i, c = c0 + threadIdx.x;
    float v = 0.f;
    if (t < len && t < t_max && c < C) {
      v = x[(start + t) * ld + c];
      if (take_log) v = logf(fmaxf(v, 1e-5f));
    }
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, t = t0 + threadIdx.x;
    if (c < C && t < t_max) out[((size_t)b * C + c) * t_max + t] = tile[threadIdx.x][i];
  }
}

int mel_frame_rows(const VsRows& rows, const float* wave, int t_max, const int32_t* n_samples, int hop, int ld, int pad,
                   float* out, cudaStream_t st) {
  mel_frame_rows_kernel<<<rows.n_rows, 192, 0, st>>>(rows, wave, t_max, n_samples, hop, ld, pad, out);
  VS_LAUNCH_CHECK();
  return VS_OK;
}
int mel_magnitude(const float* dft, int ld_in, int im_off, int n_bins, int ld_out, int R, float* mag, cudaStream_t st) {
  dim3 grid((ld_out + 255) / 256, R);
  mel_magnitude_kernel<<<grid, 256, 0, st>>>(dft, ld_in, im_off, n_bins, ld_out, R, mag);
  VS_LAUNCH_CHECK();
  return VS_OK;
}
int mel_unpack(const VsRows& rows, const float* x, int ld, int C, int t_max, int take_log, float* out, cudaStream_t st) {
  VS_REQUIRE(t_max > 0 && C > 0, "mel_unpack: empty output");
  dim3 grid((t_max + 31) / 32, (C + 31) / 32, rows.n_utt);
  mel_unpack_kernel<<<grid, dim3(32, 8), 0, st>>>(rows, x, ld, C, t_max, take_log, out);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

// ---- 8(f) waveform post-processing (reference inference_api.py:50-51: scipy wav write + `ffmpeg -ar 22050`) ------------
// out[b][t] = s16( sum_k fir[k] * x[b][decimate*t + k - (ntaps-1)/2] ), samples outside [0, n_samples[b]) are zero;
// decimate == 1 and ntaps == 0: plain float -> s16.  s16(v) = clip(rint(v * 32768)) (round half to even), ffmpeg's rule.
__global__ void pcm16_kernel(const float* __restrict__ x, int T, const int32_t* __restrict__ n_samples, int decimate,
                             const float* __restrict__ fir, int ntaps, int16_t* __restrict__ out, int T_out) {
  extern __shared__ float taps[];
  for (int i = threadIdx.x; i < ntaps; i += blockDim.x) taps[i] = fir[i];
  __syncthreads();
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T_out) return;
  const int n = n_samples[b];
  const float* xb = x + (size_t)b * T;
  float v = 0.f;
  if (ntaps == 0) {
    const int i = decimate * t;
    v = (i < n && i < T) ? xb[i] : 0.f;
  } else {
    const int c = decimate * t - (ntaps - 1) / 2;
    for (int k = 0; k < ntaps; ++k) {
      const int i = c + k;
      if (i >= 0 && i < n && i < T) v = fmaf(taps[k], xb[i], v);
    }
  }
  int q = __float2int_rn(v * 32768.f);
  q = q > 32767 ? 32767 : (q < -32768 ? -32768 : q);
  out[(size_t)b * T_out + t] = (int16_t)q;
}
int pcm16(const float* x, int B, int T, const int32_t* n_samples, int decimate, const float* fir, int ntaps, int16_t* out,
          int T_out, cudaStream_t st) {
  VS_REQUIRE(B > 0 && T > 0 && T_out > 0 && decimate >= 1 && ntaps >= 0 && ntaps <= 1024 && (ntaps == 0 || fir),
             "pcm16: bad arguments");
  dim3 grid((T_out + 255) / 256, B);
  pcm16_kernel<<<grid, 256, sizeof(float) * (ntaps > 0 ? ntaps : 1), st>>>(x, T, n_samples, decimate, fir, ntaps, out, T_out);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

}  // namespace vs
