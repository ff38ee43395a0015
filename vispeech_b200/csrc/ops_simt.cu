// fp32 CUDA-core kernels over ragged rows: generic conv1d (implicit GEMM), channel LayerNorm,
// windowed relative-position attention (streaming softmax), row dot products.
// These serve the HBM/latency-bound part of the path (text encoder, predictors, frame prior, flow);
// the decoder's dense contractions run on tcgen05 (umma_conv.cu).
#include <stdarg.h>
#include <cooperative_groups.h>
#include <mutex>
#include <unordered_set>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace vs {

unsigned long long g_launch_count = 0;
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

namespace {
std::mutex g_cfg_mutex;
int g_sm_count[64];                                   // 0 = not queried yet
std::unordered_set<uint64_t> g_smem_optin;            // (kernel address, device) pairs already configured
}  // namespace

namespace {
const char* const kOptNames[OPT_COUNT] = {"tf32_min_rows", "x3_min_rows", "tf32_prior", "wn_fused", "attention_mma", "tf32_cluster",
                                          "mrf_fused", "decoder_streams", "respair_grid_div", "fused_respair", "umma_timing_buffer", "split16", "resblock_fused", "pair_conv", "pair_fused", "coupling_fused", "pdl", "coupling_min_rows", "tap_pairs", "conv_spread", "attention_small"};
Options g_defaults = {{4096, 256, 0, 1, 1, 1, 1, 0, 1, 2, 0, 1, 1, 1, 1, 1, 33, 1, 0, 1, 1}};
thread_local Options tl_opts;
thread_local int tl_scope_depth = 0;
}  // namespace

const Options& opts() {
  if (tl_scope_depth == 0) {
    std::lock_guard<std::mutex> lock(g_cfg_mutex);
    tl_opts = g_defaults;
  }
  return tl_opts;
}

OptionScope::OptionScope(const Options* overrides) {
  if (tl_scope_depth++ == 0) {
    std::lock_guard<std::mutex> lock(g_cfg_mutex);
    tl_opts = g_defaults;
    if (overrides)
      for (int i = 0; i < OPT_COUNT; ++i)
        if (overrides->v[i] != kOptUnset) tl_opts.v[i] = overrides->v[i];
  }
}
OptionScope::~OptionScope() { --tl_scope_depth; }

thread_local int tl_pdl_extra = 0;
bool pdl_enabled(int group) { return ((opts().v[OPT_PDL] | tl_pdl_extra) & group) != 0; }
PdlExtra::PdlExtra(int groups) : saved(tl_pdl_extra) { tl_pdl_extra |= groups; }
PdlExtra::~PdlExtra() { tl_pdl_extra = saved; }

int option_set(Options* o, const char* name, int64_t value) {
  VS_REQUIRE(name, "option: null name");
  int idx = -1;
  for (int i = 0; i < OPT_COUNT; ++i)
    if (std::string(name) == kOptNames[i]) idx = i;
  VS_REQUIRE(idx >= 0, "unknown option '%s'", name);
  switch (idx) {
    case OPT_TF32_MIN_ROWS: case OPT_X3_MIN_ROWS: case OPT_COUPLING_MIN_ROWS: VS_REQUIRE(value >= 1, "option %s must be >= 1", name); break;
    case OPT_TF32_PRIOR: case OPT_WN_FUSED: case OPT_MRF_FUSED: case OPT_SPLIT16: case OPT_RESBLOCK_FUSED: case OPT_COUPLING_FUSED: case OPT_CONV_SPREAD: case OPT_ATTENTION_SMALL: value = value != 0; break;
    case OPT_TAP_PAIRS: VS_REQUIRE(value >= 0 && value <= 2, "option tap_pairs must be 0, 1 (k = 11 only) or 2 (every k)"); break;
    case OPT_PDL: VS_REQUIRE(value >= 0 && value <= 255, "option pdl is a bit mask 0..255"); break;
    case OPT_PAIR_FUSED: VS_REQUIRE(value >= 0 && value <= 2, "option pair_fused must be 0, 1 (C = 128) or 2 (also C = 64)"); break;
    case OPT_PAIR_CONV: VS_REQUIRE(value >= 0 && value <= 2, "option pair_conv must be 0 (off), 1 (C = 128) or 2 (also C = 256, k = 3)"); break;
    case OPT_ATTENTION_MMA: VS_REQUIRE(value >= 0 && value <= 4, "option attention_mma must be 0..4"); break;
    case OPT_TF32_CLUSTER: value = value == 2 ? 2 : 1; break;
    case OPT_DECODER_STREAMS: VS_REQUIRE(value >= 0 && value <= 2, "option decoder_streams must be 0 (auto), 1 or 2"); break;
    case OPT_RESPAIR_GRID_DIV: value = value < 1 ? 1 : value; break;
    case OPT_FUSED_RESPAIR: VS_REQUIRE(value >= 0 && value <= 2, "option fused_respair must be 0..2"); break;
    default: break;
  }
  std::lock_guard<std::mutex> lock(g_cfg_mutex);
  (o ? o : &g_defaults)->v[idx] = value;
  return VS_OK;
}

int device_sm_count(int* n_sm) {
  int dev = 0;
  VS_CUDA_CHECK(cudaGetDevice(&dev));
  VS_REQUIRE(dev >= 0 && dev < 64, "device index %d out of range", dev);
  std::lock_guard<std::mutex> lock(g_cfg_mutex);
  if (!g_sm_count[dev]) VS_CUDA_CHECK(cudaDeviceGetAttribute(&g_sm_count[dev], cudaDevAttrMultiProcessorCount, dev));
  *n_sm = g_sm_count[dev];
  return VS_OK;
}

int ensure_dynamic_smem(const void* kernel, int bytes) {
  int dev = 0;
  VS_CUDA_CHECK(cudaGetDevice(&dev));
  VS_REQUIRE(dev >= 0 && dev < 64, "device index %d out of range", dev);
  const uint64_t key = (reinterpret_cast<uint64_t>(kernel) << 6) ^ (uint64_t)dev;
  std::lock_guard<std::mutex> lock(g_cfg_mutex);
  if (g_smem_optin.count(key)) return VS_OK;
  VS_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  g_smem_optin.insert(key);
  return VS_OK;
}

// ------------------------------------------------------------------------------------------------
// conv1d_f32: out[orow(r)][co] = f( sum_{j<k} sum_ci W[j][ci][co] * lrelu(in[r + (j-pad_l)*dil][ci]) )
// Tile 64 rows x 64 couts per CTA, K-step 16 input channels per tap, 256 threads x (4x4) outputs.
// Rows outside [0,R) read as zero; gap rows of the input are zero by the ragged-rows invariant.
// ------------------------------------------------------------------------------------------------
constexpr int CV_BM = 64, CV_BN = 64, CV_BK = 16;

// Split-K over a thread-block cluster: the latency path (batch 1, a few hundred rows at most) has 1-6 row tiles per
// conv, so the K loop (up to 3 x 768 channels) is cut into `nchunk` slices, one CTA of a (1,1,nchunk) cluster each;
// the partial tiles are summed in rank order through distributed shared memory by rank 0 (deterministic - no atomics),
// which then runs the epilogue.  `nchunk` depends on the conv's shape only, and the un-clustered form (many row tiles)
// accumulates the same chunks from zero and adds them in the same order, so a row's result is bit-identical whatever
// else is in the batch (the serving queue and the batch-invariance tests rely on it).
// Operands arrive through a 4-deep cp.async ring (the latency path is a chain of 6-18 dependent K-steps per CTA; with one
// step of register prefetch each step cost a full L2 round trip).
constexpr int CV_ST = 4;                 // cp.async ring depth: three K-steps of loads are in flight under the FMAs of one
constexpr int CV_ALD = CV_BK + 4;        // A tile row pitch (floats): 16-byte aligned rows, warp-broadcast reads

__device__ __forceinline__ void cv_cp_async16(void* dst, const void* src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  const int n = valid ? 16 : 0;                      // src-size 0: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}

template <bool CLUSTER>
__global__ void __launch_bounds__(256) conv1d_f32_kernel(ConvF32 a, int nchunk) {
  __shared__ __align__(16) float ring[CV_ST * (CV_BM * CV_ALD + CV_BK * CV_BN)];   // 36 KB; reused for the cluster reduce
  pdl_trigger();
  pdl_wait();
  float* const As = ring;                                  // [stage][row][CV_ALD]  (row-major: k contiguous)
  float* const Bs = ring + CV_ST * CV_BM * CV_ALD;         // [stage][k][CV_BN]
  const int tid = threadIdx.x;
  const int r0 = blockIdx.x * CV_BM, c0 = blockIdx.y * CV_BN;
  const int ty = tid / 16, tx = tid % 16;
  float acc[4][4], tot[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][j] = 0.f; tot[i][j] = 0.f; }

  const int am = tid / 4, akq = tid % 4;          // A loader: row am, channels 4*akq..+3
  const int bk = tid / 16, bn4 = tid % 16;        // B loader: k row bk, couts 4*bn4..+3
  const bool vec_b = (a.Cout % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.w) & 15) == 0);
  const int per_tap = a.Cin / CV_BK, n_it = a.k * per_tap;
  // this CTA's K-steps: one chunk (cluster rank) or all of them, chunk by chunk
  const int it_lo = CLUSTER ? (int)((long long)n_it * blockIdx.z / nchunk) : 0;
  const int it_hi = CLUSTER ? (int)((long long)n_it * (blockIdx.z + 1) / nchunk) : n_it;

  auto issue = [&](int it) {
    const int st = (it - it_lo) % CV_ST;
    const int j = it / per_tap, ci0 = (it % per_tap) * CV_BK;
    const int ar = r0 + am + (j - a.pad_l) * a.dil;
    const bool a_ok = ar >= 0 && ar < a.R;
    cv_cp_async16(As + (st * CV_BM + am) * CV_ALD + 4 * akq, a.in + (size_t)(a_ok ? ar : 0) * a.in_ld + ci0 + 4 * akq, a_ok);
    const float* wrow = a.w + ((size_t)j * a.Cin + ci0 + bk) * a.Cout + c0 + 4 * bn4;
    float* bdst = Bs + (st * CV_BK + bk) * CV_BN + 4 * bn4;
    const int cc = c0 + 4 * bn4;
    if (vec_b) {
      const bool b_ok = cc + 3 < a.Cout;
      cv_cp_async16(bdst, b_ok ? wrow : a.w, b_ok);
    } else {                                      // ragged / unaligned Cout (the 1-channel convs): plain loads
#pragma unroll
      for (int e = 0; e < 4; ++e) bdst[e] = (cc + e < a.Cout) ? wrow[e] : 0.f;
    }
  };

#pragma unroll
  for (int s = 0; s < CV_ST - 1; ++s) {
    if (it_lo + s < it_hi) issue(it_lo + s);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  int ch = 0, ch_end = CLUSTER ? it_hi : (int)((long long)n_it / nchunk);
  for (int it = it_lo; it < it_hi; ++it) {
    asm volatile("cp.async.wait_group %0;" ::"n"(CV_ST - 2) : "memory");
    __syncthreads();                              // K-step `it` has landed for everyone; K-step it-1 is consumed
    if (it + CV_ST - 1 < it_hi) issue(it + CV_ST - 1);
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (!CLUSTER && it == ch_end) {               // chunk sums are added in chunk order, like rank 0 does below
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          tot[i][jj] = (ch == 0) ? acc[i][jj] : tot[i][jj] + acc[i][jj];
          acc[i][jj] = 0.f;
        }
      ++ch;
      ch_end = (int)((long long)n_it * (ch + 1) / nchunk);
    }
    const int st = (it - it_lo) % CV_ST;
    const float* As_s = As + (st * CV_BM + 4 * ty) * CV_ALD;
    const float* Bs_s = Bs + st * CV_BK * CV_BN + 4 * tx;
    float xa[4][CV_BK];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int q = 0; q < CV_BK / 4; ++q) {
        float4 v = *reinterpret_cast<const float4*>(As_s + i * CV_ALD + 4 * q);
        if (a.in_slope != 1.f) {
          v.x = lrelu(v.x, a.in_slope); v.y = lrelu(v.y, a.in_slope);
          v.z = lrelu(v.z, a.in_slope); v.w = lrelu(v.w, a.in_slope);
        }
        xa[i][4 * q] = v.x; xa[i][4 * q + 1] = v.y; xa[i][4 * q + 2] = v.z; xa[i][4 * q + 3] = v.w;
      }
#pragma unroll
    for (int kk = 0; kk < CV_BK; ++kk) {
      const float4 y = *reinterpret_cast<const float4*>(Bs_s + kk * CV_BN);
      const float ya[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc[i][jj] = fmaf(xa[i][kk], ya[jj], acc[i][jj]);
    }
  }
  if (!CLUSTER) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) acc[i][jj] = (ch == 0) ? acc[i][jj] : tot[i][jj] + acc[i][jj];
  }

  if (CLUSTER) {
    static_assert(sizeof(ring) >= 16 * 256 * sizeof(float), "partial tile must fit in the ring");
    float* const Ps = ring;                             // this CTA's partial tile, [i*4+jj][tid]; the ring is drained
    cg::cluster_group cluster = cg::this_cluster();
    __syncthreads();                                    // every thread is done reading the ring
    const int rank = (int)blockIdx.z;
    if (rank != 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) Ps[(i * 4 + jj) * 256 + tid] = acc[i][jj];
    }
    cluster.sync();
    if (rank == 0) {
      for (int r = 1; r < nchunk; ++r) {
        const float* remote = cluster.map_shared_rank(Ps, r);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) acc[i][jj] += remote[(i * 4 + jj) * 256 + tid];
      }
    }
    cluster.sync();                                     // remote tiles stay alive until rank 0 has read them
    if (rank != 0) return;
  }

  const int R_out = a.R_out ? a.R_out : a.R;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + 4 * ty + i;
    if (r >= a.R) continue;
    const int orow = r * a.out_row_mul + a.out_row_off;
    if (orow >= R_out) continue;
    int utt = 0;
    if (a.row_utt) utt = a.row_utt[orow / a.row_div];
    const bool valid = utt >= 0;
    const float* ub = nullptr;
    if (a.ubias && valid) ub = a.ubias + (size_t)(a.ubias_idx ? a.ubias_idx[utt] : utt) * a.ubias_ld;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int c = c0 + 4 * tx + jj;
      if (c >= a.Cout) continue;
      float* o = a.out + (size_t)orow * a.out_ld + c;
      if (!valid) { if (!a.accumulate) *o = 0.f; continue; }
      float y = acc[i][jj];
      if (a.bias) y += a.bias[c];
      if (ub) y += ub[c];
      if (a.act == 1) y = fmaxf(y, 0.f);
      else if (a.act == 2) y = tanhf(y);
      if (a.res) y += a.res[(size_t)orow * a.res_ld + c];
      y *= a.out_scale;
      *o = a.accumulate ? (*o + y) : y;
    }
  }
}

int conv1d_f32(const ConvF32& a, cudaStream_t st) {
  VS_REQUIRE(a.Cin % CV_BK == 0, "conv1d_f32: Cin=%d must be a multiple of %d", a.Cin, CV_BK);
  VS_REQUIRE(a.in_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(a.in) & 15) == 0, "conv1d_f32: input not 16B aligned");
  VS_REQUIRE(a.R > 0 && a.Cout > 0, "conv1d_f32: empty problem");
  dim3 grid((a.R + CV_BM - 1) / CV_BM, (a.Cout + CV_BN - 1) / CV_BN);
  // K is always accumulated in `nchunk` slices (a function of the conv's shape only); the slices run on a cluster's CTAs
  // when the tile grid alone would leave most SMs idle (latency path), else one after the other in one CTA
  const int n_it = a.k * (a.Cin / CV_BK), tiles = (int)(grid.x * grid.y);
  int nchunk = 1;
  while (nchunk < 8 && n_it / (nchunk * 2) >= 4) nchunk *= 2;
  if (nchunk == 1 || tiles * nchunk > 296) {
    VS_CUDA_CHECK(launch_pdl(conv1d_f32_kernel<false>, grid, dim3(256), 0, st, a, nchunk));
  } else {
    grid.z = nchunk;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = nchunk;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled(1) ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 2;
    VS_CUDA_CHECK(cudaLaunchKernelEx(&cfg, conv1d_f32_kernel<true>, a, nchunk));
  }
  VS_LAUNCH_CHECK();
  return VS_OK;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over channels of (a + b), one warp per row; invalid rows are written as zero.
// modules.py:29-32 (gamma/beta) and nn.LayerNorm(768) of frame_prior_network.py:83,95 share this.
// ------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(256) layernorm_rows_kernel(const float* a, const float* b,   // a/b may alias out (in-place residual LN)
                                                             int n_b, int64_t b_stride,        // b = sum of n_b partials (K-slices of umma_split.cu)
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             float* out, __half* __restrict__ out_hi, __half* __restrict__ out_lo,
                                                             int R, const int32_t* __restrict__ row_utt, int pre_relu) {   // pre_relu: LN(relu(a + sum b))
  constexpr int PER = C / 64;                          // float2 per lane: one warp owns a row
  pdl_trigger();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x % 32;
  if (warp >= R) return;
  const float2* a2 = a ? reinterpret_cast<const float2*>(a + (size_t)warp * C) : nullptr;     // a may be null (= 0)
  const float2* b2 = b ? reinterpret_cast<const float2*>(b + (size_t)warp * C) : nullptr;
  float2* o2 = reinterpret_cast<float2*>(out + (size_t)warp * C);
  // optional second output: the row as planar fp16 hi / lo [C/8][R][8] (the next conv's operand, umma_split.cu); the float2 with
  // index k = lane + 32 i holds channels 2k, 2k+1, i.e. 4 bytes of plane k / 4
  auto planar = [&](int k) { return ((size_t)(k >> 2) * R + warp) * 8 + (size_t)(k & 3) * 2; };
  if (row_utt && row_utt[warp] < 0) {
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      o2[lane + 32 * i] = make_float2(0.f, 0.f);
      if (out_hi) {
        *reinterpret_cast<uint32_t*>(out_hi + planar(lane + 32 * i)) = 0u;
        *reinterpret_cast<uint32_t*>(out_lo + planar(lane + 32 * i)) = 0u;
      }
    }
    return;
  }
  float2 v[PER];
#pragma unroll
  for (int i = 0; i < PER; ++i) v[i] = a2 ? a2[lane + 32 * i] : make_float2(0.f, 0.f);
  if (b2) {
    for (int s = 0; s < n_b; ++s) {                      // fixed order: deterministic
      const float2* bs = b2 + (size_t)s * (b_stride / 2);
#pragma unroll
      for (int i = 0; i < PER; ++i) { const float2 w = bs[lane + 32 * i]; v[i].x += w.x; v[i].y += w.y; }
    }
  }
  if (pre_relu) {
#pragma unroll
    for (int i = 0; i < PER; ++i) { v[i].x = fmaxf(v[i].x, 0.f); v[i].y = fmaxf(v[i].y, 0.f); }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) s += v[i].x + v[i].y;
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) { const float dx = v[i].x - mean, dy = v[i].y - mean; q += dx * dx + dy * dy; }
#pragma unroll
  for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / C + 1e-5f);
  const float2* g2 = reinterpret_cast<const float2*>(gamma);
  const float2* be2 = reinterpret_cast<const float2*>(beta);
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const float2 g = g2[lane + 32 * i], be = be2[lane + 32 * i];
    const float2 y = make_float2((v[i].x - mean) * rstd * g.x + be.x, (v[i].y - mean) * rstd * g.y + be.y);
    o2[lane + 32 * i] = y;
    if (out_hi) {
      const __half2 h = __floats2half2_rn(y.x, y.y);
      const float2 hf = __half22float2(h);
      const __half2 l = __floats2half2_rn(y.x - hf.x, y.y - hf.y);
      *reinterpret_cast<__half2*>(out_hi + planar(lane + 32 * i)) = h;
      *reinterpret_cast<__half2*>(out_lo + planar(lane + 32 * i)) = l;
    }
  }
}

int layernorm_rows_ex(const float* a, const float* b, int n_b, int64_t b_stride, const float* gamma, const float* beta, float* out,
                      __half* out_hi, __half* out_lo, int R, int C, const int32_t* row_utt, cudaStream_t st, int pre_relu) {
  VS_REQUIRE(C == 192 || C == 256 || C == 768, "layernorm: C=%d unsupported (192, 256 or 768)", C);
  VS_REQUIRE(a || b, "layernorm: no input");
  VS_REQUIRE((out_hi == nullptr) == (out_lo == nullptr) && b_stride % 2 == 0, "layernorm: bad planar outputs / partial stride");
  const int warps_per_block = 8, grid = (R + warps_per_block - 1) / warps_per_block;
  if (C == 192) VS_CUDA_CHECK(launch_pdl(layernorm_rows_kernel<192>, dim3(grid), dim3(warps_per_block * 32), 0, st, a, b, n_b, b_stride, gamma, beta, out, out_hi, out_lo, R, row_utt, pre_relu));
  else if (C == 256) VS_CUDA_CHECK(launch_pdl(layernorm_rows_kernel<256>, dim3(grid), dim3(warps_per_block * 32), 0, st, a, b, n_b, b_stride, gamma, beta, out, out_hi, out_lo, R, row_utt, pre_relu));
  else VS_CUDA_CHECK(launch_pdl(layernorm_rows_kernel<768>, dim3(grid), dim3(warps_per_block * 32), 0, st, a, b, n_b, b_stride, gamma, beta, out, out_hi, out_lo, R, row_utt, pre_relu));
  VS_LAUNCH_CHECK();
  return VS_OK;
}
int layernorm_rows(const float* a, const float* b, const float* gamma, const float* beta, float* out, int R, int C,
                   const int32_t* row_utt, cudaStream_t st) {
  return layernorm_rows_ex(a, b, 1, 0, gamma, beta, out, nullptr, nullptr, R, C, row_utt, st);
}

// ------------------------------------------------------------------------------------------------
// Windowed relative-position self-attention (attentions.py:148-179), banded form, streaming softmax.
// One CTA = 64 queries of one (utterance, head).  qkv rows are [q(192) | k(192) | v(192)], head h owns
// channels [96h, 96h+96) of each.  scores = (q/sqrt(96)).k + (q/sqrt(96)).Ek[j-i+4] for |j-i|<=4;
// out = softmax(scores).(v) + sum_d p[i,i+d] Ev[d+4].  Keys are the utterance's own rows only, so the
// reference's -1e4 pad fill (attentions.py:166) never triggers (batch-1 semantics).
// ------------------------------------------------------------------------------------------------
constexpr int AT_BK = 64, AT_D = kHeadDim;
constexpr int AT_LD = 100;   // row pitch of Q/K/V tiles: 16 B aligned, 400 B = 4 banks apart -> LDS.128 conflict-free
constexpr int AT_SLD = 68;   // row pitch of the score tile (16 B aligned)

// RI = query rows per thread: 4 (64 queries per CTA) or 1 (16 per CTA: the batch-1 path, where 64-query CTAs leave a 70-phoneme call on
// 4 CTAs of serial work)
template <int RI>
__global__ void __launch_bounds__(256) rel_attention_kernel(VsRows rows, const float* __restrict__ qkv,
                                                            const float* __restrict__ ek,
                                                            const float* __restrict__ ev, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  constexpr int AT_BQ = 16 * RI;     // queries per CTA
  extern __shared__ __align__(16) float sm[];
  float* Qs = sm;                          // [BQ][100]
  float* Ks = Qs + AT_BQ * AT_LD;          // [64][100]
  float* Vs = Ks + AT_BK * AT_LD;          // [64][100]
  float* Ss = Vs + AT_BK * AT_LD;          // [BQ][68]
  float* Ev = Ss + AT_BQ * AT_SLD;         // [9][96]
  float* Ek = Ev + kRel * AT_D;            // [9][96]
  float* relq = Ek + kRel * AT_D;          // [BQ][9]
  float* row_m = relq + AT_BQ * kRel;      // [BQ]
  float* row_l = row_m + AT_BQ;            // [BQ]
  float* row_alpha = row_l + AT_BQ;        // [BQ]

  const int b = blockIdx.z, h = blockIdx.y;
  const int T = rows.utt_len[b], start = rows.utt_start[b];
  const int q0 = blockIdx.x * AT_BQ;
  if (q0 >= T) return;
  const int tid = threadIdx.x;
  const int ld = 3 * kHidden;
  const float scale = rsqrtf((float)AT_D);

  for (int i = tid; i < AT_BQ * (AT_D / 4); i += 256) {       // float4 loads: 24 per row
    const int r = i / (AT_D / 4), d4 = i % (AT_D / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + r < T) v = *reinterpret_cast<const float4*>(qkv + (size_t)(start + q0 + r) * ld + h * AT_D + 4 * d4);
    v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    *reinterpret_cast<float4*>(Qs + r * AT_LD + 4 * d4) = v;
  }
  for (int i = tid; i < kRel * AT_D; i += 256) { Ev[i] = ev[i]; Ek[i] = ek[i]; }
  if (tid < AT_BQ) { row_m[tid] = -INFINITY; row_l[tid] = 0.f; }
  __syncthreads();
  for (int i = tid; i < AT_BQ * kRel; i += 256) {
    const int r = i / kRel, w = i % kRel;
    float s0 = 0.f, s1 = 0.f;                  // both operands from shared memory, 16 bytes at a time (the scalar loop over global ek was ~1/4 of the kernel)
#pragma unroll 6
    for (int d = 0; d < AT_D; d += 4) {
      const float4 qv = *reinterpret_cast<const float4*>(Qs + r * AT_LD + d), kv = *reinterpret_cast<const float4*>(Ek + w * AT_D + d);
      s0 = fmaf(qv.x, kv.x, s0); s1 = fmaf(qv.y, kv.y, s1); s0 = fmaf(qv.z, kv.z, s0); s1 = fmaf(qv.w, kv.w, s1);
    }
    relq[i] = s0 + s1;
  }

  // S tile: rows RI*ty..+RI-1, key columns tx+16c (c<4).  O tile: the same rows, head dims 6tx..6tx+5.
  const int ty = tid / 16, tx = tid % 16;
  float o_acc[RI][6];
#pragma unroll
  for (int i = 0; i < RI; ++i)
#pragma unroll
    for (int c = 0; c < 6; ++c) o_acc[i][c] = 0.f;

  for (int k0 = 0; k0 < T; k0 += AT_BK) {
    __syncthreads();
    for (int i = tid; i < AT_BK * (AT_D / 4); i += 256) {
      const int r = i / (AT_D / 4), d4 = i % (AT_D / 4);
      float4 kk = make_float4(0.f, 0.f, 0.f, 0.f), vv = kk;
      if (k0 + r < T) {
        const float* g = qkv + (size_t)(start + k0 + r) * ld + h * AT_D + 4 * d4;
        kk = *reinterpret_cast<const float4*>(g + kHidden);
        vv = *reinterpret_cast<const float4*>(g + 2 * kHidden);
      }
      *reinterpret_cast<float4*>(Ks + r * AT_LD + 4 * d4) = kk;
      *reinterpret_cast<float4*>(Vs + r * AT_LD + 4 * d4) = vv;
    }
    __syncthreads();
    float s[RI][4];
#pragma unroll
    for (int i = 0; i < RI; ++i)
#pragma unroll
      for (int c = 0; c < 4; ++c) s[i][c] = 0.f;
#pragma unroll 2
    for (int d = 0; d < AT_D; d += 4) {            // 8 LDS.128 per 64 FMA
      float4 qa[RI], kb[4];
#pragma unroll
      for (int i = 0; i < RI; ++i) qa[i] = *reinterpret_cast<const float4*>(Qs + (RI * ty + i) * AT_LD + d);
#pragma unroll
      for (int c = 0; c < 4; ++c) kb[c] = *reinterpret_cast<const float4*>(Ks + (tx + 16 * c) * AT_LD + d);
#pragma unroll
      for (int i = 0; i < RI; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          s[i][c] = fmaf(qa[i].x, kb[c].x, s[i][c]);
          s[i][c] = fmaf(qa[i].y, kb[c].y, s[i][c]);
          s[i][c] = fmaf(qa[i].z, kb[c].z, s[i][c]);
          s[i][c] = fmaf(qa[i].w, kb[c].w, s[i][c]);
        }
    }
#pragma unroll
    for (int i = 0; i < RI; ++i)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int qi = q0 + RI * ty + i, kj = k0 + tx + 16 * c;
        float v = s[i][c];
        const int dd = kj - qi;
        if (dd >= -kWindow && dd <= kWindow) v += relq[(RI * ty + i) * kRel + dd + kWindow];
        if (kj >= T) v = -INFINITY;
        Ss[(RI * ty + i) * AT_SLD + tx + 16 * c] = v;
      }
    __syncthreads();
    // row-wise streaming softmax: 4 threads per row in BOTH variants (with 16 queries per CTA only two warps work here), so the row
    // sums associate the same way and an utterance's bits do not depend on which variant its batch selects
    if (tid < AT_BQ * 4) {
      const int r = tid / 4, part = tid % 4;
      float mx = -INFINITY;
      for (int c = part; c < AT_BK; c += 4) mx = fmaxf(mx, Ss[r * AT_SLD + c]);
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const float m_old = row_m[r];
      const float m_new = fmaxf(m_old, mx);
      float sum = 0.f;
      for (int c = part; c < AT_BK; c += 4) {
        const float p = __expf(Ss[r * AT_SLD + c] - m_new);
        Ss[r * AT_SLD + c] = p;
        sum += p;
      }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      __syncwarp();
      if (part == 0) {
        const float alpha = (m_old == -INFINITY) ? 0.f : __expf(m_old - m_new);
        row_alpha[r] = alpha;
        row_l[r] = row_l[r] * alpha + sum;
        row_m[r] = m_new;
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < RI; ++i) {
      const float al = row_alpha[RI * ty + i];
#pragma unroll
      for (int c = 0; c < 6; ++c) o_acc[i][c] *= al;
    }
#pragma unroll 2
    for (int j = 0; j < AT_BK; j += 4) {            // 4 LDS.128 (P) + 12 LDS.64 (V) per 96 FMA
      float4 p4[RI];
#pragma unroll
      for (int i = 0; i < RI; ++i) p4[i] = *reinterpret_cast<const float4*>(Ss + (RI * ty + i) * AT_SLD + j);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const float* vr = Vs + (j + jj) * AT_LD + 6 * tx;
        const float2 v0 = *reinterpret_cast<const float2*>(vr), v1 = *reinterpret_cast<const float2*>(vr + 2),
                     v2 = *reinterpret_cast<const float2*>(vr + 4);
#pragma unroll
        for (int i = 0; i < RI; ++i) {
          const float p = jj == 0 ? p4[i].x : jj == 1 ? p4[i].y : jj == 2 ? p4[i].z : p4[i].w;
          o_acc[i][0] = fmaf(p, v0.x, o_acc[i][0]); o_acc[i][1] = fmaf(p, v0.y, o_acc[i][1]);
          o_acc[i][2] = fmaf(p, v1.x, o_acc[i][2]); o_acc[i][3] = fmaf(p, v1.y, o_acc[i][3]);
          o_acc[i][4] = fmaf(p, v2.x, o_acc[i][4]); o_acc[i][5] = fmaf(p, v2.y, o_acc[i][5]);
        }
      }
    }
    // relative values on the band (attentions.py:174-177)
#pragma unroll
    for (int i = 0; i < RI; ++i) {
      const int qi = q0 + RI * ty + i;
      for (int dd = -kWindow; dd <= kWindow; ++dd) {
        const int j = qi + dd - k0;
        if (j < 0 || j >= AT_BK || qi + dd >= T) continue;
        const float p = Ss[(RI * ty + i) * AT_SLD + j];
#pragma unroll
        for (int c = 0; c < 6; ++c) o_acc[i][c] = fmaf(p, Ev[(dd + kWindow) * AT_D + 6 * tx + c], o_acc[i][c]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < RI; ++i) {
    const int qi = q0 + RI * ty + i;
    if (qi >= T) continue;
    const float inv = 1.f / row_l[RI * ty + i];
    float* o = out + (size_t)(start + qi) * kHidden + h * AT_D + 6 * tx;
#pragma unroll
    for (int c = 0; c < 6; c += 2) *reinterpret_cast<float2*>(o + c) = make_float2(o_acc[i][c] * inv, o_acc[i][c + 1] * inv);
  }
}

static size_t attention_smem_bytes(int bq) {
  return sizeof(float) * ((bq + 2 * AT_BK) * AT_LD + bq * AT_SLD + 2 * kRel * AT_D + bq * kRel + 3 * bq);
}

int rel_attention(const VsRows& rows, const float* qkv, const float* ek, const float* ev, float* out, cudaStream_t st,
                  Workspace* ws, __half* out_hi, __half* out_lo, bool* wrote_planar, bool gaps_dont_care) {
  if (wrote_planar) *wrote_planar = false;
  // option "attention_mma": 1 = auto (from 128 rows per utterance up: tcgen05 when the caller gave a workspace, else mma.sync),
  // 0 = CUDA cores, 2 / 3 = always mma.sync (3xTF32 / plain TF32), 4 = always tcgen05
  const int mode = (int)opts().v[OPT_ATTENTION_MMA];
  if (ws && (mode == 4 || (mode == 1 && rows.max_len >= 128 && rel_attention_umma_fits(rows, *ws)))) {
    Workspace scratch = *ws;                       // a copy: the caller's allocations stay where they are
    const bool planar = out_hi && out_lo && wrote_planar;
    if (planar) *wrote_planar = true;
    return rel_attention_umma(rows, qkv, ek, ev, out, scratch, st, planar ? out_hi : nullptr, planar ? out_lo : nullptr);
  }
  if (mode >= 2 || (mode == 1 && rows.max_len >= 128)) return rel_attention_mma(rows, qkv, ek, ev, out, st);
  int n_sm = 0;
  VS_TRY(device_sm_count(&n_sm));
  // too few 64-query CTAs for half the SMs (the batch-1 latency path): 16 queries per CTA
  const bool small = opts().v[OPT_ATTENTION_SMALL] && ((rows.max_len + 63) / 64) * kHeads * rows.n_utt * 2 <= n_sm;
  const int bq = small ? 16 : 64;
  const size_t smem = attention_smem_bytes(bq);
  VS_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(small ? rel_attention_kernel<1> : rel_attention_kernel<4>), (int)smem));
  // gap rows of `out` must be zero: the caller feeds out into a k=1 conv whose epilogue masks, but keep it clean
  if (!gaps_dont_care) VS_CUDA_CHECK(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)rows.n_rows * kHidden, st));
  // grid.x covers the longest utterance; CTAs beyond an utterance's length exit immediately
  VS_REQUIRE(rows.max_len > 0 && rows.max_len <= rows.n_rows, "rel_attention: bad max_len %d", rows.max_len);
  dim3 grid((rows.max_len + bq - 1) / bq, kHeads, rows.n_utt);
  if (small) VS_CUDA_CHECK(launch_pdl<16>(rel_attention_kernel<1>, dim3(grid), dim3(256), smem, st, rows, qkv, ek, ev, out));
  else VS_CUDA_CHECK(launch_pdl<16>(rel_attention_kernel<4>, dim3(grid), dim3(256), smem, st, rows, qkv, ek, ev, out));
  VS_LAUNCH_CHECK();
  return VS_OK;
}

// ------------------------------------------------------------------------------------------------
// out[r] = bias + w . x[r]   (the 1-channel projections: duration proj, proj_f0, energy linear)
// ------------------------------------------------------------------------------------------------
__global__ void row_dot_kernel(const float* __restrict__ x, int ld, const float* __restrict__ w,
                               const float* __restrict__ bias, float* __restrict__ out, int R, int C,
                               const int32_t* __restrict__ row_utt) {
  pdl_trigger();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x % 32;
  if (warp >= R) return;
  float s = 0.f;
  const bool valid = !row_utt || row_utt[warp] >= 0;
  if (valid)
    for (int c = lane; c < C; c += 32) s = fmaf(x[(size_t)warp * ld + c], w[c], s);
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[warp] = valid ? s + bias[0] : 0.f;
}

int row_dot(const float* x, int ld, const float* w, const float* bias, float* out, int R, int C,
            const int32_t* row_utt, cudaStream_t st) {
  VS_CUDA_CHECK(launch_pdl<16>(row_dot_kernel, dim3((R + 7) / 8), dim3(256), 0, st, x, ld, w, bias, out, R, C, row_utt));
  VS_LAUNCH_CHECK();
  return VS_OK;
}

}  // namespace vs
