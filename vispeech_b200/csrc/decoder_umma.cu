// f16 tcgen05 HiFi-GAN decoder: Generator.forward (reference models.py:271-290) as a chain of
// umma_conv1d launches over planar f16 activations.  All leaky-relus, biases, residual adds, the MRF
// sum (/3) and the ConvTranspose1d phase scatter are fused into conv epilogues; no elementwise pass
// touches HBM between convs.
#include "decoder.cuh"
#include "ops_misc.cuh"
#include "umma_conv.cuh"

namespace vs {

int ups_taps(int i, int* pad_l) {
  const int s = kUpRate[i], K = kUpKernel[i], p = kUpPad[i];
  int dmax = -1000, dmin = 1000;
  for (int ph = 0; ph < s; ++ph) {
    const int hi = (ph + p) / s, lo = hi - K / s + 1;
    dmax = hi > dmax ? hi : dmax;
    dmin = lo < dmin ? lo : dmin;
  }
  *pad_l = -dmin;
  return dmax - dmin + 1;
}

int resolve_decoder_f16(const FetchFn& fetch, DecoderW* d) {
#define DF16(field, name, numel) VS_TRY(fetch(name, numel, VS_DTYPE_F16, reinterpret_cast<const void**>(&(field))))
  DF16(d->pre16.w, "dec16.pre.w", 7 * 192 * 512);
  d->pre16.b = d->pre.b;
  for (int i = 0; i < kDecStages; ++i) {
    const int cin = kStageC[i], cout = kStageC[i + 1];
    int pad_l = 0;
    const int taps = ups_taps(i, &pad_l);
    DF16(d->ups16[i].w, "dec16.ups." + std::to_string(i) + ".w", (int64_t)taps * cin * cout * kUpRate[i]);
    d->ups16[i].b = d->ups[i].b;
    for (int j = 0; j < kDecKernels; ++j) {
      const int n = i * kDecKernels + j;
      for (int mth = 0; mth < kDecDils; ++mth) {
        const std::string q = "dec16.rb." + std::to_string(n) + ".";
        const int64_t numel = (int64_t)kResK[j] * cout * cout;
        DF16(d->c1_16[n][mth].w, q + "c1." + std::to_string(mth) + ".w", numel);
        DF16(d->c2_16[n][mth].w, q + "c2." + std::to_string(mth) + ".w", numel);
        d->c1_16[n][mth].b = d->c1[n][mth].b;
        d->c2_16[n][mth].b = d->c2[n][mth].b;
        if (cout <= 128) {   // parameters of the fused kernels (umma_respair.cu, umma_pairfused.cu, umma_resblock.cu, umma_mrf.cu)
          VS_CUDA_CHECK(cudaMemcpy(d->bias_host[n][mth][0], d->c1[n][mth].b, cout * sizeof(float), cudaMemcpyDeviceToHost));
          VS_CUDA_CHECK(cudaMemcpy(d->bias_host[n][mth][1], d->c2[n][mth].b, cout * sizeof(float), cudaMemcpyDeviceToHost));
        }
      }
    }
  }
#undef DF16
  VS_CUDA_CHECK(cudaMemcpy(d->post_w_host, d->post_w, sizeof(d->post_w_host), cudaMemcpyDeviceToHost));
  return VS_OK;
}

// fp32 row-major [R][C] -> planar f16 [C/8][R][8], zero on invalid rows
__global__ void to_planar_f16_kernel(const float* __restrict__ x, const int32_t* __restrict__ row_utt,
                                      __half* __restrict__ out, int R, int C) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;      // one thread per (plane, row)
  if (i >= (C / 8) * R) return;
  const int pl = i / R, r = i % R;
  uint4 o = make_uint4(0, 0, 0, 0);
  if (row_utt[r] >= 0) {
    const float4 a = *reinterpret_cast<const float4*>(x + (size_t)r * C + pl * 8);
    const float4 b = *reinterpret_cast<const float4*>(x + (size_t)r * C + pl * 8 + 4);
    __half2 h0 = __floats2half2_rn(a.x, a.y), h1 = __floats2half2_rn(a.z, a.w);
    __half2 h2 = __floats2half2_rn(b.x, b.y), h3 = __floats2half2_rn(b.z, b.w);
    o = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                   *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
  }
  *reinterpret_cast<uint4*>(out + (size_t)i * 8) = o;
}

// conv_post (32 -> 1, k7, no bias) + tanh on planar f16 input that already carries leaky_relu(.,0.01)
// (models.py:286-288).  224 MACs per sample: CUDA cores, one thread per output sample.
__global__ void conv_post_kernel(const __half* __restrict__ x, const float* __restrict__ w,
                                 const int32_t* __restrict__ row_utt, int row_div, float* __restrict__ wave, int R) {
  __shared__ float ws[7 * 32];
  for (int i = threadIdx.x; i < 7 * 32; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  if (row_utt[r / row_div] < 0) { wave[r] = 0.f; return; }
  float acc = 0.f;
#pragma unroll
  for (int t = 0; t < 7; ++t) {
    const int rr = r + t - 3;
    if (rr < 0 || rr >= R) continue;
#pragma unroll
    for (int pl = 0; pl < 4; ++pl) {
      const uint4 u = *reinterpret_cast<const uint4*>(x + ((size_t)pl * R + rr) * 8);
      const uint32_t wd[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&wd[e]));
        acc = fmaf(f.x, ws[t * 32 + pl * 8 + 2 * e], acc);
        acc = fmaf(f.y, ws[t * 32 + pl * 8 + 2 * e + 1], acc);
      }
    }
  }
  wave[r] = tanhf(acc);
}

// The three ResBlock chains of an MRF stage (k = 3, 7, 11) are independent until their sums meet (models.py:280-284).
// With two streams the k = 11 chain runs next to k = 3 then k = 7: every kernel here is a persistent full-grid launch, so
// nothing runs "in parallel", but the tail of each launch (SMs idle while the last tiles finish) is filled by the other
// chain's CTAs - 3-5 % per pair of launches in isolation (tools/cosched_pairs.py), but nothing inside the decoder (measured
// 25.77 vs 25.74 ms per step), so it is OFF by default.  Fork / join through events; the side chain has its own intermediate
// buffers; the sum keeps its order (k3 + k7) + k11.
// options "mrf_fused" (0 | 1: the last MRF stage + conv_post as ONE kernel, umma_mrf.cu) and "decoder_streams" (1 | 2; in situ 2
// gains nothing at batch size: 25.77 vs 25.74 ms) come from opts().  Small calls are different: below ~2048 frame rows the C = 256
// stage's kernels are 24 - 96 CTAs each and the two chains really run side by side (chunked 60 s decode 32.8 -> 28.1 ms, C1 2.78 ->
// 2.74), so decoder_streams = 0 (auto, the default) takes the side stream there and only there.
bool decoder_two_streams(int frame_rows) {
  const int64_t v = opts().v[OPT_DECODER_STREAMS];
  return v == 2 || (v == 0 && frame_rows < 2048);
}

struct SideStream { cudaStream_t s = nullptr; cudaEvent_t fork = nullptr, sum = nullptr, join = nullptr; };
static int side_stream(SideStream** out) {
  static thread_local SideStream tab[16];        // per thread: two threads decoding on one device must not share the fork / join events
  int dev = 0;
  VS_CUDA_CHECK(cudaGetDevice(&dev));
  VS_REQUIRE(dev >= 0 && dev < 16, "decode: device index %d", dev);
  SideStream& t = tab[dev];
  if (!t.s) {
    VS_CUDA_CHECK(cudaStreamCreateWithFlags(&t.s, cudaStreamNonBlocking));
    VS_CUDA_CHECK(cudaEventCreateWithFlags(&t.fork, cudaEventDisableTiming));
    VS_CUDA_CHECK(cudaEventCreateWithFlags(&t.sum, cudaEventDisableTiming));
    VS_CUDA_CHECK(cudaEventCreateWithFlags(&t.join, cudaEventDisableTiming));
  }
  *out = &t;
  return VS_OK;
}

int decode_f16(const DecoderW& w, const VsRows& rows, const float* z, int max_len, float* wave, Workspace& ws,
                cudaStream_t st_main) {
  cudaStream_t st = st_main;
  const int R = rows.n_rows;
  // small calls (the batch-1 latency path) are launch-bound: let every decoder kernel's launch and prologue overlap its predecessor's
  // tail (C1: 3.80 -> 3.67 ms per call).  Large batches keep the decoder kernels serialised: no gain device-resident, and in
  // throughput mode early-resident decoder CTAs get in the way of the other stream's latent stages (e2e 10.5 k -> 10.25 k).
  PdlExtra pdl_small(R < 2048 ? 4 : 0);
  int32_t* valid = ws.take<int32_t>(R);
  __half* zin = ws.take<__half>((int64_t)R * kHidden);
  const bool two = decoder_two_streams(R);
  const bool mrf_fused = opts().v[OPT_MRF_FUSED] != 0;
  __half* buf[9];
  for (int i = 0; i < (two ? 9 : 6); ++i) buf[i] = ws.take<__half>((int64_t)R * 16384);
  if (!ws.ok) { set_error("decode_f16: workspace too small"); return VS_ERR_WORKSPACE; }
  __half *XA = buf[0], *S = buf[4], *NEXT = buf[5];     // buf[1..3] / buf[6..8]: T, AA, BA of the main / side chain
  SideStream* side = nullptr;
  if (two) VS_TRY(side_stream(&side));

  VS_TRY(mask_frames(rows, max_len, valid, st));                         // (z * x_mask)[:, :, :max_len]  models.py:720
  VS_CUDA_CHECK(launch_pdl<4>(to_planar_f16_kernel, dim3(((kHidden / 8) * R + 255) / 256), dim3(256), 0, st, z, valid, zin, R, kHidden));
  VS_LAUNCH_CHECK();

  UmmaConv c;
  c.in = zin; c.w = w.pre16.w; c.bias = w.pre16.b; c.ubias = w.cond_tab; c.ubias_idx = rows.sid;
  c.out_act = NEXT; c.act_slope = 0.1f; c.row_utt = valid; c.row_div = 1; c.R = R; c.Cin = kHidden; c.N = 512;
  c.taps = 7; c.pad_l = 3;
  VS_TRY(umma_conv1d(c, st));                                            // lrelu(conv_pre(z) + cond(g))  models.py:272-276

  int mul = 1;
  for (int i = 0; i < kDecStages; ++i) {
    const int cin = kStageC[i], cout = kStageC[i + 1], s = kUpRate[i];
    int pad_l = 0;
    const int taps = ups_taps(i, &pad_l);
    // which resblocks of this stage run as fused conv pairs (umma_respair.cu)?
    bool fused[kDecKernels];
    for (int j = 0; j < kDecKernels; ++j) {
      fused[j] = true;
      for (int mth = 0; mth < kDecDils; ++mth) fused[j] = fused[j] && umma_respair_supported(cout, kResK[j], kResD[mth]);
    }
    c = UmmaConv();
    // only the ACTIVATED stream a = lrelu(x) is stored between ResBlock iterations: it is the next conv's operand as
    // is, and the residual x is recovered in the c2 epilogue as min(a, a/slope) (same f16 relative rounding as storing x)
    const bool mrf = mrf_fused && i == kDecStages - 1;                 // last stage: x0 in 22 bits (hi + lo), then umma_mrf
    c.in = NEXT; c.w = w.ups16[i].w; c.bias = w.ups16[i].b; c.out_raw = nullptr; c.out_act = XA;
    c.act_slope = 0.1f;
    if (mrf) { c.out_act = nullptr; c.out_raw = XA; c.out_lo = buf[1]; }
    c.row_utt = valid; c.row_div = mul * s; c.R = R * mul; c.Cin = cin; c.N = cout * s; c.taps = taps; c.pad_l = pad_l;
    c.up = s;
    VS_TRY(umma_conv1d(c, st));                                          // ups[i] (ConvTranspose1d)  models.py:277
    mul *= s;
    const int Rs = R * mul;
    if (mrf) {
      // 3 x ResBlock1 + sum / 3 + leaky_relu(0.01) + conv_post + tanh  (models.py:279-288): residual stream, MRF sum and
      // conv_post input stay in fp32 on chip
      UmmaMrf f;
      f.x_hi = XA; f.x_lo = buf[1]; f.wave = wave; f.row_utt = valid; f.row_div = mul; f.R = Rs; f.post_w_host = w.post_w_host;
      for (int j = 0; j < kDecKernels; ++j)
        for (int mth = 0; mth < kDecDils; ++mth) {
          const int n = i * kDecKernels + j;
          f.w[j][mth][0] = w.c1_16[n][mth].w; f.w[j][mth][1] = w.c2_16[n][mth].w;
          f.b1_host[j][mth] = w.bias_host[n][mth][0]; f.b2_host[j][mth] = w.bias_host[n][mth][1];
        }
      return umma_mrf(f, st_main);
    }
    if (two) {                                                           // fork: the side chain may start once XA is there
      VS_CUDA_CHECK(cudaEventRecord(side->fork, st_main));
      VS_CUDA_CHECK(cudaStreamWaitEvent(side->s, side->fork, 0));
    }
    for (int j = 0; j < kDecKernels; ++j) {
      const int n = i * kDecKernels + j, k = kResK[j];
      const float final_slope = (i == kDecStages - 1) ? 0.01f : 0.1f;    // final lrelu uses the default slope (Q3)
      const bool on_side = two && j == kDecKernels - 1;                  // the k = 11 chain
      st = on_side ? side->s : st_main;
      __half* T = on_side ? buf[6] : buf[1];
      __half* AA = on_side ? buf[7] : buf[2];
      __half* BA = on_side ? buf[8] : buf[3];
      if (two && j == kDecKernels - 1) VS_CUDA_CHECK(cudaEventRecord(side->sum, st_main));   // S = k3 + k7 is enqueued
      if (!on_side && opts().v[OPT_RESBLOCK_FUSED] && umma_resblock_supported(cout, k) && j == 0) {
        // the whole k = 3 ResBlock of the C = 64 stage as one kernel: residual stream in fp32 in TMEM (umma_resblock.cu)
        UmmaResBlock rb;
        rb.a = XA; rb.out_raw = S; rb.row_utt = valid; rb.row_div = mul; rb.R = Rs;
        for (int mth = 0; mth < kDecDils; ++mth) {
          rb.w[mth][0] = w.c1_16[n][mth].w; rb.w[mth][1] = w.c2_16[n][mth].w;
          rb.b1_host[mth] = w.bias_host[n][mth][0]; rb.b2_host[mth] = w.bias_host[n][mth][1];
        }
        VS_TRY(umma_resblock(rb, st));
        continue;
      }
      if (fused[j]) {
        const __half* cur = XA;                                   // a = lrelu(x): the only stream between iterations
        for (int mth = 0; mth < kDecDils; ++mth) {
          const bool last = (mth == kDecDils - 1);
          if (on_side && last) VS_CUDA_CHECK(cudaStreamWaitEvent(side->s, side->sum, 0));   // its epilogue reads S
          UmmaPair pr;
          pr.x = cur; pr.w1 = w.c1_16[n][mth].w; pr.b1 = w.c1_16[n][mth].b; pr.w2 = w.c2_16[n][mth].w; pr.b2 = w.c2_16[n][mth].b;
          pr.b1_host = w.bias_host[n][mth][0]; pr.b2_host = w.bias_host[n][mth][1];
          pr.row_utt = valid; pr.row_div = mul; pr.R = Rs; pr.C = cout; pr.taps = k; pr.dil = kResD[mth]; pr.in_slope = 0.1f;
          if (!last) { pr.out_act = (mth == 0) ? AA : BA; pr.act_slope = 0.1f; }   // lrelu(c2(lrelu(c1(a))) + x)  modules.py:211-220
          else {
            pr.res2 = (j > 0) ? S : nullptr;                             // xs += resblock_j(x)  models.py:280-284
            if (j < kDecKernels - 1) pr.out_raw = S;
            else { pr.out_act = NEXT; pr.act_scale = 1.f / kDecKernels; pr.act_slope = final_slope; }
          }
          VS_TRY(umma_respair(pr, st));
          cur = (mth == 0) ? AA : BA;
        }
        continue;
      }
      const __half* cur_act = XA;
      for (int mth = 0; mth < kDecDils; ++mth) {
        const bool last = (mth == kDecDils - 1);
        if (on_side && last) VS_CUDA_CHECK(cudaStreamWaitEvent(side->s, side->sum, 0));     // c2's epilogue reads S
        c = UmmaConv();
        c.row_utt = valid; c.row_div = mul; c.R = Rs; c.Cin = cout; c.N = cout; c.taps = k; c.pad_l = (k - 1) / 2;
        c.in = cur_act; c.w = w.c1_16[n][mth].w; c.bias = w.c1_16[n][mth].b; c.dil = kResD[mth];
        c.out_act = T; c.act_slope = 0.1f;
        VS_TRY(umma_conv1d(c, st));                                      // lrelu(c1(lrelu(x)))  modules.py:211-218
        c.in = T; c.w = w.c2_16[n][mth].w; c.bias = w.c2_16[n][mth].b; c.dil = 1;
        c.res = cur_act; c.res_inv_slope = 10.f;                         // + x, recovered from lrelu(x, 0.1)
        if (!last) {
          c.out_act = (mth == 0) ? AA : BA;                              // store lrelu(c2(.) + x)  modules.py:219-220, 211
        } else {
          c.res2 = (j > 0) ? S : nullptr;                                // xs += resblock_j(x)  models.py:280-284
          if (j < kDecKernels - 1) { c.out_raw = S; c.out_act = nullptr; }
          else {                                                         // x = xs / 3, then the next stage's leaky_relu
            c.out_raw = nullptr; c.out_act = NEXT; c.act_scale = 1.f / kDecKernels; c.act_slope = final_slope;
          }
        }
        VS_TRY(umma_conv1d(c, st));
        cur_act = (mth == 0) ? AA : BA;
      }
    }
    st = st_main;
    if (two) {                                                           // join: NEXT (written by the side chain) is complete
      VS_CUDA_CHECK(cudaEventRecord(side->join, side->s));
      VS_CUDA_CHECK(cudaStreamWaitEvent(st_main, side->join, 0));
    }
  }
  const int Rw = R * mul;
  conv_post_kernel<<<(Rw + 255) / 256, 256, 0, st>>>(NEXT, w.post_w, valid, mul, wave, Rw);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

}  // namespace vs
