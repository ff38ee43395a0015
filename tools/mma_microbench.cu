// tcgen05.mma rate microbenchmark (sm_100a).  One issuing warp, one accumulator chain per CTA, operands are
// uninitialised shared memory - only the timing matters.  Two questions:
//   1. cycles per MMA (M=128, K=16, f16) as a function of N with fixed operand addresses (tensor / smem floors);
//   2. the same with the conv kernel's real descriptor walk: A start offset = tap*dil rows (16 B each) + K-chunk
//      planes `rows_a` rows apart (LBO), B walking through contiguous weight slabs.
// Build: nvcc -O2 -gencode arch=compute_100a,code=sm_100a -I vispeech_b200/csrc -o tools/_bin/mma_mb tools/mma_microbench.cu
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <string>
#include "umma_common.cuh"

namespace vs { void set_error(const char*, ...) {} unsigned long long g_launch_count = 0; }
using namespace vs::umma;

struct Walk { int n, taps, dil, nks, rows_a, tiles, iters, b_fixed, a_fixed, variant, fill, noise; float* gbuf; };

// NK MMAs of one conv tap in ONE asm block: a single elect, descriptor low words advanced inside the block.
template <int NK>
__device__ __forceinline__ void mma_tap_block(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                              uint32_t idesc, uint32_t accumulate, uint32_t a_kstep, uint32_t b_kstep) {
  asm volatile(
      "{\n\t.reg .pred p, pe;\n\t.reg .b64 da, db;\n\t.reg .b32 al, bl;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b32 al, %1;\n\t"
      "mov.b32 bl, %3;\n\t"
      "mov.b64 da, {al, %2};\n\t"
      "mov.b64 db, {bl, %4};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "setp.ne.b32 p, 1, 0;\n\t"
      ".pragma \"nounroll\";\n\t"
      "}" ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
#pragma unroll
  for (int i = 1; i < NK; ++i) {
    a_lo += a_kstep;
    b_lo += b_kstep;
    asm volatile(
        "{\n\t.reg .pred pe;\n\t.reg .b64 da, db;\n\t"
        "elect.sync _|pe, 0xffffffff;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, 1;\n\t"
        "}" ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc)
        : "memory");
  }
}
// plain MMA without election: the caller is already a single lane
__device__ __forceinline__ void mma_plain(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
template <int NK>
__device__ __forceinline__ void walk_unrolled(const Walk& w, uint32_t tm, uint32_t a_lo0, uint32_t a_hi, uint32_t b_lo0, uint32_t b_hi,
                                              uint32_t idesc, uint32_t a_kstep, uint32_t b_kstep, uint32_t dil) {
  const int taps = w.taps, tiles = w.tiles, iters = w.iters;
  for (int it = 0; it < iters; ++it)
    for (int m = 0; m < tiles; ++m) {
      uint32_t a_tap = a_lo0 + (uint32_t)(m * 128), b_lo = b_lo0, accumulate = 0;
#pragma unroll 2
      for (int t = 0; t < taps; ++t, a_tap += dil) {
        uint32_t a_lo = a_tap;
#pragma unroll
        for (int ks = 0; ks < NK; ++ks) {
          tc_mma_f16_lohi(tm, a_lo, a_hi, b_lo, b_hi, idesc, accumulate);
          accumulate = 1;
          a_lo += a_kstep;
          b_lo += b_kstep;
        }
      }
    }
}
template <int NK>
__device__ __forceinline__ void walk_lane0(const Walk& w, uint32_t tm, uint32_t a_lo0, uint32_t a_hi, uint32_t b_lo0, uint32_t b_hi,
                                           uint32_t idesc, uint32_t a_kstep, uint32_t b_kstep, uint32_t dil) {
  const int taps = w.taps, tiles = w.tiles, iters = w.iters;
  for (int it = 0; it < iters; ++it)
    for (int m = 0; m < tiles; ++m) {
      uint32_t a_tap = a_lo0 + (uint32_t)(m * 128), b_lo = b_lo0, accumulate = 0;
#pragma unroll 2
      for (int t = 0; t < taps; ++t, a_tap += dil) {
        uint32_t a_lo = a_tap;
#pragma unroll
        for (int ks = 0; ks < NK; ++ks) {
          mma_plain(tm, a_lo, a_hi, b_lo, b_hi, idesc, accumulate);
          accumulate = 1;
          a_lo += a_kstep;
          b_lo += b_kstep;
        }
      }
    }
}

__global__ void __launch_bounds__(320, 2) mma_walk(const Walk w, long long* out_clk) {
  __shared__ volatile int done_flag;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(8) uint64_t bar2;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) { done_flag = 0; mbar_init(smem_u32(&bar), 1); mbar_init(smem_u32(&bar2), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (w.fill) {   // operands: 0 = whatever is there, 1 = zeros, 2 = pseudo-random f16 in [-1, 1)
    uint32_t* s32 = reinterpret_cast<uint32_t*>(smem);
    uint32_t x = 1234567u + threadIdx.x * 7919u + blockIdx.x * 104729u;
    for (int i = threadIdx.x; i < 190 * 256; i += blockDim.x) {
      x = x * 1664525u + 1013904223u;
      const uint32_t lo = 0x3F000000u | ((x >> 9) & 0x007F0000u) | (x & 0x80000000u);      // +-[0.5,1) f16 in the high half
      const uint32_t hi = 0x3F000000u | ((x << 3) & 0x007F0000u) | ((x << 7) & 0x80000000u);
      s32[i] = w.fill == 1 ? 0u : ((lo >> 16) | (hi & 0xFFFF0000u));
    }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_slot;
  if (warp == 1) {
    const uint32_t base = smem_u32(smem);
    const uint32_t a_lbo = (uint32_t)w.rows_a * 16u, b_lbo = (uint32_t)w.n * 16u;
    const uint32_t a_bytes = (uint32_t)w.rows_a * 16u * 2u * (uint32_t)w.nks;
    const uint64_t ad = make_desc(base, a_lbo, 128), bd = make_desc(base + ((a_bytes + 1023u) & ~1023u), b_lbo, 128);
    const uint32_t a_hi = (uint32_t)(ad >> 32), b_hi = (uint32_t)(bd >> 32);
    const uint32_t a_lo0 = (uint32_t)ad, b_lo0 = (uint32_t)bd;
    const uint32_t idesc = make_idesc(w.n);
    const uint32_t a_kstep = w.a_fixed ? 0u : 2u * (uint32_t)w.rows_a, b_kstep = w.b_fixed ? 0u : 2u * (uint32_t)w.n;
    const int taps = w.taps, nks = w.nks, tiles = w.tiles, iters = w.iters;
    const uint32_t dil = w.a_fixed ? 0u : (uint32_t)w.dil;
    const long long t0 = clock64();
    if (w.variant == 1) {
      if (nks == 2) walk_unrolled<2>(w, tm, a_lo0, a_hi, b_lo0, b_hi, idesc, a_kstep, b_kstep, dil);
      else if (nks == 4) walk_unrolled<4>(w, tm, a_lo0, a_hi, b_lo0, b_hi, idesc, a_kstep, b_kstep, dil);
      else walk_unrolled<8>(w, tm, a_lo0, a_hi, b_lo0, b_hi, idesc, a_kstep, b_kstep, dil);
    } else if (w.variant == 2) {
      if ((threadIdx.x & 31) == 0) {
        if (nks == 2) walk_lane0<2>(w, tm, a_lo0, a_hi, b_lo0, b_hi, idesc, a_kstep, b_kstep, dil);
        else if (nks == 4) walk_lane0<4>(w, tm, a_lo0, a_hi, b_lo0, b_hi, idesc, a_kstep, b_kstep, dil);
        else walk_lane0<8>(w, tm, a_lo0, a_hi, b_lo0, b_hi, idesc, a_kstep, b_kstep, dil);
      }
      __syncwarp();
    } else
    for (int it = 0; it < iters; ++it)
      for (int m = 0; m < tiles; ++m) {
        uint32_t a_tap = a_lo0 + (uint32_t)(m * 128), b_lo = b_lo0, accumulate = 0;
        for (int t = 0; t < taps; ++t, a_tap += dil) {
          uint32_t a_lo = a_tap;
          for (int ks = 0; ks < nks; ++ks) {
            tc_mma_f16_lohi(tm, a_lo, a_hi, b_lo, b_hi, idesc, accumulate);
            accumulate = 1;
            a_lo += a_kstep;
            b_lo += b_kstep;
          }
        }
      }
    tc_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0, 1);
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) { out_clk[blockIdx.x] = t1 - t0; done_flag = 1; }
  } else if (warp >= 2 && w.noise) {
    // background traffic from 4 "epilogue" warps while the MMAs run: 1 = tcgen05.ld (TMEM reads), 2 = LDS, 4 = STS, 8 = STG
    const int lane = threadIdx.x & 31, q = warp & 3;
    uint32_t v[32];
    float acc = 0.f;
    const uint32_t sbase = smem_u32(smem) + 150 * 1024 + (uint32_t)((threadIdx.x - 64) & 127) * 16u;
    long long n_it = 0;
    const long long tn0 = clock64();
    while (!done_flag) {
      ++n_it;
      if (w.noise & 1) { tmem_ld32(tm + ((uint32_t)(q * 32) << 16) + 128u, v); acc += __uint_as_float(v[lane & 31]); }
      if (w.noise & 2) {
        uint32_t x0, x1, x2, x3;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3) : "r"(sbase + r * 2048u));
          acc += __uint_as_float(x0 ^ x3);
        }
      }
      if (w.noise & 4) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
          asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(sbase + r * 2048u), "r"(lane) : "memory");
      }
      if (w.noise & 16) {    // spin on an mbarrier that never completes (what a waiting epilogue warp does)
#pragma unroll 1
        for (int r = 0; r < 64; ++r)
          if (mbar_try_wait(smem_u32(&bar2), 0)) acc += 1.f;
      }
      if (w.noise & 32) {    // TMEM writes
        tmem_ld32(tm + ((uint32_t)(q * 32) << 16) + 128u, v);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(tm + ((uint32_t)(q * 32) << 16) + 192u),
                     "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      if (w.noise & 8) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
          reinterpret_cast<float4*>(w.gbuf)[((size_t)blockIdx.x * 4 + r) * 128 + ((threadIdx.x - 64) & 127)] = make_float4(acc, 0, 0, 0);
      }
    }
    if (lane == 0 && warp == 2) out_clk[300 + blockIdx.x] = (clock64() - tn0) / (n_it > 0 ? n_it : 1);
    if (acc == 123.456f) w.gbuf[0] = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256u) : "memory");
}

static long long* d_clk;
static float* gbuf0() { static float* g = nullptr; if (!g) cudaMalloc(&g, (size_t)300 * 4 * 128 * 16); return g; }
static double g_noise_clk = 0;   // clocks per iteration of the background loop (CTA 0, warp 2)
static double run(Walk w, int per_sm) {
  static long long h[512];
  const int smem = per_sm == 1 ? 200 * 1024 : 100 * 1024;
  mma_walk<<<148 * per_sm, w.noise ? (w.noise & 64 ? 320 : 192) : 64, smem>>>(w, d_clk);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  cudaMemcpy(h, d_clk, sizeof(long long) * 512, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148 * per_sm; ++i) avg += (double)h[i];
  avg /= 148 * per_sm;
  g_noise_clk = (double)h[300];
  return avg / ((double)w.iters * w.tiles * w.taps * w.nks) / per_sm;   // clocks per MMA per SM
}

int main(int argc, char** argv) {
  cudaMalloc(&d_clk, 512 * sizeof(long long));
  cudaFuncSetAttribute(mma_walk, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (argc > 1 && std::string(argv[1]) == "pitch") {
    // 3. does the pitch between the two K-chunk planes of one MMA (LBO = rows_a * 16 B) matter?  (unrolled issue loop, 1 CTA/SM)
    printf("plane pitch sweep, N = C, k = 7, d = 1: rows_a (pitch mod 128 B) -> clk/MMA/SM\n");
    for (int c : {32, 64})
      for (int rows_a : {262, 256, 258, 260, 264, 832, 833, 834, 836, 838, 840, 784, 786}) {
        if ((size_t)rows_a * c * 2 > 150 * 1024) continue;
        Walk w{c, 7, 1, c / 16, rows_a, 2, 100, 0, 0, 1, 0, 0, nullptr};
        printf("  C=%3d rows_a=%4d (%3d) : %6.1f\n", c, rows_a, (rows_a * 16) % 128, run(w, 1));
      }
    printf("background warps (C = 32, k = 7, unrolled issue loop, 1 CTA/SM): clk/MMA/SM with 4 / 8 background warps\n");
    const char* names[] = {"none", "tcgen05.ld", "LDS", "STS", "STG", "mbarrier spin", "tcgen05.ld+st", "spin+STS+ld"};
    const int modes[] = {0, 1, 2, 4, 8, 16, 32, 16 | 4 | 1};
    for (int i = 0; i < 8; ++i) {
      Walk w{32, 7, 1, 2, 262, 2, 100, 0, 0, 1, 0, modes[i], gbuf0()};
      Walk w8 = w; w8.noise |= modes[i] ? 64 : 0;
      printf("  %-14s : %6.1f  %6.1f\n", names[i], run(w, 1), run(w8, 1));
    }
    return 0;
  }
  printf("fixed operands (floors): clk/MMA/SM at 1 and 2 CTAs/SM\n");
  for (int n : {32, 64, 128, 256}) {
    Walk w{n, 8, 1, 4, 256, 1, 200, 1, 1, 0, 0, 0, nullptr};
    printf("  N=%3d : %6.1f %6.1f\n", n, run(w, 1), n <= 128 ? run(w, 2) : 0.0);
  }
  printf("operand data (unrolled issue loop, 1 CTA/SM, 400 tiles): as found | zeros | random\n");
  for (int c : {32, 64}) {
    Walk w{c, 7, 1, c / 16, 128 * 2 + 6, 2, 200, 0, 0, 1, 0, 0, nullptr};
    Walk wz = w; wz.fill = 1;
    Walk wr = w; wr.fill = 2;
    printf("  C=%3d k=7 : %6.1f | %6.1f | %6.1f\n", c, run(w, 1), run(wz, 1), run(wr, 1));
  }
  float* gbuf;
  cudaMalloc(&gbuf, (size_t)300 * 4 * 128 * 16);
  printf("background traffic from 4 other warps (unrolled issue loop, 1 CTA/SM): none | tcgen05.ld | LDS | STS | STG | all\n");
  for (int c : {32, 64}) {
    printf("  C=%3d k=7 :", c);
    for (int noise : {0, 1, 2, 4, 8, 15}) {
      Walk w{c, 7, 1, c / 16, 128 * 2 + 6, 2, 200, 0, 0, 1, 0, noise, gbuf};
      const double r = run(w, 1);
      printf(" %6.1f (bg loop %5.0f clk)", r, noise ? g_noise_clk : 0.0);
    }
    printf("\n");
  }
  printf("conv descriptor walk, 1 CTA/SM: N=C, taps, dil, MT -> clk/MMA/SM issue-loop variants\n");
  for (int c : {32, 64, 128})
    for (int taps : {3, 7, 11})
      for (int dil : {1, 5}) {
        const int mt = c == 128 ? 1 : 2;
        const int rows_a = 128 * mt + (taps - 1) * dil;
        if ((size_t)rows_a * c * 2 + (size_t)taps * c * c * 2 > 90 * 1024) continue;
        Walk w{c, taps, dil, c / 16, rows_a, mt, 40, 0, 0, 0, 0, 0, nullptr};
        Walk w1 = w; w1.variant = 1;
        Walk w2 = w; w2.variant = 2;
        printf("  C=%3d k=%2d d=%d MT=%d : nested %6.1f | unrolled %6.1f | lane0 %6.1f   (2 CTAs/SM: %6.1f %6.1f %6.1f)\n", c, taps, dil, mt,
               run(w, 1), run(w1, 1), run(w2, 1), c <= 64 ? run(w, 2) : 0., c <= 64 ? run(w1, 2) : 0., c <= 64 ? run(w2, 2) : 0.);
      }
  return 0;
}
