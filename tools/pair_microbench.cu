// tcgen05.mma.cta_group::2 issue-rate microbenchmark (sm_100a): what does a round of the CTA-pair kernels cost the tensor pipe?
// A 2-CTA cluster; the leader's warp 0 issues `rounds` rounds of `chains` accumulation chains (different TMEM accumulators) of `m`
// MMAs (M = 256, N, K = 16, f16; operands are uninitialised shared memory - only the timing matters), with `commits` multicast
// commits per round, and waits for the last one.  Prints clocks per round and per MMA.
// Build: nvcc -O2 -gencode arch=compute_100a,code=sm_100a -I vispeech_b200/csrc -o tools/_bin/pair_mb tools/pair_microbench.cu
#include <cstdio>
#include <cstdlib>
#include "umma_common.cuh"

namespace vs { void set_error(const char*, ...) {} unsigned long long g_launch_count = 0; }
using namespace vs::umma;

__device__ __forceinline__ void commit_pair(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\t.reg .b16 m;\n\tmov.b16 m, 3;\n\telect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mma_pair(uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, pe;\n\t.reg .b64 da, db;\n\telect.sync _|pe, 0xffffffff;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t@pe tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi),
      "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_single(uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
  tc_mma_f16_lohi(d, a_lo, a_hi, b_lo, b_hi, idesc, acc);
}

struct Cfg { int n, m, chains, commits, rounds, pair, interleave; long long* out; int fence, probe; };
template <int PAIR>

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) bench(Cfg cc) {
  Cfg c = cc; c.pair = PAIR;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bars[8];
  const int warp = threadIdx.x / 32;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const uint32_t base = smem_u32(smem), bar = smem_u32(bars);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(bar + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    if (c.pair) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 0 && (rank == 0 || !c.pair)) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(c.n >> 3) << 17) | ((uint32_t)((c.pair ? 256 : 128) >> 4) << 24);
    const uint32_t rows = 178, nh = c.pair ? c.n / 2 : c.n;
    const uint32_t a_hi = (uint32_t)(make_desc(0, rows * 16, 128) >> 32), b_hi = (uint32_t)(make_desc(0, nh * 16, 128) >> 32);
    const uint32_t a0 = (uint32_t)make_desc(0, rows * 16, 128) + (base >> 4), b0 = (uint32_t)make_desc(0, nh * 16, 128) + ((base + 64 * 1024) >> 4);
    const long long t0 = clock64();
    uint32_t ph = 0;
    for (int r = 0; r < c.rounds; ++r) {
      if (!c.interleave) {
        for (int ch = 0; ch < c.chains; ++ch) {
          const uint32_t d = tmem + (uint32_t)(((r * c.chains + ch) & 3) * 128);
          uint32_t acc = 0, a = a0 + ch * 8, b = b0;
          if (c.probe) mbar_wait(bar + 8 * 6, 1, 3);        // a satisfied wait (fresh barrier, parity 1)
          if (c.fence) tc_fence_after();
#pragma unroll 1
          for (int i = 0; i < c.m; i += 4) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (PAIR) mma_pair(d, a + k * 2 * rows, a_hi, b + k * 2 * nh, b_hi, idesc, acc);
              else mma_single(d, a + k * 2 * rows, a_hi, b + k * 2 * nh, b_hi, idesc, acc);
              acc = 1;
            }
            a += 3; b += (i & 4) ? 8 * nh : 0u - 8 * nh;
          }
          if (ch < c.commits - 1) { if (PAIR) commit_pair(bar + 8 * (1 + ch)); else tc_commit(bar + 8 * (1 + ch)); }
        }
      } else {      // the chains' MMAs alternate: tap by tap
        uint32_t acc = 0, a = a0, b = b0;
#pragma unroll 1
        for (int i = 0; i < c.m; i += 4) {
          for (int ch = 0; ch < c.chains; ++ch) {
            const uint32_t d = tmem + (uint32_t)(((r * c.chains + ch) & 3) * 128);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (PAIR) mma_pair(d, a + ch * 8 + k * 2 * rows, a_hi, b + k * 2 * nh, b_hi, idesc, acc);
              else mma_single(d, a + ch * 8 + k * 2 * rows, a_hi, b + k * 2 * nh, b_hi, idesc, acc);
            }
          }
          acc = 1;
          a += 3; b += (i & 4) ? 8 * nh : 0u - 8 * nh;
        }
      }
      if (c.commits > 0) { if (PAIR) commit_pair(bar); else tc_commit(bar); }
    }
    if (PAIR) commit_pair(bar + 8 * 7); else tc_commit(bar + 8 * 7);     // tracks every MMA issued so far
    mbar_wait(bar + 8 * 7, 0, 1);
    (void)ph;
    const long long t1 = clock64();
    if (threadIdx.x == 0) c.out[blockIdx.x / 2] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 0) {
    tc_fence_after();
    if (c.pair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

int main() {
  long long* out;
  cudaMalloc(&out, 1024 * sizeof(long long));
  cudaFuncSetAttribute(bench<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int rounds = 200;
  printf("%5s %4s %4s %6s %7s %5s %10s %10s %8s\n", "pair", "N", "m", "chains", "commits", "intl", "clk/round", "clk/MMA", "ms");
  for (int pair = 0; pair < 2; ++pair)
    for (int n : {32, 64, 128})
     if (pair || n <= 64)
      for (int m : {12, 44})
        for (int variant = 0; variant < 8; ++variant) {
          // 0: one chain, no commits   1: two chains, no commits   2: two chains, 1 commit per round   3: two chains, 2 commits per round
          // 4: two chains interleaved, 1 commit per round
          // 5: as 2 + tcgen05.fence::after_thread_sync per chain   6: as 2 + a satisfied mbarrier wait per chain   7: both
          Cfg c{n, m, variant == 0 ? 1 : 2, variant <= 1 ? 0 : (variant == 3 ? 2 : 1), rounds, pair, variant == 4, out, variant == 5 || variant == 7, variant >= 6};
          cudaEvent_t e0, e1;
          cudaEventCreate(&e0); cudaEventCreate(&e1);
          if (pair) bench<1><<<148, 128, 200 * 1024>>>(c); else bench<0><<<148, 128, 200 * 1024>>>(c);
          cudaEventRecord(e0);
          if (pair) bench<1><<<148, 128, 200 * 1024>>>(c); else bench<0><<<148, 128, 200 * 1024>>>(c);
          cudaEventRecord(e1);
          if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
          float ms;
          cudaEventElapsedTime(&ms, e0, e1);
          long long h[74];
          cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
          double clk = 0;
          for (int i = 0; i < 74; ++i) clk += (double)h[i] / 74;
          printf("%5d %4d %4d %6d %7d %5d %10.0f %10.1f %8.3f%s%s\n", pair, n, m, c.chains, c.commits, c.interleave, clk / rounds,
                 clk / rounds / (c.m * c.chains), ms, c.fence ? "  +fence" : "", c.probe ? "  +wait" : "");
        }
  return 0;
}
