// PTX wrappers shared by the tcgen05 kernels (sm_100a): mbarrier, TMA bulk copy, tcgen05 mma/commit/ld, descriptors.
#pragma once
#include "common.cuh"
#include <cuda_fp16.h>

namespace vs {
namespace umma {

constexpr int kTileM = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Non-blocking probe (try_wait may suspend the thread for a while before it returns false).
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Blocking wait = spin on try_wait (which itself parks the warp for a short, hardware-chosen time).  A suspend-time hint
// (try_wait ..., ns -> NANOSLEEP.SYNCS) removes the spin instructions but wakes up later: measured slower on every conv.
// Bounded: a protocol bug must abort the kernel (trap) instead of hanging the GPU; the clock is only read every 256 probes
// so the loop is 3 instructions (spinning warps were 25 % of all issued instructions in the fused ResBlock kernel).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int tag) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  for (;;) {
#pragma unroll 1
    for (int it = 0; it < 256; ++it)
      if (mbar_try_wait(bar, parity)) return;
    if (clock64() - t0 > 4000000000LL) {
      printf("umma: mbarrier timeout tag=%d block=%d thread=%d parity=%u\n", tag, blockIdx.x, threadIdx.x, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {   // whole warp calls, one elected lane commits
  asm volatile(
      "{\n\t.reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar)
      : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers f16 inputs with fp32 accumulation.
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same, with the descriptors passed as 32-bit halves: only the low words (start address, 16 B units) change
// between MMAs, so the issue loop is two integer adds per instruction.
// Executed by the whole (converged) MMA warp: one elected lane issues, so the surrounding loop stays
// warp-uniform and the compiler keeps the descriptors in uniform registers.
__device__ __forceinline__ void tc_mma_f16_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                                 uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, pe;\n\t.reg .b64 da, db;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// All MMAs of one 128-row tile against resident weights: taps x NK K-chunks, accumulating into d_tmem.
// One warp issues every MMA of a CTA, so this loop's instruction count bounds the small-N convs (tools/mma_microbench.cu:
// runtime-nested loops cost 55-120 clk per MMA against a 40-48 clk operand-fetch floor): the K-chunk walk is unrolled.
template <int NK>
__device__ __forceinline__ void issue_tile_acc(uint32_t d_tmem, uint32_t a_tile, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                               uint32_t idesc, int taps, uint32_t dil, uint32_t a_kstep, uint32_t b_kstep,
                                               uint32_t accumulate) {   // accumulate = 1: D += (keeps what is in TMEM)
  uint32_t a_tap = a_tile;
#pragma unroll 2
  for (int t = 0; t < taps; ++t, a_tap += dil) {
    uint32_t a_lo = a_tap;
#pragma unroll
    for (int ks = 0; ks < NK; ++ks) {
      tc_mma_f16_lohi(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, accumulate);
      accumulate = 1;
      a_lo += a_kstep;
      b_lo += b_kstep;
    }
  }
}
template <int NK>
__device__ __forceinline__ void issue_tile(uint32_t d_tmem, uint32_t a_tile, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                           uint32_t idesc, int taps, uint32_t dil, uint32_t a_kstep, uint32_t b_kstep) {
  issue_tile_acc<NK>(d_tmem, a_tile, a_hi, b_lo, b_hi, idesc, taps, dil, a_kstep, b_kstep, 0u);
}

// Shared-memory matrix descriptor, K-major, SWIZZLE_NONE (cute::UMMA::SmemDescriptor bit layout):
// [0,14) start>>4, [16,30) LBO>>4 (pitch between the two 16-byte K chunks of one MMA),
// [32,46) SBO>>4 (pitch between 8-row groups), [46,48) version = 1 on sm_100, [61,64) layout = 0.
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 = 1 @4, a/b_format @7/@10 (F16 = 0, BF16 = 1),
// a/b major K = 0 @15/@16, N>>3 @17, M>>4 @24.  The decoder's operands are IEEE fp16 (11-bit significand, same tensor
// rate as bf16): with bf16 operands the log-mel of the waveform misses the 1e-2 bar by 10x (precision).
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// two fp32 -> packed fp16x2, round to nearest, saturating to +-65504 (an overflow must not turn into inf: the next
// conv would spread NaNs over its whole receptive field)
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void unpack_f16x8(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

}  // namespace umma
}  // namespace vs
