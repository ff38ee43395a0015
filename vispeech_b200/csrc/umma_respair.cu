// One ResBlock1 iteration (reference modules.py:211-220) as ONE tcgen05 kernel for the narrow decoder stages:
//
//     y = c2( lrelu( c1( lrelu(x) ) + b1 ) ) + b2 + x        (c1: k taps, dilation d;  c2: k taps, dilation 1)
//
// The unfused chain (umma_conv.cu) moves 5 activation-sized tensors through HBM per iteration (read a = lrelu(x), write
// lrelu(c1), read it back, read a again for the residual, write the result) and stages 2-3 (C = 64, 32) sit on the HBM
// roofline of that traffic.  Here ONE tensor is read and ONE written per iteration:
//   * only the ACTIVATED stream a = lrelu(x) travels between iterations.  It is the conv's A operand as is (TMA bulk copy
//     straight into the UMMA layout), and the residual is recovered in the last epilogue by inverting the leaky-relu,
//     x = min(a, a / slope): in f16 that costs the same relative rounding as storing x itself;
//   * the intermediate lrelu(c1 + b1) is written by the first epilogue directly into shared memory in the layout the
//     second conv reads (rows outside the sequence / in gaps are zeroed there: they are c2's zero padding).
//
// Per CTA, tile i covers 128 conv1 rows and L = 128 - (k-1) output rows.  Five roles run as a software pipeline over the
// CTA's tiles, each hand-off an mbarrier ring:
//   warp 0      producer   XA[i % SX]  <- rows [s0-h1, s0+128+h1) of a (zero fill outside [0,R))
//   warp 1      MMA        conv1(i) : XA -> acc1[i&1]   then   conv2(i-1) : A2[(i-1)&1] -> acc2[(i-1)&1]
//   EW warps    epilogue 1 acc1[i&1] -> +b1 -> lrelu -> mask -> f16 -> A2[i&1] (smem)
//   EW warps    epilogue 2 acc2[i&1] + b2 + lrelu^-1(XA rows) [+ MRF sum] -> lrelu -> HBM
// While the tensor pipe runs, its operand fetch owns shared memory: an LDS / STS / mbarrier probe from another warp takes
// ~250 clk and a tcgen05.ld ~300 (tools/mma_microbench.cu), so each epilogue is a chain of a few such round trips, about
// 1.5-2.5k clk per tile.  Hence no redundant barriers (acc1_full(i) already implies conv2(i-2) has released A2[i&1],
// acc2_full(i) that XA(i) has landed), biases in the kernel's constant bank, an MMA warp that probes all four of its
// barriers at once - and, because ncu shows the kernel bound by the instruction issue rate (IPC 2.8, 70 % of slots, two
// thirds of them epilogue arithmetic), epilogues with a branch-free path for fully valid warps and compile-time output
// modes.
// conv1 of the next tile is issued before conv2 of the current one, so the tensor pipe works on conv1(i+1) while epilogue 1
// turns acc1(i) into A2(i), and on conv2(i) while epilogue 2 drains tile i-1.  Both convs' weights stay resident in shared
// memory; the MMA issue loops are the unrolled ones of umma_common.cuh (issue rate bounds N = 32 / 64).
#include "umma_conv.cuh"
#include "umma_common.cuh"
#include <type_traits>

namespace vs {
namespace {

using namespace umma;

// warp 0 producer, warp 1 MMA, then EW epilogue-1 warps and EW epilogue-2 warps; every epilogue warp owns one TMEM lane
// quarter x one 32-column chunk of a tile: EW = N / 8 (4 warps per stage at N = 32, 8 at N = 64)
__host__ __device__ constexpr int threads_for(int n) { return 64 + 2 * (n / 8) * 32; }
// what epilogue 2 writes (compile-time, so its dead paths vanish: the kernel is bound by the SM's instruction issue rate)
enum { M_ACT = 0, M_RAW = 1, M_RAW_RES2 = 2, M_ACT_RES2_SCALE = 3, M_GENERIC = 4 };
constexpr int kMaxSX = 4;         // input ring depth (>= 3 keeps load(i+2), conv1(i+1) and the residual read of i apart)

struct Plan {
  int L, h1, h2, planes, rows_x, rows_a2, n_tiles, tmem_cols, ctas_per_sm, row_div_shift, SX;
  int tight;              // 1: two buffers shared by input rows and intermediate, residual re-read from L2 (C = 64, k = 11)
  int pm;                 // tap pairing (C = 64): bit 0 conv1, bit 1 conv2 run as N = 128 MMAs over PAIRS of taps (see the kernel)
  uint32_t xa_bytes, a2_bytes, w_bytes, smem_bytes;
  uint32_t off_a2, off_w1, off_w2, off_bar, off_xch;
};
struct Params {
  UmmaPair c;
  Plan p;
  float bias[2][64];      // b1, b2 in the kernel's constant bank: the epilogues' bias adds take them as immediate operands
                          // (shared memory is saturated by the MMA operand fetch, an LDS there costs ~200 clk)
  long long* dbg;         // wait-clock counters, see umma_conv.cu (only with -DVS_UMMA_TIMING)
};
#ifdef VS_UMMA_TIMING
#define VS_TIMED(var, stmt)                         \
  do {                                              \
    const long long _t0 = dbg ? clock64() : 0;      \
    stmt;                                           \
    if (dbg) var += clock64() - _t0;                \
  } while (0)
#else
#define VS_TIMED(var, stmt) stmt
#endif

// TIGHT (C = 64, k = 11): both weight sets take 180 KB, which leaves room for TWO activation buffers in all.  Each buffer
// serves one tile from start to end: it receives the input rows XA(i); once conv1(i) has consumed them, epilogue 1 writes
// the intermediate A2(i) over them (in A2's own layout); conv2(i) reads that and its commit hands the buffer back to the
// producer for tile i+2.  The residual is re-read from global memory (L2) in epilogue 2, since XA(i) is gone by then.
// The MMA warp issues tiles in pairs - conv1(a), conv1(b), conv2(a), conv2(b) - so that every hand-off (epilogue 1 of a
// under conv1(b), epilogue 1 of b under conv2(a), the load for a+2 under conv2(b), the load for b+2 under conv1(a+2)) has
// a whole conv of tensor time to hide behind.  Measured per pair at the C2 size: 1.23 ms (~1040 TFLOP/s) against 1.30-1.45
// for the two unfused kernels; a first form with ONE buffer of each kind ran at 1.27, a second MMA-issuing warp for conv2
// at 1.31 (no gain: at N = 64 the 48-clk operand fetch, not the issue loop, is what the tensor pipe waits for).
// out[j] += odd[j + d] for the paired form: vo holds this thread's row of the odd half; the value of row j + d comes from lane + d of
// the same warp, or - for the last d lanes - from the first d lanes of the next lane quarter's warp (same column chunk) through `xch`
// (f16 [quarter][chunk][d rows][32]); the last quarter's last d rows have no neighbour (those rows are outside the tile's valid range).
// Two named barriers per call: all `nthreads` threads of this epilogue stage pass them once per tile.
__device__ __forceinline__ void add_shifted_odd(uint32_t (&v)[32], const uint32_t (&vo)[32], int d, int q, int cc, int lane, uint32_t xch,
                                                int bar_id, int nthreads) {
  const uint32_t my = xch + (uint32_t)((q * 2 + cc) * d) * 64u, next = xch + (uint32_t)(((q + 1) * 2 + cc) * d) * 64u;
  if (lane < d) {
#pragma unroll
    for (int i = 0; i < 32; i += 8)
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(my + (uint32_t)lane * 64u + (uint32_t)i * 2u),
                   "r"(pack_f16x2(__uint_as_float(vo[i]), __uint_as_float(vo[i + 1]))), "r"(pack_f16x2(__uint_as_float(vo[i + 2]), __uint_as_float(vo[i + 3]))),
                   "r"(pack_f16x2(__uint_as_float(vo[i + 4]), __uint_as_float(vo[i + 5]))), "r"(pack_f16x2(__uint_as_float(vo[i + 6]), __uint_as_float(vo[i + 7])))
                   : "memory");
  }
  asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nthreads) : "memory");
  const bool from_next = lane >= 32 - d;
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    uint32_t ww[4] = {0u, 0u, 0u, 0u};
    if (from_next && q < 3)
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(ww[0]), "=r"(ww[1]), "=r"(ww[2]), "=r"(ww[3])
                   : "r"(next + (uint32_t)(lane - (32 - d)) * 64u + (uint32_t)i * 2u));
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&ww[h]));
      const float s0 = __shfl_down_sync(0xffffffffu, __uint_as_float(vo[i + 2 * h]), d);
      const float s1 = __shfl_down_sync(0xffffffffu, __uint_as_float(vo[i + 2 * h + 1]), d);
      v[i + 2 * h] = __float_as_uint(__uint_as_float(v[i + 2 * h]) + (from_next ? f.x : s0));
      v[i + 2 * h + 1] = __float_as_uint(__uint_as_float(v[i + 2 * h + 1]) + (from_next ? f.y : s1));
    }
  }
  asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nthreads) : "memory");    // nobody overwrites xch for the next tile before every read is done
}

// TAP PAIRING (PM != 0, C = 64): an SS-mode MMA re-reads its 4 KB A tile from shared memory for every K = 16 step, so at N = 64
// it costs 51 clk for 32 clk of tensor time (tools/pair_microbench.cu).  Two taps t, t + 1 of a conv read the SAME A rows if the
// second one's result is allowed to land one tap offset lower: D[i][0:64] += W_t a[i + t d] is tap t's share of out[i], and
// D[i][64:128] += W_{t+1} a[i + t d] is tap t + 1's share of out[i - d].  So the taps are issued in pairs as ONE N = 128 MMA
// (64 clk, tensor-bound) against [W_t | W_{t+1}], the A offset advances by 2 d per pair (the odd last tap runs alone at N = 64
// into the even half), and the epilogue forms out[i] = D[i][0:64] + D[i + d][64:128]: a warp shuffle down by d lanes, plus, for
// the last d rows of every 32-row lane quarter, an exchange with the next quarter's warp through a few hundred bytes of shared
// memory (f16: it is half of a sum that is rounded to f16 right afterwards anyway).  The last d rows of the tile have no
// neighbour: a tile yields d rows less (113 - 117 instead of 118 at k = 11).  MMA time per tile: -37 %.
// MEASURED (tools/pair_timing.py, C2 size): k = 11 1.244 -> 1.205 ms, k = 7 0.875 -> 1.01, k = 3 0.52 -> 0.83: the second TMEM load, the two
// named barriers and the exchange lengthen the chain conv1 -> epilogue 1 -> conv2 by more than the MMAs save (two tiles in flight per
// CTA cannot hide it; TMEM and shared memory have no room for a third).  Option "tap_pairs" (default 0) keeps the form testable.
template <int N, int MODE, bool TIGHT = false, int PM = 0>
__global__ void __launch_bounds__(threads_for(N), N == 32 ? 2 : 1) umma_respair_kernel(const __grid_constant__ Params prm) {
  static_assert(PM == 0 || N == 64, "tap pairing is for C = 64");
  constexpr int EW = N / 8, kThreads = threads_for(N);
  constexpr uint32_t ACCW = PM ? 128u : (uint32_t)N;          // accumulator slot width; acc1 ring [0, 2 ACCW), acc2 ring [2 ACCW, 4 ACCW)
  constexpr bool P1 = (PM & 1) != 0, P2 = (PM & 2) != 0;
  extern __shared__ __align__(128) uint8_t smem[];
  pdl_trigger();
  const UmmaPair& c = prm.c;
  const Plan& p = prm.p;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
#ifdef VS_UMMA_TIMING
  long long* const dbg = prm.dbg;
  long long tw0 = 0, tw1 = 0, tw2 = 0;
  const long long t_start = dbg ? clock64() : 0;
#endif

  const uint32_t smem_base = smem_u32(smem);
  const uint32_t xa = smem_base, a2 = smem_base + p.off_a2;
  const uint32_t w1 = smem_base + p.off_w1, w2 = smem_base + p.off_w2, bar = smem_base + p.off_bar;
  // barrier table (8 B each)
  const uint32_t w_full = bar;
  auto xa_full = [&](uint32_t i) { return bar + 8u * (1 + i); };
  auto xa_empty = [&](uint32_t i) { return bar + 8u * (1 + kMaxSX + i); };
  auto acc1_full = [&](uint32_t i) { return bar + 8u * (1 + 2 * kMaxSX + i); };
  auto acc1_empty = [&](uint32_t i) { return bar + 8u * (3 + 2 * kMaxSX + i); };
  auto a2_full = [&](uint32_t i) { return bar + 8u * (5 + 2 * kMaxSX + i); };
  auto acc2_full = [&](uint32_t i) { return bar + 8u * (9 + 2 * kMaxSX + i); };
  auto acc2_empty = [&](uint32_t i) { return bar + 8u * (11 + 2 * kMaxSX + i); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + p.off_bar + 8 * (13 + 2 * kMaxSX));

  if (threadIdx.x == 0) {
    mbar_init(w_full, 1);
    for (int i = 0; i < p.SX; ++i) { mbar_init(xa_full(i), 1); mbar_init(xa_empty(i), TIGHT ? 1 : 1 + EW); }   // conv1 commit (+ epilogue 2 warps)
    for (int i = 0; i < 2; ++i) {
      mbar_init(acc1_full(i), 1); mbar_init(acc1_empty(i), EW);
      mbar_init(a2_full(i), EW);
      mbar_init(acc2_full(i), 1); mbar_init(acc2_empty(i), EW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // rows [128, 128 + 2*h2) of both A2 buffers are read by the last taps of conv2 (those outputs are discarded): keep
  // them finite
  for (int i = threadIdx.x; i < (TIGHT ? 0 : 2) * p.planes * 2 * p.h2; i += kThreads) {
    const int bsel = i / (p.planes * 2 * p.h2), r = i % (p.planes * 2 * p.h2);
    const int pl = r / (2 * p.h2), j = r % (2 * p.h2);
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(a2 + bsel * p.a2_bytes + (uint32_t)(pl * p.rows_a2 + 128 + j) * 16u),
                 "r"(0)
                 : "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();                              // the prologue above overlapped the previous kernel's tail
  const uint32_t tmem_base = *tmem_slot;     // columns [0, 2N): acc1 ring, [2N, 4N): acc2 ring
  const uint32_t SX = (uint32_t)p.SX;

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    if (lane == 0) mbar_arrive_expect_tx(w_full, 2 * p.w_bytes);
    __syncwarp();
    if (PM == 0) {
      if (lane == 0) {
        bulk_g2s(w1, c.w1, p.w_bytes, w_full);
        bulk_g2s(w2, c.w2, p.w_bytes, w_full);
      }
    } else {
      // a paired conv keeps its taps as [pair g][plane][128 columns = tap 2g | tap 2g + 1][8], then the odd last tap as [plane][64][8];
      // the packed source is [tap][plane][64][8]: one 1 KB copy per (tap, plane)
      const int per_conv = c.taps * p.planes, n_pairs2 = c.taps - 1;           // taps below n_pairs2 are paired
      for (int i = lane; i < 2 * per_conv; i += 32) {
        const int cv = i / per_conv, t = (i % per_conv) / p.planes, pl = i % p.planes;
        const bool paired = cv == 0 ? P1 : P2;
        const __half* src = (cv ? c.w2 : c.w1) + (size_t)(t * p.planes + pl) * N * 8;
        uint32_t dst = cv ? w2 : w1;
        if (!paired) dst += (uint32_t)(t * p.planes + pl) * (N * 16u);
        else if (t < n_pairs2) dst += (uint32_t)(((t >> 1) * p.planes + pl) * 2 * N + (t & 1) * N) * 16u;
        else dst += (uint32_t)(n_pairs2 / 2) * (uint32_t)p.planes * 2u * N * 16u + (uint32_t)pl * (N * 16u);
        bulk_g2s(dst, src, N * 16u, w_full);
      }
    }
    uint32_t slot = 0, phase = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      VS_TIMED(tw0, mbar_wait(xa_empty(slot), phase ^ 1, 21));
      const uint32_t stage = xa + slot * p.xa_bytes;
      const int s0 = tile * p.L - p.h2;
      const int row_lo = s0 - p.h1, row_hi = row_lo + p.rows_x;
      const int c_lo = row_lo < 0 ? 0 : row_lo, c_hi = row_hi > c.R ? c.R : row_hi;
      const int n_zero_lo = c_lo - row_lo, n_zero_hi = row_hi - c_hi;
      if (n_zero_lo > 0 || n_zero_hi > 0) {
        const int per_plane = n_zero_lo + n_zero_hi;
        for (int i = lane; i < p.planes * per_plane; i += 32) {
          const int pl = i / per_plane, j = i % per_plane;
          const int row = j < n_zero_lo ? j : (p.rows_x - n_zero_hi + (j - n_zero_lo));
          asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(stage + (uint32_t)(pl * p.rows_x + row) * 16u), "r"(0)
                       : "memory");
        }
        fence_proxy_async();
      }
      __syncwarp();
      const uint32_t bytes = (uint32_t)(c_hi - c_lo) * 16u;
      if (lane == 0) mbar_arrive_expect_tx(xa_full(slot), bytes * p.planes);
      __syncwarp();
      for (int pl = lane; pl < p.planes; pl += 32)
        bulk_g2s(stage + (uint32_t)(pl * p.rows_x + n_zero_lo) * 16u, c.x + ((size_t)pl * c.R + c_lo) * 8, bytes, xa_full(slot));
      if (++slot == SX) { slot = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (warp-uniform, elected lane issues)
    const uint32_t idesc = make_idesc(N);
    constexpr uint32_t b_lbo = (uint32_t)N * 16u, b_kstep = 2u * N;
    constexpr int NK = N / 16;
    const uint32_t b_hi = (uint32_t)(make_desc(0, b_lbo, 128u) >> 32), b_lo_fixed = (uint32_t)make_desc(0, b_lbo, 128u);
    const uint32_t a1_lbo = (uint32_t)p.rows_x * 16u, a2_lbo = (uint32_t)p.rows_a2 * 16u;
    const uint32_t a1_hi = (uint32_t)(make_desc(0, a1_lbo, 128u) >> 32), a1_lo_fixed = (uint32_t)make_desc(0, a1_lbo, 128u);
    const uint32_t a2_hi = (uint32_t)(make_desc(0, a2_lbo, 128u) >> 32), a2_lo_fixed = (uint32_t)make_desc(0, a2_lbo, 128u);
    const uint32_t a1_kstep = 2u * (uint32_t)p.rows_x, a2_kstep = 2u * (uint32_t)p.rows_a2;
    const uint32_t w1_lo = b_lo_fixed + (w1 >> 4), w2_lo = b_lo_fixed + (w2 >> 4);
    const int taps = c.taps;
    const uint32_t dil = (uint32_t)c.dil;
    // paired issue (see the kernel's head comment): (taps - 1) / 2 MMA groups of N = 128 over tap pairs, then the odd last tap at N = 64
    const uint32_t idesc2 = make_idesc(2 * N);
    const uint32_t b2_hi = (uint32_t)(make_desc(0, 2u * N * 16u, 128u) >> 32), b2_lo_fixed = (uint32_t)make_desc(0, 2u * N * 16u, 128u);
    auto issue_pairs = [&](uint32_t d_tmem, uint32_t a_tile, uint32_t a_hi_, uint32_t w_addr, uint32_t step, uint32_t a_kstep_) {
      const int np = (taps - 1) / 2;
      uint32_t a_tap = a_tile, b = b2_lo_fixed + (w_addr >> 4), accumulate = 0;
#pragma unroll 1
      for (int g = 0; g < np; ++g, a_tap += 2u * step) {
#pragma unroll
        for (int k = 0; k < NK; ++k) {
          tc_mma_f16_lohi(d_tmem, a_tap + (uint32_t)k * a_kstep_, a_hi_, b + (uint32_t)k * (4u * N), b2_hi, idesc2, accumulate);
          accumulate = 1;
        }
        b += (uint32_t)p.planes * 2u * N;
      }
      uint32_t bs = b_lo_fixed + ((w_addr + (uint32_t)np * (uint32_t)p.planes * 2u * N * 16u) >> 4);
#pragma unroll
      for (int k = 0; k < NK; ++k) tc_mma_f16_lohi(d_tmem, a_tap + (uint32_t)k * a_kstep_, a_hi_, bs + (uint32_t)k * b_kstep, b_hi, idesc, 1u);
    };
    mbar_wait(w_full, 0, 22);
    tc_fence_after();
    auto conv2 = [&](uint32_t j, bool ready) {   // conv2 of this CTA's j-th tile: A2 rows o + t
      const uint32_t b = j & 1u, ph = (j >> 1) & 1u;
      if (!ready) {
        VS_TIMED(tw1, mbar_wait(a2_full(b), ph, 24));
        VS_TIMED(tw2, mbar_wait(acc2_empty(b), ph ^ 1u, 25));
      }
      tc_fence_after();
      if (P2)
        issue_pairs(tmem_base + (2u + b) * ACCW, a2_lo_fixed + ((TIGHT ? xa + b * p.xa_bytes : a2 + b * p.a2_bytes) >> 4), a2_hi, w2, 1u, a2_kstep);
      else
        issue_tile<NK>(tmem_base + (2u + b) * ACCW, a2_lo_fixed + ((TIGHT ? xa + b * p.xa_bytes : a2 + b * p.a2_bytes) >> 4), a2_hi, w2_lo,
                       b_hi, idesc, taps, 1u, a2_kstep, b_kstep);
      tc_commit(acc2_full(b));
      if (TIGHT) tc_commit(xa_empty(b));          // the buffer goes back to the producer (tile j + 2)
    };
    uint32_t slot = 0, phase = 0, i = 0;
    if (TIGHT) {
      auto conv1 = [&](uint32_t j) {               // buffer j & 1, accumulator j & 1
        const uint32_t b = j & 1u, ph = (j >> 1) & 1u;
        VS_TIMED(tw0, mbar_wait(xa_full(b), ph, 23));
        VS_TIMED(tw0, mbar_wait(acc1_empty(b), ph ^ 1u, 26));
        tc_fence_after();
        if (P1) issue_pairs(tmem_base + b * ACCW, a1_lo_fixed + ((xa + b * p.xa_bytes) >> 4), a1_hi, w1, dil, a1_kstep);
        else issue_tile<NK>(tmem_base + b * ACCW, a1_lo_fixed + ((xa + b * p.xa_bytes) >> 4), a1_hi, w1_lo, b_hi, idesc, taps, dil,
                            a1_kstep, b_kstep);
        tc_commit(acc1_full(b));
      };
      uint32_t n_mine = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) ++n_mine;
      for (uint32_t j = 0; j < n_mine; j += 2) {
        const bool two = j + 1 < n_mine;
        conv1(j);
        if (two) conv1(j + 1);
        conv2(j, false);
        if (two) conv2(j + 1, false);
      }
    } else {
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++i) {
      const uint32_t b = i & 1u, ph = (i >> 1) & 1u;
      // probe all four barriers of this iteration at once (each probe is a ~250 clk shared-memory round trip)
      const uint32_t jb = (i - 1u) & 1u, jph = ((i - 1u) >> 1) & 1u;
      const bool r0 = mbar_test_wait(xa_full(slot), phase), r1 = mbar_test_wait(acc1_empty(b), ph ^ 1u);
      bool r2 = false, r3 = false;
      if (i > 0) { r2 = mbar_test_wait(a2_full(jb), jph); r3 = mbar_test_wait(acc2_empty(jb), jph ^ 1u); }
      if (!r0) VS_TIMED(tw0, mbar_wait(xa_full(slot), phase, 23));
      if (!r1) VS_TIMED(tw0, mbar_wait(acc1_empty(b), ph ^ 1u, 26));
      tc_fence_after();
      if (P1) issue_pairs(tmem_base + b * ACCW, a1_lo_fixed + ((xa + slot * p.xa_bytes) >> 4), a1_hi, w1, dil, a1_kstep);
      else issue_tile<NK>(tmem_base + b * ACCW, a1_lo_fixed + ((xa + slot * p.xa_bytes) >> 4), a1_hi, w1_lo, b_hi, idesc, taps, dil,
                          a1_kstep, b_kstep);     // conv1: XA rows j + t*d
      tc_commit(acc1_full(b));
      tc_commit(xa_empty(slot));
      if (++slot == SX) { slot = 0; phase ^= 1; }
      if (i > 0) conv2(i - 1, r2 && r3);
    }
    if (i > 0) conv2(i - 1, false);
    }
  } else if (warp < 2 + EW) {
    // ------------------------------------------------------------------ epilogue 1: acc1 -> A2 = lrelu(c1 + b1)
    const int q = warp & 3, cc = (warp - 2) >> 2;       // TMEM lane quarter, 32-column chunk
    const float slope = c.in_slope;
    const int j = q * 32 + lane;                        // A2 local row = conv1 output row
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cc * 32);
    const uint32_t xch1 = smem_base + p.off_xch;
    const uint32_t a2_lane = (TIGHT ? xa : a2) + (uint32_t)j * 16u + (uint32_t)(cc * 4 * p.rows_a2) * 16u;
    const uint32_t a2_plane = (uint32_t)p.rows_a2 * 16u;
    uint32_t i = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++i) {
      const uint32_t b = i & 1u, ph = (i >> 1) & 1u;
      const int g = tile * p.L - p.h2 + j;
      bool valid = g >= 0 && g < c.R;                   // outside the sequence / in gap rows: c2 must see zero padding
      if (valid && c.row_utt) valid = c.row_utt[g >> p.row_div_shift] >= 0;
      const uint32_t keep = valid ? 0xFFFFFFFFu : 0u;
      VS_TIMED(tw0, mbar_wait(acc1_full(b), ph, 27));
      tc_fence_after();
      uint32_t v[32];
      tmem_ld32(t_lane + b * ACCW, v);
      if (P1) {                                          // out1[j] = even[j] + odd[j + d]
        uint32_t vo[32];
        tmem_ld32(t_lane + b * ACCW + (uint32_t)N, vo);
        add_shifted_odd(v, vo, c.dil, q, cc, lane, xch1, 1, EW * 32);
      }
      const uint32_t a2_row = a2_lane + b * (TIGHT ? p.xa_bytes : p.a2_bytes);   // TIGHT: over the tile's own input rows
      auto chunk = [&](auto cc_tag) {
        constexpr int CC = decltype(cc_tag)::value;
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) {
          float y[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float t = __uint_as_float(v[8 * gq + e]) + prm.bias[0][CC * 32 + gq * 8 + e];
            y[e] = fmaxf(t, t * slope);
          }
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a2_row + (uint32_t)gq * a2_plane),
                       "r"(pack_f16x2(y[0], y[1]) & keep), "r"(pack_f16x2(y[2], y[3]) & keep),
                       "r"(pack_f16x2(y[4], y[5]) & keep), "r"(pack_f16x2(y[6], y[7]) & keep)
                       : "memory");
        }
      };
      if (N == 32 || cc == 0) chunk(std::integral_constant<int, 0>{});
      else chunk(std::integral_constant<int, 1>{});
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(a2_full(b)); mbar_arrive(acc1_empty(b)); }
    }
  } else {
    // ------------------------------------------------------------------ epilogue 2: acc2 + b2 + x [+ MRF sum] -> HBM
    constexpr bool kGen = MODE == M_GENERIC;
    const bool has_res2 = kGen ? (c.res2 != nullptr) : (MODE == M_RAW_RES2 || MODE == M_ACT_RES2_SCALE);
    const bool has_raw = kGen ? (c.out_raw != nullptr) : (MODE == M_RAW || MODE == M_RAW_RES2);
    const bool has_act = kGen ? (c.out_act != nullptr) : (MODE == M_ACT || MODE == M_ACT_RES2_SCALE);
    const bool has_scale = kGen ? (c.act_scale != 1.f) : (MODE == M_ACT_RES2_SCALE);
    const int q = warp & 3, cc = (warp - 2 - EW) >> 2;
    const float oslope = c.act_slope, oscale = c.act_scale;
    const __half2 inv2 = __float2half2_rn(1.f / c.in_slope);
    const int o = q * 32 + lane;                        // conv2 output position within the tile
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + 2u * ACCW + (uint32_t)(cc * 32);
    const uint32_t xch2 = smem_base + p.off_xch + (P1 ? 8u * (uint32_t)c.dil * 64u : 0u);
    const uint32_t xa_lane = xa + (uint32_t)(p.h1 + p.h2 + o) * 16u + (uint32_t)(cc * 4 * p.rows_x) * 16u;
    const uint32_t xa_plane = (uint32_t)p.rows_x * 16u;
    const size_t plane_elems = (size_t)c.R * 8;         // elements between consecutive 8-channel planes in HBM
    uint32_t slot = 0, i = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++i) {
      const uint32_t b = i & 1u, ph = (i >> 1) & 1u;
      const int g = tile * p.L + o;
      const bool in_tile = o < p.L && g < c.R;
      int utt = -1;
      if (in_tile) utt = c.row_utt ? c.row_utt[g >> p.row_div_shift] : 0;
      const uint32_t keep = utt >= 0 ? 0xFFFFFFFFu : 0u;               // gap rows are written as exact zeros
      const size_t row_off = (size_t)(cc * 4) * plane_elems + (size_t)g * 8;
      uint4 rv2[4];
      if (has_res2 && in_tile) {                        // the MRF sum is fetched before waiting for the accumulator
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) rv2[gq] = *reinterpret_cast<const uint4*>(c.res2 + row_off + (size_t)gq * plane_elems);
      }
      uint4 xv[4];
      if (TIGHT) {                                      // the input stage is long gone: fetch the residual rows again (L2)
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) {
          xv[gq] = make_uint4(0u, 0u, 0u, 0u);
          if (in_tile) xv[gq] = *reinterpret_cast<const uint4*>(c.x + row_off + (size_t)gq * plane_elems);
        }
      }
      VS_TIMED(tw0, mbar_wait(acc2_full(b), ph, 29));
      tc_fence_after();
      if (!TIGHT) {                                     // the residual rows: XA(i) has landed once conv2(i) has completed
#pragma unroll
        for (int gq = 0; gq < 4; ++gq)
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(xv[gq].x), "=r"(xv[gq].y), "=r"(xv[gq].z), "=r"(xv[gq].w)
                       : "r"(xa_lane + slot * p.xa_bytes + (uint32_t)gq * xa_plane));
      }
      uint32_t v[32];
      tmem_ld32(t_lane + b * ACCW, v);
      if (P2) {                                          // out2[o] = even[o] + odd[o + 1]
        uint32_t vo[32];
        tmem_ld32(t_lane + b * ACCW + (uint32_t)N, vo);
        add_shifted_odd(v, vo, 1, q, cc, lane, xch2, 2, EW * 32);
      }
      auto chunk = [&](auto cc_tag) {
        constexpr int CC = decltype(cc_tag)::value;
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) {
          // x = lrelu^-1(a) = min(a, a / slope) on the packed pairs (1/slope = 10 is exact in f16; the product rounds like
          // a stored f16 x would have), then everything else in fp32
          const uint32_t aw[4] = {xv[gq].x, xv[gq].y, xv[gq].z, xv[gq].w};
          float y[8];
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const __half2 a2v = *reinterpret_cast<const __half2*>(&aw[h]);
            const __half2 x2 = __hmin2(a2v, __hmul2(a2v, inv2));
            const float2 xf = __half22float2(x2);
            y[2 * h] = __uint_as_float(v[8 * gq + 2 * h]) + prm.bias[1][CC * 32 + gq * 8 + 2 * h] + xf.x;
            y[2 * h + 1] = __uint_as_float(v[8 * gq + 2 * h + 1]) + prm.bias[1][CC * 32 + gq * 8 + 2 * h + 1] + xf.y;
          }
          if (has_res2) {
            float f[8];
            unpack_f16x8(rv2[gq], f);
#pragma unroll
            for (int e = 0; e < 8; ++e) y[e] += f[e];
          }
          const size_t go = row_off + (size_t)gq * plane_elems;
          if (has_raw) {
            const uint4 raw = make_uint4(pack_f16x2(y[0], y[1]) & keep, pack_f16x2(y[2], y[3]) & keep,
                                         pack_f16x2(y[4], y[5]) & keep, pack_f16x2(y[6], y[7]) & keep);
            if (in_tile) *reinterpret_cast<uint4*>(c.out_raw + go) = raw;
          }
          if (has_act) {
            float z[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float t = has_scale ? y[e] * oscale : y[e];
              z[e] = fmaxf(t, t * oslope);
            }
            const uint4 act = make_uint4(pack_f16x2(z[0], z[1]) & keep, pack_f16x2(z[2], z[3]) & keep,
                                         pack_f16x2(z[4], z[5]) & keep, pack_f16x2(z[6], z[7]) & keep);
            if (in_tile) *reinterpret_cast<uint4*>(c.out_act + go) = act;
          }
        }
      };
      if (N == 32 || cc == 0) chunk(std::integral_constant<int, 0>{});
      else chunk(std::integral_constant<int, 1>{});
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(acc2_empty(b)); if (!TIGHT) mbar_arrive(xa_empty(slot)); }
      if (++slot == SX) slot = 0;
    }
  }

#ifdef VS_UMMA_TIMING
  if (dbg && lane == 0 && (warp < 3 || warp == 2 + EW)) {   // [cta][producer | MMA | epilogue 1 | epilogue 2][total, waits]
    long long* o = dbg + ((size_t)blockIdx.x * 4 + (warp == 2 + EW ? 3 : warp)) * 4;
    o[0] = clock64() - t_start; o[1] = tw0; o[2] = tw1; o[3] = tw2;
  }
#endif
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

int make_plan(const UmmaPair& c, Plan* out) {
  Plan p{};
  VS_REQUIRE(c.C == 32 || c.C == 64, "umma_respair: C=%d (only 32 and 64 are fused)", c.C);
  VS_REQUIRE(c.taps % 2 == 1 && c.taps >= 3 && c.dil >= 1 && c.R > 0, "umma_respair: bad shape");
  VS_REQUIRE(c.in_slope > 0.f && c.in_slope < 1.f && c.act_slope > 0.f && c.act_slope <= 1.f, "umma_respair: bad slopes");
  p.h1 = c.dil * (c.taps - 1) / 2;
  p.h2 = (c.taps - 1) / 2;
  p.planes = c.C / 8;
  p.w_bytes = (uint32_t)c.taps * c.C * c.C * 2u;
  int s = 0;
  while ((1 << s) < c.row_div) ++s;
  VS_REQUIRE((1 << s) == c.row_div, "umma_respair: row_div=%d must be a power of two", c.row_div);
  p.row_div_shift = s;
  p.L = kTileM - 2 * p.h2;
  p.pm = 0;
  p.rows_x = kTileM + 2 * p.h1;
  p.rows_a2 = kTileM + 2 * p.h2;
  p.xa_bytes = (uint32_t)p.planes * p.rows_x * 16u;
  p.a2_bytes = (uint32_t)p.planes * p.rows_a2 * 16u;
  const uint32_t bar_bytes = 8u * (13 + 2 * kMaxSX) + 16u;
  const uint32_t fixed = 2 * p.a2_bytes + 2 * p.w_bytes + bar_bytes + 256u;
  // input ring: 3-4 stages; two CTAs per SM when that fits in half an SM's shared memory
  const uint32_t half_sm = 112u * 1024, full_sm = 224u * 1024;
  p.SX = 0;
  int per_sm = 1;
  if (fixed + 3 * p.xa_bytes <= half_sm) { per_sm = 2; p.SX = (fixed + 4 * p.xa_bytes <= half_sm) ? 4 : 3; }
  else if (fixed + 3 * p.xa_bytes <= full_sm) { p.SX = (fixed + 4 * p.xa_bytes <= full_sm) ? 4 : 3; }
  else if (fixed + 2 * p.xa_bytes <= full_sm) p.SX = 2;
  else if (c.C == 64 && fixed - 2 * p.a2_bytes + 2 * p.xa_bytes <= 227u * 1024 - 512u) { p.SX = 2; p.tight = 1; }
  VS_REQUIRE(p.SX > 0, "umma_respair: C=%d k=%d d=%d does not fit in shared memory", c.C, c.taps, c.dil);
  p.off_a2 = p.SX * p.xa_bytes;
  p.off_w1 = p.off_a2 + (p.tight ? 0 : 2) * p.a2_bytes;
  p.off_w2 = p.off_w1 + p.w_bytes;
  p.off_bar = (p.off_w2 + p.w_bytes + 127u) & ~127u;
  p.smem_bytes = p.off_bar + bar_bytes;
  p.tmem_cols = 4 * c.C;                               // 128 or 256: a power of two >= 32
  // tap pairing (C = 64, option "tap_pairs"): conv2 always when its 512-byte exchange buffer fits, conv1 too when its d x 512 bytes do
  p.off_xch = (p.smem_bytes + 15u) & ~15u;
  if (c.C == 64 && c.dil <= 5 && (opts().v[OPT_TAP_PAIRS] == 2 || (opts().v[OPT_TAP_PAIRS] == 1 && c.taps >= 11))) {   // 1: only where it measured faster
    const uint32_t cap = 227u * 1024;
    const uint32_t x1 = 8u * (uint32_t)c.dil * 64u, x2 = 8u * 64u;
    if (p.off_xch + x1 + x2 <= cap) p.pm = 3;
    else if (p.off_xch + x2 <= cap) p.pm = 2;
    if (p.pm) {
      p.smem_bytes = p.off_xch + ((p.pm & 1) ? x1 : 0u) + x2;
      p.tmem_cols = 512;
      if (p.pm & 1) p.L -= c.dil;                      // the last d conv1 rows of a tile have no neighbour to take their odd half from
      per_sm = 1;
    }
  }
  // never more CTAs on an SM than planned (TMEM: 512 columns)
  const uint32_t min_smem = (227u * 1024) / (uint32_t)(per_sm + 1) + 1024u;
  if (p.smem_bytes < min_smem) p.smem_bytes = min_smem;
  p.ctas_per_sm = per_sm;
  p.n_tiles = (c.R + p.L - 1) / p.L;
  *out = p;
  return VS_OK;
}

}  // namespace

// options: "respair_grid_div" (experiment knob: launch 1/div of the CTA slots, for co-scheduling tests) and "fused_respair"
// (0 off, 1 the C = 32 stage only, 2 (default) every ResBlock iteration whose two weight sets fit in smem)

bool umma_respair_supported(int C, int taps, int dil) {
  const int g_mode = (int)opts().v[OPT_FUSED_RESPAIR];
  if (g_mode == 0) return false;
  if (C == 128) return g_mode == 2 && opts().v[OPT_PAIR_FUSED] != 0 && umma_pairfused_supported(C, taps, dil);   // CTA-pair form only
  if (!(C == 32 || (C == 64 && g_mode == 2))) return false;
  UmmaPair c;
  c.C = C; c.taps = taps; c.dil = dil; c.R = 1024; c.row_div = 1; c.in_slope = 0.1f;
  Plan p;
  return make_plan(c, &p) == VS_OK;
}

int umma_respair(const UmmaPair& c, cudaStream_t st) {
  {   // option "pair_fused": 1 = C = 128 (k = 3) on the CTA-pair kernel, 2 = also C = 64 (umma_pairfused.cu)
    const int pf = (int)opts().v[OPT_PAIR_FUSED];
    if ((c.C == 128 || (c.C == 64 && pf >= 2)) && pf && umma_pairfused_supported(c.C, c.taps, c.dil)) return umma_pairfused(c, st);
  }
  Params prm;
  prm.c = c;
  prm.dbg = static_cast<long long*>(umma_conv_timing_buffer());
  VS_REQUIRE(c.x && c.w1 && c.w2 && c.b1 && c.b2 && (c.out_raw || c.out_act), "umma_respair: null pointer");
  VS_TRY(make_plan(c, &prm.p));
  if (c.b1_host && c.b2_host) {
    for (int i = 0; i < c.C; ++i) { prm.bias[0][i] = c.b1_host[i]; prm.bias[1][i] = c.b2_host[i]; }
  } else {   // op-level API (tests, tools): fetch the biases; the decoder passes host copies made at model finalize
    VS_CUDA_CHECK(cudaMemcpyAsync(prm.bias[0], c.b1, c.C * sizeof(float), cudaMemcpyDeviceToHost, st));
    VS_CUDA_CHECK(cudaMemcpyAsync(prm.bias[1], c.b2, c.C * sizeof(float), cudaMemcpyDeviceToHost, st));
    VS_CUDA_CHECK(cudaStreamSynchronize(st));
  }
  int n_sm = 0;
  VS_TRY(device_sm_count(&n_sm));
  int grid = n_sm * prm.p.ctas_per_sm / (int)opts().v[OPT_RESPAIR_GRID_DIV];
  if (grid < n_sm) grid = n_sm;
  if (grid > prm.p.n_tiles) grid = prm.p.n_tiles;
  int mode = M_GENERIC;
  const bool scale = c.act_scale != 1.f;
  if (c.out_act && !c.out_raw && !c.res2 && !scale) mode = M_ACT;
  else if (c.out_raw && !c.out_act && !c.res2) mode = M_RAW;
  else if (c.out_raw && !c.out_act && c.res2) mode = M_RAW_RES2;
  else if (c.out_act && !c.out_raw && c.res2 && scale) mode = M_ACT_RES2_SCALE;
#define VS_PAIR_CASE(NN, MM)                                                                                          \
  if (c.C == NN && mode == MM && !(NN == 64 && (prm.p.tight || prm.p.pm))) {                                          \
    VS_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(umma_respair_kernel<NN, MM>), 227 * 1024));              \
    VS_CUDA_CHECK(launch_pdl<4>(umma_respair_kernel<NN, MM>, dim3(grid), dim3(threads_for(NN)), prm.p.smem_bytes, st, prm));                                \
  }
#define VS_PAIR_TIGHT(MM)                                                                                             \
  if (c.C == 64 && mode == MM && prm.p.tight && !prm.p.pm) {                                                          \
    VS_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(umma_respair_kernel<64, MM, true>), 227 * 1024));        \
    VS_CUDA_CHECK(launch_pdl<4>(umma_respair_kernel<64, MM, true>, dim3(grid), dim3(threads_for(64)), prm.p.smem_bytes, st, prm));                          \
  }
#define VS_PAIR_PM(MM, TT, PP)                                                                                        \
  if (c.C == 64 && mode == MM && (prm.p.tight != 0) == TT && prm.p.pm == PP) {                                        \
    VS_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(umma_respair_kernel<64, MM, TT, PP>), 227 * 1024));      \
    VS_CUDA_CHECK(launch_pdl<4>(umma_respair_kernel<64, MM, TT, PP>, dim3(grid), dim3(threads_for(64)), prm.p.smem_bytes, st, prm)); \
  }
  if (prm.p.pm) {
    // tap-paired forms: the decoder's modes per branch (k = 7: act, raw + sum; k = 11: act, act + sum + scale; k = 3: act, raw) + generic
    const bool decoder_mode = mode == M_ACT || mode == M_RAW || mode == M_RAW_RES2 || mode == M_ACT_RES2_SCALE;
    if (!decoder_mode) mode = M_GENERIC;
    VS_PAIR_PM(M_ACT, false, 3) else VS_PAIR_PM(M_RAW, false, 3) else VS_PAIR_PM(M_RAW_RES2, false, 3) else VS_PAIR_PM(M_ACT_RES2_SCALE, false, 3)
    else VS_PAIR_PM(M_GENERIC, false, 3)
    else VS_PAIR_PM(M_ACT, true, 3) else VS_PAIR_PM(M_RAW, true, 3) else VS_PAIR_PM(M_RAW_RES2, true, 3) else VS_PAIR_PM(M_ACT_RES2_SCALE, true, 3)
    else VS_PAIR_PM(M_GENERIC, true, 3)
    else VS_PAIR_PM(M_ACT, true, 2) else VS_PAIR_PM(M_RAW, true, 2) else VS_PAIR_PM(M_RAW_RES2, true, 2) else VS_PAIR_PM(M_ACT_RES2_SCALE, true, 2)
    else VS_PAIR_PM(M_GENERIC, true, 2)
    else VS_PAIR_PM(M_ACT, false, 2) else VS_PAIR_PM(M_GENERIC, false, 2)
    else { set_error("umma_respair: no tap-paired instantiation for mode %d tight %d pm %d", mode, prm.p.tight, prm.p.pm); return VS_ERR_INVALID; }
  } else
  VS_PAIR_CASE(32, M_ACT) else VS_PAIR_CASE(32, M_RAW) else VS_PAIR_CASE(32, M_RAW_RES2) else VS_PAIR_CASE(32, M_ACT_RES2_SCALE)
  else VS_PAIR_CASE(32, M_GENERIC) else VS_PAIR_CASE(64, M_ACT) else VS_PAIR_CASE(64, M_RAW) else VS_PAIR_CASE(64, M_RAW_RES2)
  else VS_PAIR_CASE(64, M_ACT_RES2_SCALE) else VS_PAIR_CASE(64, M_GENERIC)
  else VS_PAIR_TIGHT(M_ACT) else VS_PAIR_TIGHT(M_RAW) else VS_PAIR_TIGHT(M_RAW_RES2) else VS_PAIR_TIGHT(M_ACT_RES2_SCALE)
  else VS_PAIR_TIGHT(M_GENERIC)
#undef VS_PAIR_PM
#undef VS_PAIR_TIGHT
#undef VS_PAIR_CASE
  VS_LAUNCH_CHECK();
  return VS_OK;
}

}  // namespace vs
