// conv1d over ragged rows on the tensor cores in TF32 (tcgen05.mma kind::tf32, fp32 accumulate in TMEM), for the
// part of the path that must stay above bf16 precision: flow WN convs, encoder QKV / O / FFN convs,
// projection.  fp32 row-major [R][C] in and out, so the surrounding memory-bound kernels are unchanged.
//
//   * 4 loader warps stage the activation tile: row-major fp32 -> registers (round-to-nearest TF32) -> the UMMA
//     "K-major, no swizzle" smem layout [channel-quad plane][row][4 x fp32]; rows outside [0,R) become the conv's zero
//     padding; taps are descriptor start offsets of tap*dil rows, as in umma_conv.cu.  Channels are staged in chunks
//     of 96 (24 planes) so that Cin = 768 fits; a 2-deep ring overlaps staging with the MMAs.
//   * 1 producer warp streams pre-rounded TF32 weight slabs (32 channels x Nblk columns, one cp.async.bulk each)
//     through a 3-deep ring; 1 warp issues the MMAs (warp-uniform loop, elected lane); 8 epilogue warps drain the
//     TMEM accumulator ring: + bias, + per-speaker bias (WN cond), ReLU, validity mask, 128 B contiguous stores per lane.
//   * Work unit = (128-row tile, n-block <= 256 columns); persistent CTAs stride over units.
#include "umma_tf32.cuh"
#include "umma_common.cuh"
#include "umma_conv.cuh"

namespace vs {
namespace {

using namespace umma;

constexpr int kLoaderWarps = 4;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 32 * (kLoaderWarps + 2 + kEpiWarps);   // loaders | weight producer | MMA | epilogue
constexpr int kSA = 2, kMaxSB = 16;   // weight ring: as many slabs as fit (>= 3): the stream is latency-bound otherwise
// plain TF32: 96 channels per activation stage, 32 per weight slab; 3xTF32: 48 / 16 with [hi|lo] pairs (same bytes)

// option "tf32_cluster" (1 | 2): CTAs sharing each weight slab by TMA multicast (off: no gain)

struct Plan {
  int rows_a, halo_l, n_ka, slabs_per_ka, Nblk, NB, NACC, tmem_cols, n_tiles, n_units;
  int cl;                 // CTAs per cluster sharing every weight slab (1 or 2); n_units counts (tile group of cl tiles, n-block)
  int SB;                 // weight ring depth
  int KA, slabC;          // channels per activation stage / per weight slab
  uint32_t a_half, b_half; // byte offset of the lo copy inside a stage / slab (split3)
  uint32_t a_bytes, b_bytes, smem_bytes, off_b, off_bar;
};
struct Params {
  UmmaTf32 c;
  Plan p;
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

__device__ __forceinline__ void tc_mma_tf32_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                                 uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, pe;\n\t.reg .b64 da, db;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Weight slabs are re-streamed from L2 for every 128-row tile: at 27.6 k rows the frame-level convs move 0.3-0.8 GB of weights
// per launch (FFN-1: 764 MB in 139 us).  With cl = 2 the two CTAs of a cluster work on neighbouring row tiles of the same
// n-block in lockstep; each fetches HALF of every slab and multicasts it into both CTAs' rings, and a ring slot is released
// by a multicast tcgen05.commit from both MMA warps.  Measured: no gain (frame prior 3.42 vs 3.41 ms, flow 2.13 vs 2.10) - L2
// is not what these kernels wait for - so it is an A/B knob, off by default.  (A first, broken version - UMMA descriptors
// built from un-masked shared-window addresses, which carry the CTA rank above bit 18 in a cluster launch - "gained" 8 % on
// the whole step: its MMAs multiplied garbage, drew less power, and the power-capped GPU clocked 1.89 instead of 1.73 GHz.)
__device__ __forceinline__ void bulk_g2s_multicast(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar), "h"(mask)
               : "memory");
}
__device__ __forceinline__ void tc_commit_multicast(uint32_t bar, uint16_t mask) {   // whole warp calls, one elected lane commits
  asm volatile(
      "{\n\t.reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(bar), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// a/b format TF32 = 2 (cute::UMMA::F16F32Format), fp32 accumulate, K-major both
__device__ __forceinline__ uint32_t make_idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
}
__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

__global__ void __launch_bounds__(kThreads, 1) umma_tf32_kernel(const __grid_constant__ Params prm) {
  extern __shared__ __align__(128) uint8_t smem[];
  const UmmaTf32& c = prm.c;
  const Plan& p = prm.p;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
#ifdef VS_UMMA_TIMING
  long long* const dbg = prm.dbg;
  long long tw0 = 0, tw1 = 0, tw2 = 0;
  const long long t_start = dbg ? clock64() : 0;
#endif
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t a_base = smem_base, b_base = smem_base + p.off_b, bar_base = smem_base + p.off_bar;
  auto a_full = [&](int i) { return bar_base + 8u * i; };
  auto a_empty = [&](int i) { return bar_base + 8u * (4 + i); };
  auto acc_full = [&](int i) { return bar_base + 8u * (8 + i); };
  auto acc_empty = [&](int i) { return bar_base + 8u * (16 + i); };
  auto b_full = [&](int i) { return bar_base + 8u * (24 + i); };
  auto b_empty = [&](int i) { return bar_base + 8u * (24 + kMaxSB + i); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + p.off_bar + 8 * (24 + 2 * kMaxSB));

  if (threadIdx.x == 0) {
    for (int i = 0; i < kSA; ++i) { mbar_init(a_full(i), kLoaderWarps); mbar_init(a_empty(i), 1); }
    for (int i = 0; i < p.SB; ++i) { mbar_init(b_full(i), 1); mbar_init(b_empty(i), (uint32_t)p.cl); }
    for (int i = 0; i < p.NACC; ++i) { mbar_init(acc_full(i), 1); mbar_init(acc_empty(i), kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kLoaderWarps + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (p.cl > 1) cluster_sync_all();               // the peer's barriers exist before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int slabs_per_unit = p.n_ka * c.taps * p.slabs_per_ka;
  // work item w = (group of cl neighbouring row tiles, n-block); this CTA takes tile `rank` of the group
  const int cl = p.cl, rank = cl > 1 ? (int)(blockIdx.x % cl) : 0;
  const int w_first = blockIdx.x / cl, w_step = gridDim.x / cl;
  auto tile_of = [&](int w) { return (w / p.NB) * cl + rank; };

  if (warp < kLoaderWarps) {
    // ------------------------------------------------------------- activation loaders (128 threads)
    // Thread t owns row t of every stage (plus one halo row for the first few threads).  The global loads of stage i+1 are
    // issued into registers right after stage i has been written to shared memory, so their latency overlaps the wait
    // for the ring slot (i.e. the MMAs of the stages before) instead of adding ~1.5 us to every stage.
    const int tid = threadIdx.x;                       // 0..127
    constexpr int kPF = 12;                            // float4 per thread held in flight (= planes of a 48-channel stage)
    uint32_t slot = 0, phase = 0;
    float4 pf[kPF];
    auto stage_src = [&](int u, int ka, int row, bool* ok) -> const float* {
      const int tile = tile_of(u);
      const int rg = tile * kTileM - p.halo_l + row;
      *ok = rg >= 0 && rg < c.R;
      return c.in + (size_t)(*ok ? rg : 0) * c.in_ld + ka * p.KA;
    };
    auto prefetch = [&](int u, int ka) {
      bool ok;
      const float* src = stage_src(u, ka, tid, &ok);
      const int n_planes = min(p.KA, c.Cin - ka * p.KA) / 4;
#pragma unroll
      for (int pl = 0; pl < kPF; ++pl) {
        pf[pl] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok && pl < n_planes) pf[pl] = *reinterpret_cast<const float4*>(src + 4 * pl);
      }
    };
    auto put = [&](uint32_t dst, const float4& v) {
      const float4 hi = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(hi.x), "f"(hi.y), "f"(hi.z), "f"(hi.w) : "memory");
      if (c.split3)
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + p.a_half), "f"(to_tf32(v.x - hi.x)),
                     "f"(to_tf32(v.y - hi.y)), "f"(to_tf32(v.z - hi.z)), "f"(to_tf32(v.w - hi.w))
                     : "memory");
    };
    int u = w_first, ka = 0;
    if (u < p.n_units) prefetch(u, 0);
    while (u < p.n_units) {
      VS_TIMED(tw0, mbar_wait(a_empty(slot), phase ^ 1, 11));
      const uint32_t stage = a_base + slot * p.a_bytes;
      const int n_planes = min(p.KA, c.Cin - ka * p.KA) / 4;
#pragma unroll
      for (int pl = 0; pl < kPF; ++pl)
        if (pl < n_planes) put(stage + (uint32_t)tid * 16u + (uint32_t)(pl * p.rows_a) * 16u, pf[pl]);
      // planes beyond the prefetch depth (96-channel stages) and the halo rows: loaded here
      for (int row = tid; row < p.rows_a; row += 32 * kLoaderWarps) {
        bool ok;
        const float* src = stage_src(u, ka, row, &ok);
        for (int pl = (row == tid ? kPF : 0); pl < n_planes; ++pl) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ok) v = *reinterpret_cast<const float4*>(src + 4 * pl);
          put(stage + (uint32_t)row * 16u + (uint32_t)(pl * p.rows_a) * 16u, v);
        }
      }
      fence_proxy_async();                         // generic-proxy smem writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full(slot));
      if (++slot == kSA) { slot = 0; phase ^= 1; }
      if (++ka == p.n_ka) { ka = 0; u += w_step; }
      if (u < p.n_units) prefetch(u, ka);
    }
  } else if (warp == kLoaderWarps) {
    // ------------------------------------------------------------- weight slab producer (TMA bulk)
    {
      // Every mbarrier probe is a ~250 clk shared-memory round trip while the tensor pipe runs, and a weight slab is only
      // 6-24 MMAs of work: the ring slots two and one ahead are probed while the current slab is being issued.
      uint32_t slot = 0, phase = 0;
      const uint32_t SB = (uint32_t)p.SB;
      auto ahead = [&](uint32_t sl, uint32_t ph, uint32_t d, uint32_t* s2, uint32_t* p2) {
        sl += d;
        if (sl >= SB) { sl -= SB; ph ^= 1; }
        *s2 = sl; *p2 = ph;
      };
      uint32_t s1, p1, s2, p2;
      ahead(slot, phase, 1 % SB, &s1, &p1);
      bool r0 = mbar_test_wait(b_empty(slot), phase ^ 1), r1 = SB > 1 && mbar_test_wait(b_empty(s1), p1 ^ 1);
      const uint32_t part = p.b_bytes / (uint32_t)cl;       // this CTA's share of every slab
      for (int u = w_first; u < p.n_units; u += w_step) {
        const int nb = u % p.NB;
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(c.w) + (size_t)nb * slabs_per_unit * p.b_bytes;
        for (int s = 0; s < slabs_per_unit; ++s) {
          if (!r0) VS_TIMED(tw0, mbar_wait(b_empty(slot), phase ^ 1, 12));
          ahead(slot, phase, 2, &s2, &p2);
          const bool r2 = SB > 2 && mbar_test_wait(b_empty(s2), p2 ^ 1);
          if (lane == 0) {
            mbar_arrive_expect_tx(b_full(slot), p.b_bytes);
            if (cl == 1) bulk_g2s(b_base + slot * p.b_bytes, wsrc + (size_t)s * p.b_bytes, p.b_bytes, b_full(slot));
            else bulk_g2s_multicast(b_base + slot * p.b_bytes + rank * part, wsrc + (size_t)s * p.b_bytes + rank * part, part,
                                    b_full(slot), (uint16_t)((1u << cl) - 1u));
          }
          ahead(slot, phase, 1, &slot, &phase);
          r0 = r1; r1 = r2;
        }
      }
    }
    __syncwarp();
  } else if (warp == kLoaderWarps + 1) {
    // ------------------------------------------------------------- MMA issuer
    const uint32_t idesc = make_idesc_tf32(p.Nblk);
    const uint32_t a_lbo = (uint32_t)p.rows_a * 16u, b_lbo = (uint32_t)p.Nblk * 16u;
    const uint32_t a_hi = (uint32_t)(make_desc(0, a_lbo, 128u) >> 32), b_hi = (uint32_t)(make_desc(0, b_lbo, 128u) >> 32);
    const uint32_t a_lo_fixed = (uint32_t)make_desc(0, a_lbo, 128u), b_lo_fixed = (uint32_t)make_desc(0, b_lbo, 128u);
    const uint32_t a_kstep = 2u * (uint32_t)p.rows_a, b_kstep = 2u * (uint32_t)p.Nblk;
    const uint32_t slab_planes = (uint32_t)p.slabC / 4;
    uint32_t a_slot = 0, a_phase = 0, b_slot = 0, b_phase = 0, acc_slot = 0, acc_phase = 0;
    const uint32_t SB = (uint32_t)p.SB;
    auto ahead = [&](uint32_t sl, uint32_t ph, uint32_t d, uint32_t* s2, uint32_t* p2) {
      sl += d;
      if (sl >= SB) { sl -= SB; ph ^= 1; }
      *s2 = sl; *p2 = ph;
    };
    uint32_t bs1, bp1, bs2, bp2;
    ahead(b_slot, b_phase, 1 % SB, &bs1, &bp1);
    bool r0 = mbar_test_wait(b_full(b_slot), b_phase), r1 = SB > 1 && mbar_test_wait(b_full(bs1), bp1);
    for (int u = w_first; u < p.n_units; u += w_step) {
      VS_TIMED(tw1, mbar_wait(acc_empty(acc_slot), acc_phase ^ 1, 13));
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc_slot * (uint32_t)p.Nblk;
      uint32_t accumulate = 0;
      for (int ka = 0; ka < p.n_ka; ++ka) {
        VS_TIMED(tw0, mbar_wait(a_full(a_slot), a_phase, 14));
        tc_fence_after();
        // descriptors take the 18-bit offset inside this CTA's shared memory: in a cluster launch the shared-window address
        // of rank > 0 has higher bits set, which an unmasked add would carry into the LBO field
        const uint32_t a_stage16 = ((a_base + a_slot * p.a_bytes) & 0x3FFFFu) >> 4;
        const int slabs_here = min(p.KA, c.Cin - ka * p.KA) / p.slabC;
        for (int t = 0; t < c.taps; ++t)
          for (int j = 0; j < slabs_here; ++j) {
            if (!r0) VS_TIMED(tw2, mbar_wait(b_full(b_slot), b_phase, 15));
            ahead(b_slot, b_phase, 2, &bs2, &bp2);
            const bool r2 = SB > 2 && mbar_test_wait(b_full(bs2), bp2);      // probe two slabs ahead (see the producer)
            tc_fence_after();
            uint32_t a_lo = a_lo_fixed + a_stage16 + (uint32_t)j * slab_planes * (uint32_t)p.rows_a + (uint32_t)(t * c.dil);
            uint32_t b_lo = b_lo_fixed + (((b_base + b_slot * p.b_bytes) & 0x3FFFFu) >> 4);
            // fully unrolled per-slab issue (one warp issues every MMA: runtime-nested loops cost ~100 clk per MMA,
            // tools/mma_microbench.cu): 3xTF32 = 2 K-steps x {hi*hi, lo*hi, hi*lo}, plain = 4 K-steps
            if (c.split3) {
              const uint32_t a_lo2 = a_lo + (p.a_half >> 4), b_lo2 = b_lo + (p.b_half >> 4);
#pragma unroll
              for (int k8 = 0; k8 < 2; ++k8) {
                tc_mma_tf32_lohi(d_tmem, a_lo + k8 * a_kstep, a_hi, b_lo + k8 * b_kstep, b_hi, idesc, k8 ? 1u : accumulate);
                tc_mma_tf32_lohi(d_tmem, a_lo2 + k8 * a_kstep, a_hi, b_lo + k8 * b_kstep, b_hi, idesc, 1u);   // a_lo * w_hi
                tc_mma_tf32_lohi(d_tmem, a_lo + k8 * a_kstep, a_hi, b_lo2 + k8 * b_kstep, b_hi, idesc, 1u);   // a_hi * w_lo
              }
            } else {
#pragma unroll
              for (int k8 = 0; k8 < 4; ++k8)
                tc_mma_tf32_lohi(d_tmem, a_lo + k8 * a_kstep, a_hi, b_lo + k8 * b_kstep, b_hi, idesc, k8 ? 1u : accumulate);
            }
            accumulate = 1;
            if (cl == 1) tc_commit(b_empty(b_slot));
            else tc_commit_multicast(b_empty(b_slot), (uint16_t)((1u << cl) - 1u));   // releases the slot in both CTAs
            ahead(b_slot, b_phase, 1, &b_slot, &b_phase);
            r0 = r1; r1 = r2;
          }
        tc_commit(a_empty(a_slot));
        if (++a_slot == kSA) { a_slot = 0; a_phase ^= 1; }
      }
      tc_commit(acc_full(acc_slot));
      if (++acc_slot == (uint32_t)p.NACC) { acc_slot = 0; acc_phase ^= 1; }
    }
  } else {
    // ------------------------------------------------------------- epilogue (8 warps)
    const int ew = warp - (kLoaderWarps + 2);
    const int q = warp & 3;                      // TMEM lane quarter this warp may access
    const int hsel = ew >> 2;
    const int n_chunks = p.Nblk / 32;
    uint32_t acc_slot = 0, acc_phase = 0;
    for (int u = w_first; u < p.n_units; u += w_step) {
      const int tile = tile_of(u), nb = u % p.NB;
      const int r = tile * kTileM + q * 32 + lane;
      const bool in_range = r < c.R;
      int utt = -1;
      if (in_range) utt = c.row_utt ? c.row_utt[r] : 0;
      const bool valid = utt >= 0;
      const float* ub = nullptr;
      if (c.ubias && valid) ub = c.ubias + (size_t)(c.ubias_idx ? c.ubias_idx[utt] : utt) * c.ubias_ld;
      VS_TIMED(tw0, mbar_wait(acc_full(acc_slot), acc_phase, 16));
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc_slot * (uint32_t)p.Nblk;
      for (int cc = hsel; cc < n_chunks; cc += 2) {
        uint32_t v[32];
        tmem_ld32(t_row + (uint32_t)(cc * 32), v);
        if (in_range) {
          const int col0 = nb * p.Nblk + cc * 32;
          if (c.epi == 1) {
            // WN gate: this chunk = [16 tanh pre-activations | 16 sigmoid pre-activations] of channels ch0..ch0+15
            float* o = c.out + (size_t)r * c.out_ld + (col0 >> 1);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
              if (valid) {
                float t[4], sg[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  t[e] = __uint_as_float(v[4 * g + e]);
                  sg[e] = __uint_as_float(v[16 + 4 * g + e]);
                }
                if (c.bias) {
                  const float4 bt = __ldg(reinterpret_cast<const float4*>(c.bias + col0 + 4 * g));
                  const float4 bs = __ldg(reinterpret_cast<const float4*>(c.bias + col0 + 16 + 4 * g));
                  t[0] += bt.x; t[1] += bt.y; t[2] += bt.z; t[3] += bt.w;
                  sg[0] += bs.x; sg[1] += bs.y; sg[2] += bs.z; sg[3] += bs.w;
                }
                if (ub) {
                  const float4 bt = __ldg(reinterpret_cast<const float4*>(ub + col0 + 4 * g));
                  const float4 bs = __ldg(reinterpret_cast<const float4*>(ub + col0 + 16 + 4 * g));
                  t[0] += bt.x; t[1] += bt.y; t[2] += bt.z; t[3] += bt.w;
                  sg[0] += bs.x; sg[1] += bs.y; sg[2] += bs.z; sg[3] += bs.w;
                }
                y.x = tanhf(t[0]) * (1.f / (1.f + expf(-sg[0])));
                y.y = tanhf(t[1]) * (1.f / (1.f + expf(-sg[1])));
                y.z = tanhf(t[2]) * (1.f / (1.f + expf(-sg[2])));
                y.w = tanhf(t[3]) * (1.f / (1.f + expf(-sg[3])));
              }
              *reinterpret_cast<float4*>(o + 4 * g) = y;
            }
          } else if (c.epi == 2) {
            const bool to_h = nb < c.nb_split;
            float* o = to_h ? c.out + (size_t)r * c.out_ld + col0
                            : c.out2 + (size_t)r * c.out2_ld + (col0 - c.nb_split * p.Nblk);
            const bool add = to_h || c.accumulate2;
            if (valid) {
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                float4 y = make_float4(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]), __uint_as_float(v[4 * g + 2]),
                                       __uint_as_float(v[4 * g + 3]));
                if (c.bias) {
                  const float4 b = __ldg(reinterpret_cast<const float4*>(c.bias + col0 + 4 * g));
                  y.x += b.x; y.y += b.y; y.z += b.z; y.w += b.w;
                }
                if (add) {
                  const float4 old = *reinterpret_cast<const float4*>(o + 4 * g);
                  y.x += old.x; y.y += old.y; y.z += old.z; y.w += old.w;
                }
                *reinterpret_cast<float4*>(o + 4 * g) = y;
              }
            } else if (!add) {
#pragma unroll
              for (int g = 0; g < 8; ++g) *reinterpret_cast<float4*>(o + 4 * g) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
          } else {
          float* o = c.out + (size_t)r * c.out_ld + col0;
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) {
              y = make_float4(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]), __uint_as_float(v[4 * g + 2]),
                              __uint_as_float(v[4 * g + 3]));
              if (c.bias) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(c.bias + col0 + 4 * g));
                y.x += b.x; y.y += b.y; y.z += b.z; y.w += b.w;
              }
              if (ub) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(ub + col0 + 4 * g));
                y.x += b.x; y.y += b.y; y.z += b.z; y.w += b.w;
              }
              if (c.act == 1) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f); }
            }
            *reinterpret_cast<float4*>(o + 4 * g) = y;
          }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(acc_slot));
      if (++acc_slot == (uint32_t)p.NACC) { acc_slot = 0; acc_phase ^= 1; }
    }
  }

#ifdef VS_UMMA_TIMING
  if (dbg && lane == 0 && (warp == 0 || (warp >= kLoaderWarps && warp <= kLoaderWarps + 2))) {
    // [cta][loader | weight producer | MMA | first epilogue warp][total, waits]
    long long* o = dbg + ((size_t)blockIdx.x * 4 + (warp == 0 ? 0 : warp - kLoaderWarps + 1)) * 4;
    o[0] = clock64() - t_start; o[1] = tw0; o[2] = tw1; o[3] = tw2;
  }
#endif
  tc_fence_before();
  __syncthreads();
  if (p.cl > 1) cluster_sync_all();               // no CTA leaves while its peer may still multicast into it
  if (warp == kLoaderWarps + 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

int make_plan(const UmmaTf32& c, Plan* out) {
  Plan p{};
  VS_REQUIRE(c.N % 32 == 0, "umma_tf32: N=%d must be a multiple of 32", c.N);
  VS_REQUIRE(c.R > 0 && c.taps >= 1 && c.dil >= 1 && c.pad_l >= 0, "umma_tf32: bad shape");
  VS_REQUIRE(c.in_ld % 4 == 0 && c.out_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(c.in) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(c.out) & 15) == 0,
             "umma_tf32: rows must be 16-byte aligned");
  VS_REQUIRE(!c.ubias || (c.ubias_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(c.ubias) & 15) == 0),
             "umma_tf32: per-speaker bias must be 16-byte aligned");
  p.Nblk = 0;
  for (int nb = 1; nb <= 16 && !p.Nblk; ++nb)                   // fewest n-blocks with Nblk <= 256, multiple of 32
    if (c.N % nb == 0 && c.N / nb <= 256 && (c.N / nb) % 32 == 0) p.Nblk = c.N / nb;
  VS_REQUIRE(p.Nblk > 0, "umma_tf32: cannot split N=%d into <= 256-column blocks", c.N);
  p.NB = c.N / p.Nblk;
  p.n_tiles = (c.R + kTileM - 1) / kTileM;
  p.KA = c.split3 ? 48 : 96;                                      // must match packing.pack_tf32
  p.slabC = c.split3 ? 16 : 32;
  VS_REQUIRE(c.Cin % p.slabC == 0, "umma_tf32: Cin=%d must be a multiple of %d", c.Cin, p.slabC);
  p.n_ka = (c.Cin + p.KA - 1) / p.KA;
  p.slabs_per_ka = p.KA / p.slabC;
  VS_REQUIRE(c.Cin % p.KA == 0 || c.Cin < p.KA, "umma_tf32: Cin=%d must be < or a multiple of %d", c.Cin, p.KA);
  if (c.Cin < p.KA) p.slabs_per_ka = c.Cin / p.slabC;
  p.halo_l = c.pad_l * c.dil;
  p.rows_a = kTileM + (c.taps - 1) * c.dil;
  const int planes_stage = (c.Cin < p.KA ? c.Cin : p.KA) / 4;
  p.a_half = (uint32_t)planes_stage * p.rows_a * 16u;
  p.a_bytes = p.a_half * (c.split3 ? 2u : 1u);
  p.b_half = (uint32_t)p.slabC * p.Nblk * 4u;
  p.b_bytes = p.b_half * (c.split3 ? 2u : 1u);
  p.NACC = 512 / p.Nblk;
  if (p.NACC > 8) p.NACC = 8;
  VS_REQUIRE(p.NACC >= 2, "umma_tf32: TMEM too small");
  int cols = 32;
  while (cols < p.NACC * p.Nblk) cols *= 2;
  p.tmem_cols = cols;
  p.off_b = kSA * p.a_bytes;
  const uint32_t bar_bytes = 8u * (24 + 2 * kMaxSB) + 16u;
  p.SB = (int)((220u * 1024 - p.off_b - bar_bytes) / p.b_bytes);
  if (p.SB > kMaxSB) p.SB = kMaxSB;
  VS_REQUIRE(p.SB >= 2, "umma_tf32: tile does not fit in shared memory");
  p.off_bar = (p.off_b + p.SB * p.b_bytes + 127u) & ~127u;
  p.smem_bytes = p.off_bar + bar_bytes;
  VS_REQUIRE(p.smem_bytes <= 227u * 1024, "umma_tf32: tile does not fit in shared memory");
  if (p.smem_bytes < 120u * 1024) p.smem_bytes = 120u * 1024;   // one CTA per SM (it owns all 512 TMEM columns)
  p.cl = (opts().v[OPT_TF32_CLUSTER] == 2 && p.n_tiles >= 2) ? 2 : 1;
  p.n_units = ((p.n_tiles + p.cl - 1) / p.cl) * p.NB;
  *out = p;
  return VS_OK;
}

}  // namespace

int umma_tf32(const UmmaTf32& c, cudaStream_t st) {
  Params prm;
  prm.c = c;
  VS_REQUIRE(c.in && c.w && c.out, "umma_tf32: null pointer");
  VS_REQUIRE(c.epi != 2 || (c.out2 && c.out2_ld % 4 == 0), "umma_tf32: epi=2 needs out2");
  int n_sm = 0;
  VS_TRY(device_sm_count(&n_sm));
  VS_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(umma_tf32_kernel), 227 * 1024));
  VS_TRY(make_plan(c, &prm.p));
  prm.dbg = static_cast<long long*>(umma_conv_timing_buffer());
  const int cl = prm.p.cl;
  int groups = n_sm / cl;
  if (groups > prm.p.n_units) groups = prm.p.n_units;
  const int grid = groups * cl;
  if (cl == 1) {
    umma_tf32_kernel<<<grid, kThreads, prm.p.smem_bytes, st>>>(prm);
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = prm.p.smem_bytes; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    VS_CUDA_CHECK(cudaLaunchKernelEx(&cfg, umma_tf32_kernel, prm));
  }
  VS_LAUNCH_CHECK();
  return VS_OK;
}

}  // namespace vs
