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

namespace vs {
namespace {

using namespace umma;

constexpr int kLoaderWarps = 4;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 32 * (kLoaderWarps + 2 + kEpiWarps);   // loaders | weight producer | MMA | epilogue
constexpr int kSA = 2, kSB = 3;
// plain TF32: 96 channels per activation stage, 32 per weight slab; 3xTF32: 48 / 16 with [hi|lo] pairs (same bytes)

struct Plan {
  int rows_a, halo_l, n_ka, slabs_per_ka, Nblk, NB, NACC, tmem_cols, n_tiles, n_units;
  int NS, Npack;          // the weights are packed in NB blocks of Npack columns; a unit computes Nblk = Npack / NS of them
                          // (NS > 1 when there are too few row tiles to fill the GPU: phoneme-level convs)
  uint32_t bp_half, bp_bytes;   // packed slab geometry (bytes of one [hi] half / of the whole slab in HBM)
  int KA, slabC;          // channels per activation stage / per weight slab
  uint32_t a_half, b_half; // byte offset of the lo copy inside a stage / slab (split3)
  uint32_t a_bytes, b_bytes, smem_bytes, off_b, off_bar;
};
struct Params {
  UmmaTf32 c;
  Plan p;
};

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
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t a_base = smem_base, b_base = smem_base + p.off_b, bar_base = smem_base + p.off_bar;
  auto a_full = [&](int i) { return bar_base + 8u * i; };
  auto a_empty = [&](int i) { return bar_base + 8u * (4 + i); };
  auto b_full = [&](int i) { return bar_base + 8u * (8 + i); };
  auto b_empty = [&](int i) { return bar_base + 8u * (12 + i); };
  auto acc_full = [&](int i) { return bar_base + 8u * (16 + i); };
  auto acc_empty = [&](int i) { return bar_base + 8u * (24 + i); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + p.off_bar + 8 * 32);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kSA; ++i) { mbar_init(a_full(i), kLoaderWarps); mbar_init(a_empty(i), 1); }
    for (int i = 0; i < kSB; ++i) { mbar_init(b_full(i), 1); mbar_init(b_empty(i), 1); }
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
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int slabs_per_unit = p.n_ka * c.taps * p.slabs_per_ka;

  if (warp < kLoaderWarps) {
    // ------------------------------------------------------------- activation loaders (128 threads)
    const int tid = threadIdx.x;                       // 0..127
    uint32_t slot = 0, phase = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const int tile = u / (p.NB * p.NS);
      const int row_lo = tile * kTileM - p.halo_l;
      for (int ka = 0; ka < p.n_ka; ++ka) {
        mbar_wait(a_empty(slot), phase ^ 1, 11);
        const uint32_t stage = a_base + slot * p.a_bytes;
        const int ch0 = ka * p.KA;
        const int n_planes = min(p.KA, c.Cin - ch0) / 4;
        for (int row = tid; row < p.rows_a; row += 32 * kLoaderWarps) {
          const int rg = row_lo + row;
          const bool ok = rg >= 0 && rg < c.R;
          const float* src = c.in + (size_t)(ok ? rg : 0) * c.in_ld + ch0;
          const uint32_t dst = stage + (uint32_t)row * 16u;
#pragma unroll 8
          for (int pl = 0; pl < n_planes; ++pl) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok) v = *reinterpret_cast<const float4*>(src + 4 * pl);
            const float4 hi = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (uint32_t)(pl * p.rows_a) * 16u),
                         "f"(hi.x), "f"(hi.y), "f"(hi.z), "f"(hi.w)
                         : "memory");
            if (c.split3)
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + p.a_half + (uint32_t)(pl * p.rows_a) * 16u),
                           "f"(to_tf32(v.x - hi.x)), "f"(to_tf32(v.y - hi.y)), "f"(to_tf32(v.z - hi.z)), "f"(to_tf32(v.w - hi.w))
                           : "memory");
          }
        }
        fence_proxy_async();                         // generic-proxy smem writes -> visible to the tensor core
        __syncwarp();
        if (lane == 0) mbar_arrive(a_full(slot));
        if (++slot == kSA) { slot = 0; phase ^= 1; }
      }
    }
  } else if (warp == kLoaderWarps) {
    // ------------------------------------------------------------- weight slab producer (TMA bulk)
    if (lane == 0) {
      uint32_t slot = 0, phase = 0;
      const int halves = c.split3 ? 2 : 1, planes = p.slabC / 4;
      for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
        const int nbe = u % (p.NB * p.NS), nb = nbe / p.NS, sub = nbe % p.NS;
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(c.w) + (size_t)nb * slabs_per_unit * p.bp_bytes;
        for (int s = 0; s < slabs_per_unit; ++s) {
          mbar_wait(b_empty(slot), phase ^ 1, 12);
          mbar_arrive_expect_tx(b_full(slot), p.b_bytes);
          if (p.NS == 1) {
            bulk_g2s(b_base + slot * p.b_bytes, wsrc + (size_t)s * p.bp_bytes, p.b_bytes, b_full(slot));
          } else {   // column sub-block of a packed slab: one copy per 4-channel plane (Nblk * 16 B each)
            for (int hf = 0; hf < halves; ++hf)
              for (int pl = 0; pl < planes; ++pl)
                bulk_g2s(b_base + slot * p.b_bytes + hf * p.b_half + (uint32_t)(pl * p.Nblk) * 16u,
                         wsrc + (size_t)s * p.bp_bytes + hf * p.bp_half + (size_t)(pl * p.Npack + sub * p.Nblk) * 16u,
                         (uint32_t)p.Nblk * 16u, b_full(slot));
          }
          if (++slot == kSB) { slot = 0; phase ^= 1; }
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
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      mbar_wait(acc_empty(acc_slot), acc_phase ^ 1, 13);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc_slot * (uint32_t)p.Nblk;
      uint32_t accumulate = 0;
      for (int ka = 0; ka < p.n_ka; ++ka) {
        mbar_wait(a_full(a_slot), a_phase, 14);
        tc_fence_after();
        const uint32_t a_stage16 = (a_base + a_slot * p.a_bytes) >> 4;
        const int slabs_here = min(p.KA, c.Cin - ka * p.KA) / p.slabC;
        for (int t = 0; t < c.taps; ++t)
          for (int j = 0; j < slabs_here; ++j) {
            mbar_wait(b_full(b_slot), b_phase, 15);
            tc_fence_after();
            uint32_t a_lo = a_lo_fixed + a_stage16 + (uint32_t)j * slab_planes * (uint32_t)p.rows_a + (uint32_t)(t * c.dil);
            uint32_t b_lo = b_lo_fixed + ((b_base + b_slot * p.b_bytes) >> 4);
            for (int k8 = 0; k8 < p.slabC / 8; ++k8) {
              tc_mma_tf32_lohi(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, accumulate);
              accumulate = 1;
              if (c.split3) {
                tc_mma_tf32_lohi(d_tmem, a_lo + (p.a_half >> 4), a_hi, b_lo, b_hi, idesc, 1u);   // a_lo * w_hi
                tc_mma_tf32_lohi(d_tmem, a_lo, a_hi, b_lo + (p.b_half >> 4), b_hi, idesc, 1u);   // a_hi * w_lo
              }
              a_lo += a_kstep;
              b_lo += b_kstep;
            }
            tc_commit(b_empty(b_slot));
            if (++b_slot == kSB) { b_slot = 0; b_phase ^= 1; }
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
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const int nbe = u % (p.NB * p.NS), tile = u / (p.NB * p.NS), nb = nbe / p.NS;
      const int nb_col0 = nb * p.Npack + (nbe % p.NS) * p.Nblk;
      const int r = tile * kTileM + q * 32 + lane;
      const bool in_range = r < c.R;
      int utt = -1;
      if (in_range) utt = c.row_utt ? c.row_utt[r] : 0;
      const bool valid = utt >= 0;
      const float* ub = nullptr;
      if (c.ubias && valid) ub = c.ubias + (size_t)(c.ubias_idx ? c.ubias_idx[utt] : utt) * c.ubias_ld;
      mbar_wait(acc_full(acc_slot), acc_phase, 16);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc_slot * (uint32_t)p.Nblk;
      for (int cc = hsel; cc < n_chunks; cc += 2) {
        uint32_t v[32];
        tmem_ld32(t_row + (uint32_t)(cc * 32), v);
        if (in_range) {
          const int col0 = nb_col0 + cc * 32;
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

  tc_fence_before();
  __syncthreads();
  if (warp == kLoaderWarps + 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

int make_plan(const UmmaTf32& c, int n_sm, Plan* out) {
  Plan p{};
  VS_REQUIRE(c.N % 32 == 0, "umma_tf32: N=%d must be a multiple of 32", c.N);
  VS_REQUIRE(c.R > 0 && c.taps >= 1 && c.dil >= 1 && c.pad_l >= 0, "umma_tf32: bad shape");
  VS_REQUIRE(c.in_ld % 4 == 0 && c.out_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(c.in) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(c.out) & 15) == 0,
             "umma_tf32: rows must be 16-byte aligned");
  VS_REQUIRE(!c.ubias || (c.ubias_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(c.ubias) & 15) == 0),
             "umma_tf32: per-speaker bias must be 16-byte aligned");
  p.Nblk = 0;
  for (int nb = 1; nb <= 8 && !p.Nblk; ++nb)                    // fewest n-blocks with Nblk <= 256, multiple of 32
    if (c.N % nb == 0 && c.N / nb <= 256 && (c.N / nb) % 32 == 0) p.Nblk = c.N / nb;
  VS_REQUIRE(p.Nblk > 0, "umma_tf32: cannot split N=%d into <= 256-column blocks", c.N);
  p.NB = c.N / p.Nblk;
  p.Npack = p.Nblk;
  p.n_tiles = (c.R + kTileM - 1) / kTileM;
  p.NS = 1;
  if (c.epi == 0 && n_sm > 0) {
    // too few row tiles for the GPU (phoneme level: 22 tiles) or a ragged last wave: let several CTAs share a row tile, each
    // computing a column sub-block.  Modelled time = waves * (1 / NS + the per-unit cost of loading the A tile again).
    double best = 1e30;
    for (int ns = 1; ns <= 8; ++ns) {
      if (p.Npack % ns || (p.Npack / ns) % 32) continue;
      const int units = p.n_tiles * p.NB * ns;
      const double cost = (double)((units + n_sm - 1) / n_sm) * (1.0 / ns + 0.15);
      if (cost < best - 1e-9) { best = cost; p.NS = ns; }
    }
    p.Nblk = p.Npack / p.NS;
  }
  p.KA = c.split3 ? 48 : 96;
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
  p.bp_half = (uint32_t)p.slabC * p.Npack * 4u;
  p.bp_bytes = p.bp_half * (c.split3 ? 2u : 1u);
  p.NACC = 512 / p.Nblk;
  if (p.NACC > 8) p.NACC = 8;
  VS_REQUIRE(p.NACC >= 2, "umma_tf32: TMEM too small");
  int cols = 32;
  while (cols < p.NACC * p.Nblk) cols *= 2;
  p.tmem_cols = cols;
  p.off_b = kSA * p.a_bytes;
  p.off_bar = (p.off_b + kSB * p.b_bytes + 127u) & ~127u;
  p.smem_bytes = p.off_bar + 8u * 32 + 16u;
  VS_REQUIRE(p.smem_bytes <= 227u * 1024, "umma_tf32: tile does not fit in shared memory");
  if (p.smem_bytes < 120u * 1024) p.smem_bytes = 120u * 1024;   // one CTA per SM (it owns all 512 TMEM columns)
  p.n_units = p.n_tiles * p.NB * p.NS;
  *out = p;
  return VS_OK;
}

}  // namespace

int umma_tf32(const UmmaTf32& c, cudaStream_t st) {
  Params prm;
  prm.c = c;
  VS_REQUIRE(c.in && c.w && c.out, "umma_tf32: null pointer");
  VS_REQUIRE(c.epi != 2 || (c.out2 && c.out2_ld % 4 == 0), "umma_tf32: epi=2 needs out2");
  static int n_sm = 0;
  static bool configured = false;
  if (!configured) {
    int dev = 0;
    VS_CUDA_CHECK(cudaGetDevice(&dev));
    VS_CUDA_CHECK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    VS_CUDA_CHECK(cudaFuncSetAttribute(umma_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  VS_TRY(make_plan(c, n_sm, &prm.p));
  int grid = n_sm < prm.p.n_units ? n_sm : prm.p.n_units;
  umma_tf32_kernel<<<grid, kThreads, prm.p.smem_bytes, st>>>(prm);
  VS_LAUNCH_CHECK();
  return VS_OK;
}

}  // namespace vs
