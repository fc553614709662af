// K3 / K4: Dense + bias + dropout + residual add + LayerNorm in ONE kernel (sm_100a):
//
//     z = x_res + dropout(A . W + b)        out = LayerNorm(z) * gamma + beta          (eps 1e-6, biased variance)
//
//   reference: the output projection of MultiHeadAttention followed by dropout1 / layernorm1 (Attention.py:77 ->
//   Encoder.py:22-23, Decoder.py:30-31, 35-36) and the second Dense of point_wise_feed_forward_network followed by
//   dropout / layernorm (point_ffn.py:6 -> Encoder.py:26-27, Decoder.py:39-40).
//
// A LayerNorm row is a whole row of the GEMM output, so a CTA must own complete rows: a CTA PAIR (tcgen05 cta_group::2)
// works on 256 rows x ALL N = d_model <= 512 columns -- each CTA keeps its 128 rows x N fp32 accumulators in its 512
// tensor-memory columns (N = 512: two 256-column MMA tiles, computed one after the other through the same TMA ring).
// Epilogue, two warpgroups (each owns half of the columns of every row, thread = row):
//   pass A   acc + bias -> dropout (same Philox mask and the same 1-bit keep bytes as the stand-alone kernel) -> + residual
//            = z ; row sum / sum of squares ; z -> bf16 -> smem -> TMA store (saved for the backward) ; z (fp32) is written
//            BACK into tensor memory over the accumulator
//   exchange the two warpgroups swap their partial row sums through shared memory -> mean, rstd
//   pass B   z from tensor memory -> (z - mean) * rstd * gamma + beta -> bf16 -> smem -> TMA store
// Pass A of the first column half runs while the tensor core is still working on the second half; the next row block's
// mainloop starts as soon as a half has been normalised.  The stand-alone residual_ln_fwd kernel (elementwise.cu) remains
// for d_model > 512.  What this removes per call: the GEMM's bf16 output round trip (write + read of M x N) and a kernel
// that ran at ~0.55 of the HBM copy peak.
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace bx {

constexpr int LN_BM = 128;          // rows per CTA (256 per pair)
constexpr int LN_BK = 64;
constexpr int LN_THREADS = 384;     // w0 TMA, w1 MMA, w2 TMEM alloc, w3 idle, w4-7 epilogue warpgroup 0, w8-11 warpgroup 1
constexpr int LN_STAGES = 4;

template <int BN>
struct LnCfg {
  static constexpr uint32_t A_BYTES = LN_BM * LN_BK * 2;
  static constexpr uint32_t B_BYTES = (BN / 2) * LN_BK * 2;          // this CTA's half of the B tile
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr uint32_t D_BYTES = LN_BM * 128;                   // one [128 rows x 64 bf16] staging box
  static constexpr uint32_t OFF_D = LN_STAGES * STAGE_BYTES;         // 2 warpgroups x 2 buffers
  static constexpr uint32_t OFF_X = OFF_D + 4 * D_BYTES;             // row-sum exchange: [2 parities][2 warpgroups][128 rows] float2
  static constexpr uint32_t OFF_BAR = OFF_X + 2 * 2 * 128 * 8;
  static constexpr uint32_t SMEM_BYTES = OFF_BAR + 256 + 1024;
};

// N = NT * BN output columns (NT = 1 or 2 MMA tiles per row block); A [M,K] K-major, W [K,N] MN-major (keras layout)
template <int BN, int NT>
__global__ void __launch_bounds__(LN_THREADS, 1)
gemm_ln_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmO, int M, int K,
               const float* __restrict__ bias, const __nv_bfloat16* __restrict__ xres, const float* __restrict__ gamma,
               const float* __restrict__ beta, const void* z_ptr, float* __restrict__ mean_out, float* __restrict__ rstd_out,
               uint8_t* __restrict__ keep_out, float eps, float rate, unsigned long long seed, uint32_t site,
               const unsigned long long* __restrict__ seed_dev) {
  using Cfg = LnCfg<BN>;
  constexpr int N = BN * NT;
  constexpr int WGC = BN / 2;           // columns of EACH MMA tile owned by one epilogue warpgroup (both warpgroups work on every tile:
                                        // tile 0's pass A runs under tile 1's mainloop, and only HALF of tile 1's pass A is exposed)
  constexpr int CH = WGC / 64;          // 64-column chunks per warpgroup and tile
  static_assert(WGC % 64 == 0 && N <= 512, "gemm_ln: N must be 128, 256 or 512");
  if (seed_dev != nullptr) seed = *seed_dev;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* empty_bar = full_bar + LN_STAGES;
  uint64_t* tmem_full = empty_bar + LN_STAGES;     // [2] one per MMA tile of the row block
  uint64_t* tmem_empty = tmem_full + 2;            // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float2* xch = reinterpret_cast<float2*>(smem + Cfg::OFF_X);

  const uint32_t warp = warp_idx();
  const uint32_t lane = lane_idx();
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0u;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmB); prefetch_tmap(&tmZ); prefetch_tmap(&tmO);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < LN_STAGES; ++i) { mbar_init(&full_bar[i], 2); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 256 * 2);      // the leader's barrier collects both warpgroups of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm<(N <= 128 ? 128 : (N <= 256 ? 256 : 512))>(tmem_ptr);
  tcgen05_before_thread_sync();
  cluster_sync_all();
  tcgen05_after_thread_sync();
  const uint32_t tmem_base = *tmem_ptr;
  const int unit = blockIdx.x / 2, nunits = gridDim.x / 2;
  const int m_tiles = (M + 2 * LN_BM - 1) / (2 * LN_BM);
  const int kblocks = (K + LN_BK - 1) / LN_BK;

  if (warp == 0 && lane == 0) {
    // ===================================== TMA producer =====================================
    uint32_t stage = 0, phase = 0;
    for (int t = unit; t < m_tiles; t += nunits) {
      const int m0 = t * (2 * LN_BM) + (int)rank * LN_BM;
      for (int nt = 0; nt < NT; ++nt) {
        const int n0 = nt * BN + (int)rank * (BN / 2);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          const int k0 = kb * LN_BK;
          const uint32_t lbar = mapa_u32(smem_u32(&full_bar[stage]), 0);
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
          else mbar_arrive_cluster(lbar);
          tma_load_2d_2sm(sa, &tmA, lbar, k0, m0);
#pragma unroll
          for (int p = 0; p < (BN / 2) / 64; ++p) tma_load_2d_2sm(sb + p * 8192, &tmB, lbar, n0 + p * 64, k0);
          if (++stage == LN_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && leader) {
    // ===================================== MMA issuer (leader CTA) =====================================
    constexpr uint32_t idesc = make_idesc_bf16(2 * LN_BM, BN, 0, 1);
    uint32_t stage = 0, phase = 0, it = 0;
    for (int t = unit; t < m_tiles; t += nunits, ++it) {
      for (int nt = 0; nt < NT; ++nt) {
        mbar_wait(&tmem_empty[nt], (it & 1u) ^ 1u);        // this half of the previous row block has been normalised
        tcgen05_after_thread_sync();
        const uint32_t d_tmem = tmem_base + nt * BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_after_thread_sync();
          if (lane == 0) {
            const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
            const uint32_t sb = sa + Cfg::A_BYTES;
#pragma unroll
            for (int k = 0; k < LN_BK / 16; ++k) {
              const uint64_t adesc = make_smem_desc(sa + k * 32, 0, 1024, kSwz128);
              const uint64_t bdesc = make_smem_desc(sb + k * 2048, LN_BK * 128, 1024, kSwz128);
              umma_f16_2sm(d_tmem, adesc, bdesc, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit_2sm(&empty_bar[stage]);
            if (kb == kblocks - 1) umma_commit_2sm(&tmem_full[nt]);
          }
          __syncwarp();
          if (++stage == LN_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================================== epilogue: two warpgroups, thread = row =====================================
    const uint32_t g = (warp - 4) >> 2;            // warpgroup: columns [g * WGC, (g + 1) * WGC)
    const uint32_t q = warp & 3u;                  // TMEM lane quadrant
    const uint32_t row = q * 32u + lane;
    const uint32_t wg_tid = (threadIdx.x - 128u) & 127u;
    const uint32_t lane_sel = (q * 32u) << 16;
    const uint32_t stg_base = smem_u32(smem + Cfg::OFF_D) + g * 2 * Cfg::D_BYTES;
    uint8_t* const stg_ptr = smem + Cfg::OFF_D + g * 2 * Cfg::D_BYTES;
    const float inv_keep = 1.f / (1.f - rate);
    const float inv_n = 1.f / (float)N;
    uint32_t it = 0, buf = 0;
    for (int t = unit; t < m_tiles; t += nunits, ++it) {
      const int m0 = t * (2 * LN_BM) + (int)rank * LN_BM;
      const long long grow = (long long)m0 + row;
      const bool valid = grow < M;
      float s1 = 0.f, s2 = 0.f;
      // ------------------------------ pass A: z, statistics, z -> global ------------------------------
      // residual rows of a chunk: 128 bytes per thread, requested before the accumulator chunk is read
      uint4 xr[8];
      auto load_x = [&](int col, uint4 (&dst)[8]) {
        const uint4* xp = reinterpret_cast<const uint4*>(xres + grow * N + col);
#pragma unroll
        for (int j = 0; j < 8; ++j) dst[j] = valid ? __ldg(xp + j) : make_uint4(0, 0, 0, 0);
      };
      load_x((int)g * WGC, xr);   // (first chunk: in flight before the accumulator is even ready)
#pragma unroll 1
      for (int ci = 0; ci < NT * CH; ++ci) {
        const int tt = ci / CH, cc = ci % CH;
        if (cc == 0) {
          mbar_wait(&tmem_full[tt], it & 1u);
          tcgen05_after_thread_sync();
        }
        const int col0 = tt * BN + (int)g * WGC + cc * 64;       // first column of this chunk
        const uint32_t taddr = tmem_base + lane_sel + col0;
        float v[64];
        {
          uint32_t r0[32], r1[32];
          tmem_ld_x32(taddr, r0);
          tmem_ld_x32(taddr + 32, r1);
          if (ci > 0) load_x(col0, xr);      // in flight under the tensor-memory load, the bias add and the dropout mask
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) { v[j] = __uint_as_float(r0[j]); v[32 + j] = __uint_as_float(r1[j]); }
        }
        {
          const float4* b4 = reinterpret_cast<const float4*>(bias + col0);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float4 bb = __ldg(b4 + j);
            v[4 * j + 0] += bb.x; v[4 * j + 1] += bb.y; v[4 * j + 2] += bb.z; v[4 * j + 3] += bb.w;
          }
        }
        if (rate > 0.f) {
          unsigned long long bytes = 0ull;
#pragma unroll
          for (int k8 = 0; k8 < 8; ++k8) {
            const uint32_t keep = dropout_keep8(seed, site, (unsigned long long)(grow * (N / 8) + (col0 >> 3) + k8), rate);
            bytes |= (unsigned long long)keep << (8 * k8);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[8 * k8 + e] = ((keep >> e) & 1u) ? v[8 * k8 + e] * inv_keep : 0.f;
          }
          if (valid && keep_out != nullptr) *reinterpret_cast<unsigned long long*>(keep_out + grow * (N / 8) + (col0 >> 3)) = bytes;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t w[4] = {xr[j].x, xr[j].y, xr[j].z, xr[j].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = unpack_bf16x2(w[e]);
            v[8 * j + 2 * e] += f.x; v[8 * j + 2 * e + 1] += f.y;
          }
        }
#pragma unroll
        for (int j = 0; j < 64; ++j) { s1 += v[j]; s2 += v[j] * v[j]; }
        {   // z (fp32) back into tensor memory, over the accumulator
          uint32_t r0[32], r1[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) { r0[j] = __float_as_uint(v[j]); r1[j] = __float_as_uint(v[32 + j]); }
          tmem_st_x32(taddr, r0);
          tmem_st_x32(taddr + 32, r1);
        }
        // z -> bf16 -> staging -> TMA store (the staging buffer `buf` is free: see the matching comment in gemm.cu)
        const uint32_t stg = stg_base + buf * Cfg::D_BYTES;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          st_shared_v4(stg + swz128_off(row, j), pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                       pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
        fence_proxy_async_smem();
        if (wg_tid == 0) tma_store_wait_read<0>();
        named_bar_sync(1 + g, 128);
        if (wg_tid == 0) {
          tma_store_2d(&tmZ, stg_ptr + buf * Cfg::D_BYTES, col0, m0);
          tma_store_commit();
        }
        buf ^= 1u;
      }
      // ------------------------------ row statistics: swap partial sums with the other warpgroup ------------------------------
      float2* xb = xch + (it & 1u) * 256;
      xb[g * 128 + row] = make_float2(s1, s2);
      tmem_st_wait();
      named_bar_sync(3, 256);
      const float2 o = xb[(g ^ 1u) * 128 + row];
      const float mean = (s1 + o.x) * inv_n;
      const float var = fmaxf((s2 + o.y) * inv_n - mean * mean, 0.f);
      const float rstd = rsqrtf(var + eps);
      if (g == 0 && valid) { mean_out[grow] = mean; rstd_out[grow] = rstd; }
      // ------------------------------ pass B: normalise from tensor memory ------------------------------
      const float nmr = -mean * rstd;
#pragma unroll 1
      for (int ci = 0; ci < NT * CH; ++ci) {
        const int tt = ci / CH, cc = ci % CH;
        const int col0 = tt * BN + (int)g * WGC + cc * 64;
        const uint32_t taddr = tmem_base + lane_sel + col0;
        float v[64];
        {
          uint32_t r0[32], r1[32];
          tmem_ld_x32(taddr, r0);
          tmem_ld_x32(taddr + 32, r1);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) { v[j] = __uint_as_float(r0[j]); v[32 + j] = __uint_as_float(r1[j]); }
        }
        if (cc == CH - 1) {        // this warpgroup has read its columns of tile tt: the tensor core may overwrite them
          tcgen05_before_thread_sync();
          mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[tt]), 0));
        }
        const float4* g4 = reinterpret_cast<const float4*>(gamma + col0);
        const float4* b4 = reinterpret_cast<const float4*>(beta + col0);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float4 gg = __ldg(g4 + j), bb = __ldg(b4 + j);
          v[4 * j + 0] = fmaf(v[4 * j + 0], rstd, nmr) * gg.x + bb.x;
          v[4 * j + 1] = fmaf(v[4 * j + 1], rstd, nmr) * gg.y + bb.y;
          v[4 * j + 2] = fmaf(v[4 * j + 2], rstd, nmr) * gg.z + bb.z;
          v[4 * j + 3] = fmaf(v[4 * j + 3], rstd, nmr) * gg.w + bb.w;
        }
        const uint32_t stg = stg_base + buf * Cfg::D_BYTES;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          st_shared_v4(stg + swz128_off(row, j), pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                       pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
        fence_proxy_async_smem();
        if (wg_tid == 0) tma_store_wait_read<0>();
        named_bar_sync(1 + g, 128);
        if (wg_tid == 0) {
          tma_store_2d(&tmO, stg_ptr + buf * Cfg::D_BYTES, col0, m0);
          tma_store_commit();
        }
        buf ^= 1u;
      }
    }
    if (wg_tid == 0) tma_store_wait_all<0>();
  }

  tcgen05_before_thread_sync();
  cluster_sync_all();
  tcgen05_after_thread_sync();
  if (warp == 2) tmem_dealloc_2sm<(N <= 128 ? 128 : (N <= 256 ? 256 : 512))>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// v2 (N = 256 / 512), the default: a cluster of FOUR CTAs = two cta_group::2 pairs on the same 256 rows, each pair owning one half
// (BNC = N / 2) of the columns; every CTA keeps 128 rows x BNC accumulators DOUBLE-BUFFERED (2 x BNC <= 512 tensor-memory columns), so
// the tensor core works on the next row block while this one is normalised.  What v1 paid and v2 does not (measured, see
// profiles/r02_gemm_ln_v2.txt; tools/microbench/ldtm.cu: a tensor-memory read runs at only ~55-90 B/cycle/SM):
//   * every accumulator is read ONCE: z stays in registers as packed bf16 (32 registers for 64 columns) between the statistics pass
//     and the normalisation (v1: 128 x 512 fp32 read twice + written back = 8200 cycles per row block, next to a 4100-cycle K = 512
//     mainloop; the normalisation uses the bf16-rounded z that is also what the backward reads);
//   * BNC / 64 epilogue warpgroups of 64 columns each (16 warps for N = 512): with two warps per scheduler the epilogue ran at IPC ~0.5;
//   * the residual comes through TMA into the staging box, z is written over it in place, the normalised output over that (v1: one
//     16-byte global load per ROW per instruction, 8192 LSU wavefronts per row block); bias / gamma / beta sit in shared memory
//     (broadcast reads);
//   * the row sums of the other half of the columns arrive from CTA rank ^ 2 through distributed shared memory (st.async +
//     complete_tx on the peer's mbarrier) -- 8 bytes per row, no cluster-wide barrier, no acquire at cluster scope (that one
//     invalidated L1 on every wait).
// Tried on the way and dropped: two INDEPENDENT cta_group::1 CTAs per row block (each must take in the whole 32 KB B tile per k-block:
// 48 KB per 512 MMA cycles, more than an SM ingests; multicast across a cluster of four halves the L2 reads but not the ingest).
// ------------------------------------------------------------------------------------------------
constexpr int LN2_STAGES = 4;       // w0 TMA, w1 MMA, w2 TMEM alloc, w3 idle, then BNC / 64 epilogue warpgroups (64 columns each, thread = row)

template <int BNC>
struct Ln2Cfg {
  static constexpr uint32_t A_BYTES = LN_BM * LN_BK * 2;
  static constexpr uint32_t B_BYTES = (BNC / 2) * LN_BK * 2;         // this CTA's half of the pair's B tile
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr uint32_t BOX_BYTES = LN_BM * 128;                 // one [128 rows x 64 bf16] SWIZZLE_128B box (64-byte rows halve the TMA store rate)
  static constexpr int NWG = BNC / 64;                               // epilogue warpgroups
  static constexpr int THREADS = 128 + NWG * 128;
  static constexpr uint32_t OFF_STG = LN2_STAGES * STAGE_BYTES;      // one box per warpgroup: residual in (TMA) -> z in place -> out
  static constexpr uint32_t OFF_XL = OFF_STG + NWG * BOX_BYTES;      // row sums between the warpgroups: [2 parities][NWG][128] float2
  static constexpr uint32_t OFF_XP = OFF_XL + 2 * NWG * 128 * 8;     // row sums from the peer CTA: [2 parities][128] float2
  static constexpr uint32_t OFF_VEC = OFF_XP + 2 * 128 * 8;           // bias | gamma | beta of this CTA's BNC columns (fp32)
  static constexpr uint32_t OFF_BAR = OFF_VEC + 3 * BNC * 4;
  static constexpr uint32_t SMEM_BYTES = OFF_BAR + 256 + 1024;
};

// 8 bytes into the peer CTA's shared memory; the peer's mbarrier is credited with the bytes when they have landed
__device__ __forceinline__ void st_async_f2(uint32_t cluster_addr, float a, float b, uint32_t cluster_mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];"
               ::"r"(cluster_addr), "f"(a), "f"(b), "r"(cluster_mbar) : "memory");
}
// commit to the mbarrier at this offset in both CTAs of the cta_group::2 pair given by `mask` (cluster ranks)
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}
__device__ __forceinline__ uint4 ld_shared_u4(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}

#ifdef BX_LN2_TRACE
__device__ long long g_ln2_trace[2][20][64];
#define BX_TR() do { if (lane == 0 && blockIdx.x < 2 && tr_n < 64) g_ln2_trace[blockIdx.x][warp][tr_n++] = clock64(); } while (0)
#else
#define BX_TR() do { } while (0)
#endif

template <int BNC>
__global__ void __launch_bounds__(Ln2Cfg<BNC>::THREADS, 1)
gemm_ln2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmX,
                const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmO, int M, int K,
                const float* __restrict__ bias, const float* __restrict__ gamma, const float* __restrict__ beta,
                float* __restrict__ mean_out, float* __restrict__ rstd_out, uint8_t* __restrict__ keep_out, float eps, float rate,
                unsigned long long seed, uint32_t site, const unsigned long long* __restrict__ seed_dev) {
  using Cfg = Ln2Cfg<BNC>;
  constexpr int N = 2 * BNC;
  constexpr int NWG = Cfg::NWG;
  constexpr int WGC = 64;               // columns owned by one epilogue warpgroup (thread = row): with only two warps per scheduler the
                                        // epilogue was issue-latency bound (IPC ~0.5), so BNC = 256 runs four warpgroups
  static_assert(BNC == 128 || BNC == 256, "gemm_ln2: N must be 256 or 512");
  if (seed_dev != nullptr) seed = *seed_dev;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* empty_bar = full_bar + LN2_STAGES;
  uint64_t* tmem_full = empty_bar + LN2_STAGES;    // [2] accumulator buffers
  uint64_t* tmem_empty = tmem_full + 2;            // [2]
  uint64_t* peer_full = tmem_empty + 2;            // [2] parities of the peer's row sums
  uint64_t* x_full = peer_full + 2;                // [NWG] residual boxes
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(x_full + 4);
  float2* xl = reinterpret_cast<float2*>(smem + Cfg::OFF_XL);
  float2* xp = reinterpret_cast<float2*>(smem + Cfg::OFF_XP);

  const uint32_t warp = warp_idx();
  const uint32_t lane = lane_idx();
  // cluster of four = two cta_group::2 PAIRS (ranks {0,1} and {2,3}).  A pair works on 256 rows x BNC columns (column half ch = rank / 2)
  // exactly like the plain GEMM: each CTA stages its 128 rows of A and HALF of the B tile (32 KB per k-block -- two independent CTAs
  // would each pull the whole B tile, 48 KB per 512-cycle k-block, more than an SM takes in), and keeps its 128 rows x BNC accumulators.
  // The other pair has the other BNC columns of the same rows: row sums are swapped with CTA rank ^ 2.
  const uint32_t rank = cluster_ctarank();
  const uint32_t ch = rank >> 1, rb = rank & 1u;
  const bool leader = rb == 0u;
  const uint32_t lrank = rank & ~1u;
  const int n_base = (int)ch * BNC;
  const uint16_t pair_mask = (uint16_t)(3u << lrank);
#ifdef BX_LN2_TRACE
  int tr_n = 0;
#endif

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmB); prefetch_tmap(&tmX); prefetch_tmap(&tmZ); prefetch_tmap(&tmO);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < LN2_STAGES; ++i) { mbar_init(&full_bar[i], 2); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 2 * NWG * 128); mbar_init(&peer_full[i], 1); }
    for (int i = 0; i < 4; ++i) mbar_init(&x_full[i], 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm<2 * BNC>(tmem_ptr);
  {   // per-column vectors: every epilogue thread reads the same address at the same time -> shared-memory broadcast, not 24 global loads per sub-chunk
    float* vec = reinterpret_cast<float*>(smem + Cfg::OFF_VEC);
    for (int i = threadIdx.x; i < BNC; i += Cfg::THREADS) {
      vec[i] = bias[n_base + i]; vec[BNC + i] = gamma[n_base + i]; vec[2 * BNC + i] = beta[n_base + i];
    }
  }
  tcgen05_before_thread_sync();
  cluster_sync_all();                 // (also: the peer's mbarriers exist before anything is sent to them)
  tcgen05_after_thread_sync();
  const uint32_t tmem_base = *tmem_ptr;
  const int unit = blockIdx.x / 4, nunits = gridDim.x / 4;
  const int m_tiles = (M + 2 * LN_BM - 1) / (2 * LN_BM);     // PAIRS of row blocks (a missing second block is all out-of-bounds: zero loads, clipped stores)
  const int kblocks = (K + LN_BK - 1) / LN_BK;

  if (warp == 0 && lane == 0) {
    // ===================================== TMA producer =====================================
    uint32_t stage = 0, phase = 0;
    for (int t = unit; t < m_tiles; t += nunits) {
      const int m0 = (2 * t + (int)rb) * LN_BM;
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
        uint8_t* sb = sa + Cfg::A_BYTES;
        const int k0 = kb * LN_BK;
        const uint32_t lbar = mapa_u32(smem_u32(&full_bar[stage]), lrank);      // the pair leader's barrier collects both CTAs' bytes
        if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
        else mbar_arrive_cluster(lbar);
        tma_load_2d_2sm(sa, &tmA, lbar, k0, m0);
#pragma unroll
        for (int p = 0; p < BNC / 128; ++p) tma_load_2d_2sm(sb + p * 8192, &tmB, lbar, n_base + (int)rb * (BNC / 2) + p * 64, k0);
        if (++stage == LN2_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && leader) {
    // ===================================== MMA issuer (pair leader) =====================================
    constexpr uint32_t idesc = make_idesc_bf16(2 * LN_BM, BNC, 0, 1);
    uint32_t stage = 0, phase = 0, it = 0;
    for (int t = unit; t < m_tiles; t += nunits, ++it) {
      const uint32_t acc = it & 1u;
      BX_TR();
      mbar_wait(&tmem_empty[acc], ((it >> 1) & 1u) ^ 1u);        // both CTAs' epilogues have read this accumulator buffer
      tcgen05_after_thread_sync();
      BX_TR();
      const uint32_t d_tmem = tmem_base + acc * BNC;
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_after_thread_sync();
        if (kb == 0) BX_TR();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < LN_BK / 16; ++k) {
            const uint64_t adesc = make_smem_desc(sa + k * 32, 0, 1024, kSwz128);
            const uint64_t bdesc = make_smem_desc(sb + k * 2048, LN_BK * 128, 1024, kSwz128);
            umma_f16_2sm(d_tmem, adesc, bdesc, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit_pair(&empty_bar[stage], pair_mask);
          if (kb == kblocks - 1) umma_commit_pair(&tmem_full[acc], pair_mask);
        }
        __syncwarp();
        if (++stage == LN2_STAGES) { stage = 0; phase ^= 1; }
      }
      BX_TR();
    }
  } else if (warp >= 4) {
    // ===================================== epilogue: NWG warpgroups, thread = row =====================================
    const uint32_t g = (warp - 4) >> 2;            // warpgroup: columns n_base + [g * 64, g * 64 + 64)
    const uint32_t q = warp & 3u;                  // TMEM lane quadrant
    const uint32_t row = q * 32u + lane;
    const uint32_t wg_tid = (threadIdx.x - 128u) & 127u;
    const uint32_t lane_sel = (q * 32u) << 16;
    const int colw = n_base + (int)g * WGC;
    uint8_t* const stg_ptr = smem + Cfg::OFF_STG + g * Cfg::BOX_BYTES;
    const uint32_t stg = smem_u32(stg_ptr);
    uint64_t* const xf = x_full + g;
    const float inv_keep = 1.f / (1.f - rate);
    const float inv_n = 1.f / (float)N;
    const uint32_t vec_base = smem_u32(smem + Cfg::OFF_VEC);
    const uint32_t peer_xp = mapa_u32(smem_u32(xp), rank ^ 2u);
    const uint32_t peer_bar = mapa_u32(smem_u32(peer_full), rank ^ 2u);
    const uint32_t lead_empty = mapa_u32(smem_u32(tmem_empty), lrank);
    uint32_t it = 0;
    if (wg_tid == 0 && unit < m_tiles) {
      mbar_arrive_expect_tx(xf, Cfg::BOX_BYTES);
      tma_load_2d(stg_ptr, &tmX, xf, colw, (2 * unit + (int)rb) * LN_BM);
    }
    for (int t = unit; t < m_tiles; t += nunits, ++it) {
      const uint32_t acc = it & 1u, par = it & 1u;
      const int m0 = (2 * t + (int)rb) * LN_BM;
      const long long grow = (long long)m0 + row;
      const bool valid = grow < M;
      uint32_t zpk[WGC / 2];
      float s1 = 0.f, s2 = 0.f;
      BX_TR();                                           // 0: row block start
      mbar_wait(&tmem_full[acc], (it >> 1) & 1u);
      tcgen05_after_thread_sync();
      BX_TR();                                           // 1: accumulator ready
      // ------------------------------ pass A: z, statistics, z -> global; z stays in registers as bf16 ------------------------------
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int col0 = colw + h * 32;
        float v[32];
        {
          uint32_t r[32];
          tmem_ld_x32(tmem_base + lane_sel + acc * BNC + g * WGC + h * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        }
        BX_TR();                                         // 2, 5: tensor-memory load done
        if (h == 1) {                    // the tensor core may reuse this accumulator buffer
          tcgen05_before_thread_sync();
          mbar_arrive_cluster(lead_empty + acc * 8);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bb = ld_shared_f4(vec_base + (g * WGC + h * 32 + 4 * j) * 4);
          v[4 * j + 0] += bb.x; v[4 * j + 1] += bb.y; v[4 * j + 2] += bb.z; v[4 * j + 3] += bb.w;
        }
        if (rate > 0.f) {
          uint32_t bytes = 0u;
#pragma unroll
          for (int k8 = 0; k8 < 4; ++k8) {
            const uint32_t keep = dropout_keep8(seed, site, (unsigned long long)(grow * (N / 8) + (col0 >> 3) + k8), rate);
            bytes |= keep << (8 * k8);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[8 * k8 + e] = ((keep >> e) & 1u) ? v[8 * k8 + e] * inv_keep : 0.f;
          }
          if (valid && keep_out != nullptr) *reinterpret_cast<uint32_t*>(keep_out + grow * (N / 8) + (col0 >> 3)) = bytes;
        }
        BX_TR();                                         // 3, 6: bias + dropout done
        if (h == 0) mbar_wait(xf, it & 1u);          // this row block's residual box (TMA, requested during the previous block)
        BX_TR();                                         // 4, 7: residual there
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 xr = ld_shared_u4(stg + swz128_off(row, 4 * h + j));
          const uint32_t w[4] = {xr.x, xr.y, xr.z, xr.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = unpack_bf16x2(w[e]);
            v[8 * j + 2 * e] += f.x; v[8 * j + 2 * e + 1] += f.y;
          }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) { s1 += v[j]; s2 += v[j] * v[j]; }
#pragma unroll
        for (int j = 0; j < 16; ++j) zpk[h * 16 + j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
#pragma unroll
        for (int j = 0; j < 4; ++j)      // z over this row's residual bytes
          st_shared_v4(stg + swz128_off(row, 4 * h + j), zpk[h * 16 + 4 * j], zpk[h * 16 + 4 * j + 1], zpk[h * 16 + 4 * j + 2], zpk[h * 16 + 4 * j + 3]);
      }
      BX_TR();                                           // 8: z written
      fence_proxy_async_smem();
      named_bar_sync(1 + g, 128);
      if (wg_tid == 0) {
        tma_store_2d(&tmZ, stg_ptr, colw, m0);
        tma_store_commit();
      }
      BX_TR();                                           // 9: z store issued
      // ------------------------------ row statistics: the other warpgroups (smem), the other CTA (DSMEM) ------------------------------
      float2* xlb = xl + par * (NWG * 128);
      xlb[g * 128 + row] = make_float2(s1, s2);
      named_bar_sync(1 + NWG, NWG * 128);
      BX_TR();                                           // 10: all warpgroups have their sums
      float c1 = 0.f, c2 = 0.f;                          // this CTA's BNC columns (same summation order in every warpgroup)
#pragma unroll
      for (int i = 0; i < NWG; ++i) { const float2 o = xlb[i * 128 + row]; c1 += o.x; c2 += o.y; }
      if (g == 0) {
        if (wg_tid == 0) mbar_arrive_expect_tx(&peer_full[par], 128 * 8);
        st_async_f2(peer_xp + (par * 128 + row) * 8, c1, c2, peer_bar + par * 8);
      }
      mbar_wait(&peer_full[par], (it >> 1) & 1u);      // (complete_tx makes the st.async data visible, like a TMA load's)
      const float2 pv = xp[par * 128 + row];
      BX_TR();                                           // 11: peer sums there
      // column half 0 first on both CTAs, so that both compute bit-identical statistics
      const float t1 = ch == 0 ? c1 + pv.x : pv.x + c1, t2 = ch == 0 ? c2 + pv.y : pv.y + c2;
      const float mean = t1 * inv_n;
      const float var = fmaxf(t2 * inv_n - mean * mean, 0.f);
      const float rstd = rsqrtf(var + eps);
      if (ch == 0 && g == 0 && valid) { mean_out[grow] = mean; rstd_out[grow] = rstd; }
      // ------------------------------ pass B: normalise the bf16 z held in registers ------------------------------
      const float nmr = -mean * rstd;
      if (wg_tid == 0) tma_store_wait_read<0>();       // the z store has read the box out
      named_bar_sync(1 + g, 128);
      BX_TR();                                           // 12: box free
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t gaddr = vec_base + (BNC + g * WGC + h * 32) * 4, baddr = gaddr + BNC * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 ga = ld_shared_f4(gaddr + 32 * j), gb = ld_shared_f4(gaddr + 32 * j + 16);
          const float4 ba = ld_shared_f4(baddr + 32 * j), bb = ld_shared_f4(baddr + 32 * j + 16);
          const float2 z0 = unpack_bf16x2(zpk[h * 16 + 4 * j]), z1 = unpack_bf16x2(zpk[h * 16 + 4 * j + 1]);
          const float2 z2 = unpack_bf16x2(zpk[h * 16 + 4 * j + 2]), z3 = unpack_bf16x2(zpk[h * 16 + 4 * j + 3]);
          st_shared_v4(stg + swz128_off(row, 4 * h + j),
                       pack_bf16x2(fmaf(z0.x, rstd, nmr) * ga.x + ba.x, fmaf(z0.y, rstd, nmr) * ga.y + ba.y),
                       pack_bf16x2(fmaf(z1.x, rstd, nmr) * ga.z + ba.z, fmaf(z1.y, rstd, nmr) * ga.w + ba.w),
                       pack_bf16x2(fmaf(z2.x, rstd, nmr) * gb.x + bb.x, fmaf(z2.y, rstd, nmr) * gb.y + bb.y),
                       pack_bf16x2(fmaf(z3.x, rstd, nmr) * gb.z + bb.z, fmaf(z3.y, rstd, nmr) * gb.w + bb.w));
        }
      }
      BX_TR();                                           // 13: out written
      fence_proxy_async_smem();
      named_bar_sync(1 + g, 128);
      BX_TR();                                           // 14: out barrier
      if (wg_tid == 0) {
        tma_store_2d(&tmO, stg_ptr, colw, m0);
        tma_store_commit();
        if (t + nunits < m_tiles) {                    // the next row block's residual, as soon as the box has been read out
          tma_store_wait_read<0>();
          mbar_arrive_expect_tx(xf, Cfg::BOX_BYTES);
          tma_load_2d(stg_ptr, &tmX, xf, colw, (2 * (t + nunits) + (int)rb) * LN_BM);
        }
      }
    }
    if (wg_tid == 0) tma_store_wait_all<0>();
  }

  tcgen05_before_thread_sync();
  cluster_sync_all();                 // (no CTA exits while its peer may still send to it)
  tcgen05_after_thread_sync();
  if (warp == 2) tmem_dealloc_2sm<2 * BNC>(tmem_base);
}

template <int BNC>
static int launch_gemm_ln2(const void* A, long long lda, const void* W, long long ldw, const float* bias, const void* xres, const float* gamma,
                           const float* beta, void* z, void* out, float* mean, float* rstd, void* keep_bits, long long M, long long K,
                           float eps, float rate, unsigned long long seed, unsigned site, cudaStream_t stream,
                           const unsigned long long* seed_dev) {
  using Cfg = Ln2Cfg<BNC>;
  constexpr int N = 2 * BNC;
  CUtensorMap tmA, tmB, tmX, tmZ, tmO;
  int rc;
  if ((rc = encode_tmap_2d(&tmA, 0, A, K, M, lda, 64, LN_BM, 128))) return rc;
  if ((rc = encode_tmap_2d(&tmB, 0, W, N, K, ldw, 64, 64, 128))) return rc;
  if ((rc = encode_tmap_2d(&tmX, 0, xres, N, M, N, 64, LN_BM, 128))) return rc;
  if ((rc = encode_tmap_2d(&tmZ, 0, z, N, M, N, 64, LN_BM, 128))) return rc;
  if ((rc = encode_tmap_2d(&tmO, 0, out, N, M, N, 64, LN_BM, 128))) return rc;
  auto kern = gemm_ln2_kernel<BNC>;
  static std::mutex mu;
  static bool attr_done[kMaxDevices] = {};
  static int max_units_dev[kMaxDevices] = {};
  const int dev = current_device();
  std::lock_guard<std::mutex> lock(mu);
  if (!attr_done[dev]) {
    BX_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_done[dev] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(Cfg::THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 4; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int& max_units = max_units_dev[dev];
  if (max_units == 0) {
    max_units = sm_count() / 4;
    cfg.gridDim = dim3(max_units * 4);
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) == cudaSuccess && n > 0 && n < max_units) max_units = n;
    (void)cudaGetLastError();
    if (getenv("B200XF_LN_DEBUG")) fprintf(stderr, "gemm_ln2<%d>: %d SMs, occupancy query %d clusters -> %d units\n", BNC, sm_count(), n, max_units);
  }
  const long long m_tiles = (M + 2 * LN_BM - 1) / (2 * LN_BM);
  cfg.gridDim = dim3((unsigned)(m_tiles < max_units ? m_tiles : max_units) * 4);
  void* pb = prof_on() ? prof_begin(stream) : nullptr;
  BX_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmX, tmZ, tmO, (int)M, (int)K, bias, gamma, beta, mean, rstd, (uint8_t*)keep_bits, eps, rate,
                                   seed, (uint32_t)site, seed_dev));
  if (pb) prof_end(pb, 4, 2.0 * (double)M * (double)N * (double)K, stream);
  BX_POST_LAUNCH();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// host launcher
// ------------------------------------------------------------------------------------------------
template <int BN, int NT>
static int launch_gemm_ln(const void* A, long long lda, const void* W, long long ldw, const float* bias, const void* xres, const float* gamma,
                          const float* beta, void* z, void* out, float* mean, float* rstd, void* keep_bits, long long M, long long K,
                          float eps, float rate, unsigned long long seed, unsigned site, cudaStream_t stream,
                          const unsigned long long* seed_dev) {
  using Cfg = LnCfg<BN>;
  constexpr int N = BN * NT;
  CUtensorMap tmA, tmB, tmZ, tmO;
  int rc;
  if ((rc = encode_tmap_2d(&tmA, 0, A, K, M, lda, 64, LN_BM, 128))) return rc;
  if ((rc = encode_tmap_2d(&tmB, 0, W, N, K, ldw, 64, 64, 128))) return rc;
  if ((rc = encode_tmap_2d(&tmZ, 0, z, N, M, N, 64, LN_BM, 128))) return rc;
  if ((rc = encode_tmap_2d(&tmO, 0, out, N, M, N, 64, LN_BM, 128))) return rc;
  auto kern = gemm_ln_kernel<BN, NT>;
  static std::mutex mu;
  static bool attr_done[kMaxDevices] = {};
  static int max_units_dev[kMaxDevices] = {};
  const int dev = current_device();
  std::lock_guard<std::mutex> lock(mu);
  if (!attr_done[dev]) {
    BX_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_done[dev] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(LN_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int& max_units = max_units_dev[dev];
  if (max_units == 0) {
    max_units = sm_count() / 2;
    cfg.gridDim = dim3(max_units * 2);
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) == cudaSuccess && n > 0 && n < max_units) max_units = n;
    (void)cudaGetLastError();
  }
  const long long m_tiles = (M + 2 * LN_BM - 1) / (2 * LN_BM);
  cfg.gridDim = dim3((unsigned)(m_tiles < max_units ? m_tiles : max_units) * 2);
  void* pb = prof_on() ? prof_begin(stream) : nullptr;
  BX_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmZ, tmO, (int)M, (int)K, bias, (const __nv_bfloat16*)xres, gamma, beta, (const void*)z, mean, rstd,
                                   (uint8_t*)keep_bits, eps, rate, seed, (uint32_t)site, seed_dev));
  if (pb) prof_end(pb, 4, 2.0 * (double)M * (double)N * (double)K, stream);      // category 4: Dense + LayerNorm fused (its time contains the LayerNorm)
  BX_POST_LAUNCH();
  return 0;
}

#ifdef BX_LN2_TRACE
extern "C" int b200xf_debug_ln2_trace(long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_ln2_trace, sizeof(long long) * 2 * 20 * 64);
}
#endif

bool gemm_ln_supported(long long M, long long N, long long K) {
  return (N == 128 || N == 256 || N == 512) && K % 8 == 0 && M > 0 && M < (1ll << 31) && K > 0 && K < (1ll << 31);
}

// z = xres + dropout(A.W + bias); out = LN(z)*gamma + beta.  A [M,K] bf16 (row pitch lda), W [K,N] bf16 keras layout (row pitch
// ldw), xres / z / out bf16 [M,N] contiguous, keep_bits uint8 [M, N/8] (written when rate > 0 and non-NULL).
int gemm_ln(const void* A, long long lda, const void* W, long long ldw, const float* bias, const void* xres, const float* gamma,
            const float* beta, void* z, void* out, float* mean, float* rstd, void* keep_bits, long long M, long long N, long long K, float eps,
            float rate, unsigned long long seed, unsigned site, cudaStream_t stream, const unsigned long long* seed_dev) {
  BX_REQUIRE(gemm_ln_supported(M, N, K), "gemm_ln: unsupported shape M %lld N %lld K %lld (N must be 128, 256 or 512)", M, N, K);
  BX_REQUIRE(bias && xres && gamma && beta && z && out && mean && rstd, "gemm_ln: null buffer");
  BX_REQUIRE(((reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta) |
               reinterpret_cast<uintptr_t>(xres) | reinterpret_cast<uintptr_t>(keep_bits)) & 15) == 0,
             "gemm_ln: bias / gamma / beta / residual / keep bits must be 16-byte aligned");
  BX_REQUIRE(rate >= 0.f && rate < 1.f, "gemm_ln: dropout rate out of range");
#define BX_LN_GO(bn, nt) \
  return launch_gemm_ln<bn, nt>(A, lda, W, ldw, bias, xres, gamma, beta, z, out, mean, rstd, keep_bits, M, K, eps, rate, seed, site, stream, seed_dev)
  static const bool use_v1 = [] { const char* e = getenv("B200XF_LN_V1"); return e && e[0] == '1'; }();
  if (!use_v1 && N == 512)
    return launch_gemm_ln2<256>(A, lda, W, ldw, bias, xres, gamma, beta, z, out, mean, rstd, keep_bits, M, K, eps, rate, seed, site, stream, seed_dev);
  if (!use_v1 && N == 256)
    return launch_gemm_ln2<128>(A, lda, W, ldw, bias, xres, gamma, beta, z, out, mean, rstd, keep_bits, M, K, eps, rate, seed, site, stream, seed_dev);
  if (N == 512) BX_LN_GO(256, 2);
  if (N == 256) BX_LN_GO(256, 1);
  BX_LN_GO(128, 1);
#undef BX_LN_GO
}

}  // namespace bx
