// Persistent warp-specialised tcgen05 GEMM for sm_100a:   D[M,N] (+)= epilogue( A[M,K] . B[K,N] )
//
//   bf16 operands staged HBM -> smem by TMA (SWIZZLE_128B, mbarrier full/empty ring), tcgen05.mma kind::f16
//   issued by one thread (UMMA 128 x BN x 16), fp32 accumulators double-buffered in TMEM, epilogue warps
//   read them back with tcgen05.ld, apply bias / ReLU / ReLU-mask, and TMA-store (or TMA reduce-add for
//   split-K / residual accumulation) through a swizzled smem staging tile.
//
// This one kernel serves every dense contraction on the hot path (reference call sites):
//   forward   y = x.W + b          Attention.py:62-64,77  point_ffn.py:5-6  Transformer.py:30   (A K-major, B MN-major)
//   dgrad     dx = dy.W^T          reverse-mode of the above (train.py:137)                     (A K-major, B K-major)
//   wgrad     dW = x^T.dy          ditto, split-K over tokens with TMA reduce-add               (A MN-major, B MN-major)
// Operand "major": 0 = K-major (K contiguous in memory), 1 = MN-major (M resp. N contiguous).
#include <mutex>

#include "common.cuh"

namespace bx {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 256;  // w0 TMA, w1 MMA, w2 TMEM alloc, w3 idle, w4-7 epilogue

constexpr int GEMM_FLAG_RELU = 1;       // out = max(out, 0)
constexpr int GEMM_FLAG_ACCUM = 2;      // D += out (TMA reduce-add; fp32 D only)
constexpr int GEMM_FLAG_AUX_MASK = 4;   // out *= (aux[m,n] > 0)   (ReLU backward; aux bf16)
constexpr int GEMM_FLAG_BITS_OUT = 8;   // with RELU, bf16 out: also store a 1-bit-per-element (out > 0) mask, 64 columns per uint64
constexpr int GEMM_FLAG_BITS_MASK = 16; // out *= mask bit (ReLU backward from the compact mask; aux = uint64 words, ld_aux in words)

// CG = CTAs cooperating on one tile: 1 (128 x BN tile) or 2 (a CTA pair, cta_group::2: 256 x BN tile; each CTA stages
// its own 128 rows of A and its own BN/2 columns of B, so L2->smem traffic per flop drops by a third)
template <int BN, int CG>
struct GemmCfg {
  static constexpr uint32_t A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr uint32_t B_BYTES = (BN / CG) * GEMM_BK * 2;
  static constexpr int kStages = (A_BYTES + B_BYTES) == 49152 ? 4 : ((A_BYTES + B_BYTES) == 32768 ? 6 : 8);
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr uint32_t D_BYTES = GEMM_BM * 128;  // one 128-byte-wide staging box
  static constexpr uint32_t OFF_D = kStages * STAGE_BYTES;
  static constexpr uint32_t OFF_BAR = OFF_D + 2 * D_BYTES;
  static constexpr uint32_t SMEM_BYTES = OFF_BAR + 256 + 1024;  // + alignment slack
  static constexpr uint32_t TMEM_COLS = 2 * BN;
};

template <int MA, int MB, int BN, bool kOutF32, int CG>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmD, int M, int N, int K, int splits,
            const float* __restrict__ bias, void* aux_v, long long ld_aux, int flags, int chunk, float* __restrict__ colsum) {
  const __nv_bfloat16* aux = reinterpret_cast<const __nv_bfloat16*>(aux_v);
  using Cfg = GemmCfg<BN, CG>;
  constexpr int kStages = Cfg::kStages;
  constexpr int OUT_COLS = kOutF32 ? 32 : 64;  // columns per 128-byte staging row
  constexpr int BNC = BN / CG;                 // B columns staged by this CTA

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const uint32_t warp = warp_idx();
  const uint32_t lane = lane_idx();
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;   // rank 0 of a pair issues the MMAs
  const bool leader = rank == 0u;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmD);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], CG);       // pair: the leader's barrier collects one arrive per CTA (+ both CTAs' bytes)
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 128 * CG);   // pair: the leader's barrier collects both CTAs' epilogue threads
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (CG == 2) tmem_alloc_2sm<Cfg::TMEM_COLS>(tmem_ptr);
    else tmem_alloc<Cfg::TMEM_COLS>(tmem_ptr);
  }
  tcgen05_before_thread_sync();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tcgen05_after_thread_sync();
  const uint32_t tmem_base = *tmem_ptr;
  // Tile schedule of a unit (a CTA, or a CTA pair).  chunk == 0: PERSISTENT -- one unit per SM (pair), tiles unit, unit + units, ...
  // chunk > 0: unit u works on the `chunk` consecutive tiles from u * chunk and there are more units than SMs, so the hardware
  // block scheduler hands chunks to whichever SMs are free: used while NCCL collectives share the GPU (data-parallel backward),
  // where a persistent grid's CTAs that find their SM occupied would start a whole wave late.
  const int unit_id = blockIdx.x / CG;
  const int t_begin = chunk > 0 ? unit_id * chunk : unit_id;
  const int t_step = chunk > 0 ? 1 : (int)(gridDim.x / CG);

  const int m_tiles = (M + GEMM_BM * CG - 1) / (GEMM_BM * CG);
  const int n_tiles = (N + BN - 1) / BN;
  const int kblocks = (K + GEMM_BK - 1) / GEMM_BK;
  const int tiles_mn = m_tiles * n_tiles;
  const int total = tiles_mn * splits;
  const int t_end = chunk > 0 ? min(total, t_begin + chunk) : total;

  if (warp == 0 && lane == 0) {
    // ===================================== TMA producer =====================================
    uint32_t stage = 0, phase = 0;
    for (int t = t_begin; t < t_end; t += t_step) {
      const int split = t / tiles_mn, rem = t % tiles_mn;
      const int m0 = (rem / n_tiles) * (GEMM_BM * CG) + (int)rank * GEMM_BM;   // this CTA's 128 rows of A
      const int n0 = (rem % n_tiles) * BN + (int)rank * BNC;                     // this CTA's BN/CG columns of B
      const int kb0 = (int)((long long)split * kblocks / splits), kb1 = (int)((long long)(split + 1) * kblocks / splits);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
        uint8_t* sb = sa + Cfg::A_BYTES;
        const int k0 = kb * GEMM_BK;
        if (CG == 1) {
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          if (MA == 0) {
            tma_load_2d(sa, &tmA, &full_bar[stage], k0, m0);
          } else {
#pragma unroll
            for (int p = 0; p < GEMM_BM / 64; ++p) tma_load_2d(sa + p * 8192, &tmA, &full_bar[stage], m0 + p * 64, k0);
          }
          if (MB == 0) {
            tma_load_2d(sb, &tmB, &full_bar[stage], k0, n0);
          } else {
#pragma unroll
            for (int p = 0; p < BNC / 64; ++p) tma_load_2d(sb + p * 8192, &tmB, &full_bar[stage], n0 + p * 64, k0);
          }
        } else {
          // both CTAs load into their OWN smem; all bytes are credited to the LEADER's full barrier
          const uint32_t lbar = mapa_u32(smem_u32(&full_bar[stage]), 0);
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
          else mbar_arrive_cluster(lbar);
          if (MA == 0) {
            tma_load_2d_2sm(sa, &tmA, lbar, k0, m0);
          } else {
#pragma unroll
            for (int p = 0; p < GEMM_BM / 64; ++p) tma_load_2d_2sm(sa + p * 8192, &tmA, lbar, m0 + p * 64, k0);
          }
          if (MB == 0) {
            tma_load_2d_2sm(sb, &tmB, lbar, k0, n0);
          } else {
#pragma unroll
            for (int p = 0; p < BNC / 64; ++p) tma_load_2d_2sm(sb + p * 8192, &tmB, lbar, n0 + p * 64, k0);
          }
        }
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && leader) {
    // ===================================== MMA issuer (pair: leader CTA only) =====================================
    constexpr uint32_t idesc = make_idesc_bf16(GEMM_BM * CG, BN, MA, MB);
    uint32_t stage = 0, phase = 0, it = 0;
    for (int t = t_begin; t < t_end; t += t_step, ++it) {
      const int split = t / tiles_mn;
      const int kb0 = (int)((long long)split * kblocks / splits), kb1 = (int)((long long)(split + 1) * kblocks / splits);
      const uint32_t acc = it & 1u, acc_phase = (it >> 1) & 1u;
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tcgen05_after_thread_sync();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_after_thread_sync();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            const uint64_t adesc = MA == 0 ? make_smem_desc(sa + k * 32, 0, 1024, kSwz128)
                                           : make_smem_desc(sa + k * 2048, GEMM_BK * 128, 1024, kSwz128);
            const uint64_t bdesc = MB == 0 ? make_smem_desc(sb + k * 32, 0, 1024, kSwz128)
                                           : make_smem_desc(sb + k * 2048, GEMM_BK * 128, 1024, kSwz128);
            if (CG == 2) umma_f16_2sm(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            else umma_f16(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          // frees the smem slot once these MMAs retire; on the last k-block the accumulator is ready for the epilogue
          if (CG == 2) {
            umma_commit_2sm(&empty_bar[stage]);              // multicast: same barrier offset in both CTAs
            if (kb == kb1 - 1) umma_commit_2sm(&tmem_full[acc]);
          } else {
            umma_commit(&empty_bar[stage]);
            if (kb == kb1 - 1) umma_commit(&tmem_full[acc]);
          }
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================================== epilogue =====================================
    const uint32_t q = warp & 3u;                // TMEM lane quadrant this warp may access
    const uint32_t row = q * 32u + lane;         // accumulator row owned by this thread
    const uint32_t epi_tid = threadIdx.x - 128u;
    const uint32_t stg_base = smem_u32(smem + Cfg::OFF_D);
    uint32_t it = 0, buf = 0;
    for (int t = t_begin; t < t_end; t += t_step, ++it) {
      const int rem = t % tiles_mn;
      const int m0 = (rem / n_tiles) * (GEMM_BM * CG) + (int)rank * GEMM_BM, n0 = (rem % n_tiles) * BN;
      const uint32_t acc = it & 1u, acc_phase = (it >> 1) & 1u;
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_after_thread_sync();
      const uint32_t t_row = tmem_base + ((q * 32u) << 16) + acc * BN;
      const long long grow = (long long)m0 + row;
#pragma unroll 1
      for (int c = 0; c < BN / OUT_COLS; ++c) {
        const int nc = n0 + c * OUT_COLS;
        if (nc >= N) break;
        float v[OUT_COLS];
        {
          uint32_t r[OUT_COLS / 32][32];
#pragma unroll
          for (int h = 0; h < OUT_COLS / 32; ++h) tmem_ld_x32(t_row + c * OUT_COLS + h * 32, r[h]);
          tmem_ld_wait();
#pragma unroll
          for (int h = 0; h < OUT_COLS / 32; ++h)
#pragma unroll
            for (int j = 0; j < 32; ++j) v[h * 32 + j] = __uint_as_float(r[h][j]);
        }
        if (bias != nullptr) {
          if (nc + OUT_COLS <= N) {
            const float4* b4 = reinterpret_cast<const float4*>(bias + nc);
#pragma unroll
            for (int j = 0; j < OUT_COLS / 4; ++j) {
              const float4 b = __ldg(b4 + j);
              v[4 * j + 0] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < OUT_COLS; ++j)
              if (nc + j < N) v[j] += __ldg(bias + nc + j);
          }
        }
        if (flags & GEMM_FLAG_RELU) {
#pragma unroll
          for (int j = 0; j < OUT_COLS; ++j) v[j] = fmaxf(v[j], 0.f);
          if (!kOutF32 && (flags & GEMM_FLAG_BITS_OUT) && grow < M) {   // compact ReLU gate for the backward pass
            unsigned long long bits = 0ull;
#pragma unroll
            for (int j = 0; j < OUT_COLS; ++j) bits |= (unsigned long long)(v[j] > 0.f) << j;
            reinterpret_cast<unsigned long long*>(aux_v)[grow * ld_aux + (nc >> 6)] = bits;
          }
        }
        if (!kOutF32 && (flags & GEMM_FLAG_BITS_MASK)) {
          const unsigned long long bits = grow < M ? __ldg(reinterpret_cast<const unsigned long long*>(aux_v) + grow * ld_aux + (nc >> 6)) : 0ull;
#pragma unroll
          for (int j = 0; j < OUT_COLS; ++j)
            if (!((bits >> j) & 1ull)) v[j] = 0.f;
        }
        if ((flags & GEMM_FLAG_AUX_MASK) && grow < M) {
          const __nv_bfloat16* ap = aux + grow * ld_aux + nc;
          if (nc + OUT_COLS <= N) {
#pragma unroll
            for (int j = 0; j < OUT_COLS / 8; ++j) {
              const uint4 a = __ldg(reinterpret_cast<const uint4*>(ap) + j);
              const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f = unpack_bf16x2(w[e]);
                if (!(f.x > 0.f)) v[8 * j + 2 * e] = 0.f;
                if (!(f.y > 0.f)) v[8 * j + 2 * e + 1] = 0.f;
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < OUT_COLS; ++j)
              if (nc + j < N && !(__bfloat162float(ap[j]) > 0.f)) v[j] = 0.f;
          }
        }
        // staging buffer `buf` is free: the barrier of the previous chunk was passed only after thread 0 had seen
        // the TMA store that last read it complete (two buffers, ONE named barrier per chunk).
        const uint32_t stg = stg_base + buf * Cfg::D_BYTES;
        if (kOutF32) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            st_shared_v4(stg + swz128_off(row, j), __float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]),
                         __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            st_shared_v4(stg + swz128_off(row, j), pack_bf16x2(v[8 * j], v[8 * j + 1]),
                         pack_bf16x2(v[8 * j + 2], v[8 * j + 3]), pack_bf16x2(v[8 * j + 4], v[8 * j + 5]),
                         pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
        }
        fence_proxy_async_smem();
        if (epi_tid == 0) tma_store_wait_read<0>();   // the previous chunk's store has drained the OTHER buffer
        named_bar_sync(1, 128);
        if (epi_tid == 0) {
          if (flags & GEMM_FLAG_ACCUM) tma_reduce_add_2d(&tmD, smem + Cfg::OFF_D + buf * Cfg::D_BYTES, nc, m0);
          else tma_store_2d(&tmD, smem + Cfg::OFF_D + buf * Cfg::D_BYTES, nc, m0);
          tma_store_commit();
        }
        if (!kOutF32 && colsum != nullptr) {
          // colsum[n] += sum over rows of the bf16 tile just staged (the bias gradient of the layer whose dy this GEMM produces):
          // warp q sums its 32 rows, lane = one bf16 pair of the 128-byte row (conflict-free), one fp32 atomic per column and warp.
          // The box is next written two chunks from now, behind the next chunk's barrier.
          const int rows_here = M - m0;
          float a0 = 0.f, a1 = 0.f;
#pragma unroll 8
          for (uint32_t rr = 0; rr < 32u; ++rr) {
            const uint32_t r2 = q * 32u + rr;
            if ((int)r2 < rows_here) {
              uint32_t wv;
              asm volatile("ld.shared.b32 %0, [%1];" : "=r"(wv) : "r"(stg + swz128_off(r2, lane >> 2) + (lane & 3u) * 4u));
              const float2 f = unpack_bf16x2(wv);
              a0 += f.x; a1 += f.y;
            }
          }
          const int col = nc + 2 * (int)lane;
          if (col < N) atomicAdd(colsum + col, a0);
          if (col + 1 < N) atomicAdd(colsum + col + 1, a1);
        }
        buf ^= 1u;
      }
      tcgen05_before_thread_sync();
      if (CG == 2) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[acc]), 0));   // the leader's barrier
      else mbar_arrive(&tmem_empty[acc]);
    }
    if (epi_tid == 0) tma_store_wait_all<0>();
  }

  tcgen05_before_thread_sync();
  if (CG == 2) cluster_sync_all(); else __syncthreads();   // pair: the peer's smem/TMEM stay alive until both are done
  tcgen05_after_thread_sync();
  if (warp == 2) {
    if (CG == 2) tmem_dealloc_2sm<Cfg::TMEM_COLS>(tmem_base);
    else tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// host launcher
// ------------------------------------------------------------------------------------------------
static int g_chunks_per_unit = 0;   // 0 = persistent static schedule; k > 0: ~k chunks of consecutive tiles per SM (pair), handed out by the hardware

// split-K factor for a reduce-add GEMM: aim at ~2 full waves of tile units
static int auto_splits(long long tiles, long long kblocks, int groups) {
  if (tiles >= 2 * groups) return 1;
  long long s = (2 * groups + tiles / 2) / tiles;
  if (s > kblocks / 4) s = kblocks / 4;
  return (int)(s < 1 ? 1 : s);
}

template <int MA, int MB, int BN, bool kOutF32, int CG>
static int launch_gemm(const void* A, long long lda, const void* B, long long ldb, void* D, long long ldd,
                       const float* bias, const void* aux, long long ld_aux, int flags, int splits, long long M,
                       long long N, long long K, cudaStream_t stream, float* colsum) {
  using Cfg = GemmCfg<BN, CG>;
  CUtensorMap tmA, tmB, tmD;
  int rc;
  if (MA == 0) rc = encode_tmap_2d(&tmA, 0, A, K, M, lda, 64, GEMM_BM, 128);
  else rc = encode_tmap_2d(&tmA, 0, A, M, K, lda, 64, 64, 128);
  if (rc) return rc;
  if (MB == 0) rc = encode_tmap_2d(&tmB, 0, B, K, N, ldb, 64, BN / CG, 128);
  else rc = encode_tmap_2d(&tmB, 0, B, N, K, ldb, 64, 64, 128);
  if (rc) return rc;
  rc = encode_tmap_2d(&tmD, kOutF32 ? 1 : 0, D, N, M, ldd, kOutF32 ? 32 : 64, GEMM_BM, 128);
  if (rc) return rc;
  auto kern = gemm_kernel<MA, MB, BN, kOutF32, CG>;
  // per-device launch state of this instantiation (function attributes and cluster occupancy are per device), under a lock:
  // several host threads / devices may drive the library
  static std::mutex mu;
  static bool attr_done[kMaxDevices] = {};
  static int max_units_dev[kMaxDevices] = {};
  const int dev = current_device();
  std::lock_guard<std::mutex> lock(mu);
  if (!attr_done[dev]) {
    BX_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_done[dev] = true;
  }
  const long long m_tiles = (M + GEMM_BM * CG - 1) / (GEMM_BM * CG), n_tiles = (N + BN - 1) / BN;
  const long long kblocks = (K + GEMM_BK - 1) / GEMM_BK;
  const int groups = sm_count() / CG;
  const int splits0 = splits;
  if (splits <= 0) splits = (flags & GEMM_FLAG_ACCUM) ? auto_splits(m_tiles * n_tiles, kblocks, groups) : 1;
  if (splits > kblocks) splits = (int)kblocks;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // The tile schedule is static (tile = unit + i * units), so every unit must be co-resident.  Not every TPC can
  // host a 2-CTA cluster (SMs are yield-harvested individually): ask the runtime how many clusters fit at once.
  int& max_units = max_units_dev[dev];
  if (max_units == 0) {
    max_units = groups;
    if (CG == 2) {
      cfg.gridDim = dim3(groups * CG);
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) == cudaSuccess && n > 0 && n < max_units) max_units = n;
      (void)cudaGetLastError();
    }
  }
  if (splits0 <= 0 && (flags & GEMM_FLAG_ACCUM)) {   // re-balance split-K against the real number of co-resident units
    splits = auto_splits(m_tiles * n_tiles, kblocks, max_units);
    if (splits > kblocks) splits = (int)kblocks;
  }
  const long long total2 = m_tiles * n_tiles * splits;
  int grid = (int)(total2 < max_units ? total2 : max_units) * CG;
  int chunk = 0;
  if (g_chunks_per_unit > 0 && total2 > max_units) {        // more, shorter-lived units than SMs (see the kernel's schedule comment)
    chunk = (int)((total2 + (long long)g_chunks_per_unit * max_units - 1) / ((long long)g_chunks_per_unit * max_units));
    if (chunk < 2) chunk = 2;                                  // keep the accumulator double-buffering worth its prologue
    grid = (int)((total2 + chunk - 1) / chunk) * CG;
  }
  cfg.gridDim = dim3(grid);
  void* pb = prof_on() ? prof_begin(stream) : nullptr;
  BX_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmD, (int)M, (int)N, (int)K, splits, bias, const_cast<void*>(aux), ld_aux,
                                   flags, chunk, colsum));
  if (pb) prof_end(pb, prof_gemm_cat(), 2.0 * (double)M * (double)N * (double)K, stream);
  BX_POST_LAUNCH();
  return 0;
}

static int g_force_cg = 0;   // 0 = heuristic, 1 / 2 = force (tests and benchmarks)
void gemm_force_cta_group(int cg) { g_force_cg = cg; }
void gemm_set_chunks_per_unit(int n) { g_chunks_per_unit = n < 0 ? 0 : n; }

int gemm_bf16(const void* A, long long lda, int a_major, const void* B, long long ldb, int b_major, void* D,
              long long ldd, int d_is_f32, const float* bias, const void* aux, long long ld_aux, int flags,
              int splits, long long M, long long N, long long K, cudaStream_t stream, float* colsum) {
  BX_REQUIRE(colsum == nullptr || (!d_is_f32 && splits <= 1), "gemm: column sums come from a bf16 output tile, no split-K");
  BX_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem %lld x %lld x %lld", M, N, K);
  BX_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "gemm: dimension too large");
  BX_REQUIRE(splits <= 1 || ((flags & GEMM_FLAG_ACCUM) && d_is_f32), "gemm: split-K needs fp32 D with ACCUM");
  BX_REQUIRE(!(flags & GEMM_FLAG_ACCUM) || d_is_f32, "gemm: ACCUM needs fp32 D");
  BX_REQUIRE(!(flags & GEMM_FLAG_AUX_MASK) || aux != nullptr, "gemm: AUX_MASK without aux");
  BX_REQUIRE(!(flags & GEMM_FLAG_AUX_MASK) || (ld_aux % 8 == 0 && (reinterpret_cast<uintptr_t>(aux) & 15) == 0),
             "gemm: aux must be 16-byte aligned with ld %% 8 == 0");
  BX_REQUIRE(bias == nullptr || (reinterpret_cast<uintptr_t>(bias) & 15) == 0, "gemm: bias must be 16-byte aligned");
  BX_REQUIRE(!(flags & (GEMM_FLAG_BITS_OUT | GEMM_FLAG_BITS_MASK)) ||
                 (aux != nullptr && !d_is_f32 && N % 64 == 0 && (reinterpret_cast<uintptr_t>(aux) & 7) == 0),
             "gemm: bit-mask epilogues need bf16 D, N %% 64 == 0 and an 8-byte aligned mask buffer");
  BX_REQUIRE(a_major >= 0 && a_major <= 1 && b_major >= 0 && b_major <= 1, "gemm: bad majors %d/%d", a_major, b_major);
  const bool wide = N > 128;
  const bool pair = g_force_cg ? g_force_cg == 2 : (M > 128 && wide);   // CTA pairs need >= 2 row tiles; 256-wide N tile only
#define BX_GEMM_GO(ma, mb, bn, f32, cg) \
  return launch_gemm<ma, mb, bn, f32, cg>(A, lda, B, ldb, D, ldd, bias, aux, ld_aux, flags, splits, M, N, K, stream, colsum)
#define BX_GEMM_CASE(ma, mb)                                              \
  if (a_major == ma && b_major == mb) {                                   \
    if (wide && pair) { if (d_is_f32) BX_GEMM_GO(ma, mb, 256, true, 2); else BX_GEMM_GO(ma, mb, 256, false, 2); } \
    if (wide) { if (d_is_f32) BX_GEMM_GO(ma, mb, 256, true, 1); else BX_GEMM_GO(ma, mb, 256, false, 1); }         \
    if (d_is_f32) BX_GEMM_GO(ma, mb, 128, true, 1); else BX_GEMM_GO(ma, mb, 128, false, 1);                       \
  }
  BX_GEMM_CASE(0, 1)
  BX_GEMM_CASE(0, 0)
  BX_GEMM_CASE(1, 1)
  BX_GEMM_CASE(1, 0)
#undef BX_GEMM_CASE
#undef BX_GEMM_GO
  return 1;
}

}  // namespace bx
