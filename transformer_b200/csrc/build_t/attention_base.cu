// K2: fused flash-style scaled-dot-product attention for sm_100a, forward and backward.
//   reference: scaled_dot_product_attention (Attention.py:3-34) inside MultiHeadAttention.call (Attention.py:59-78)
//
//   logits = q.k^T / sqrt(dk) + mask * (-1e9) ; P = softmax(logits) ; out = P.v
//
// Semantics kept from the reference: masking ADDS -1e9 (finite), so in fp32 every masked
// logit equals -1e9 exactly and a fully-masked row is a UNIFORM distribution over all Sk keys (never NaN).
// The canonical masks of positionalencoding.py:25-52 are derived in-kernel from compact key-padding bytes
// (1 = key token id 0) and a causal flag; an arbitrary float mask (the six-argument API) is also accepted.
//
// Layout: Q/K/V are column blocks [h*dk, (h+1)*dk) of token-major bf16 matrices [B*S, ld] (the packed output of
// the QKV projection GEMM) -- split_heads/transpose (Attention.py:52-57,74-76) never materialise.  TMA 3-D maps
// {column, token, batch} zero-fill ragged tails.  S = Q.K^T and O = P.V run on tcgen05 (accumulators in TMEM);
// softmax warps own one query row per thread (tcgen05.ld 32x32b).
#include <mutex>

#include "common.cuh"

#ifndef BX_FWD_SINGLE_PASS
#define BX_FWD_SINGLE_PASS 0
#endif

namespace bx {

// optional debug timeline: clock64 stamps written by the LAST CTA of the grid (last wave: warm instruction cache; with a
// look-ahead mask the forward CTA with the most tiles) when a buffer has been registered
__device__ long long* g_timeline = nullptr;
// the pointer is fetched ONCE per thread at kernel entry (tl_begin): a stamp site then costs a register test, not a global
// load inside the hot loops
__device__ __forceinline__ long long* tl_begin() {
  long long* t = g_timeline;
  if (t == nullptr || blockIdx.y != gridDim.y - 1 || blockIdx.z != gridDim.z - 1) return nullptr;
  const long long sel = t[255];                           // which blockIdx.x to trace; < 0: the last one
  return (long long)blockIdx.x == (sel < 0 ? (long long)gridDim.x - 1 : sel) ? t : nullptr;
}
#define tl_stamp(slot) do { if (tlp != nullptr) tlp[slot] = clock64(); } while (0)
int set_timeline(long long* dev_ptr) {
  BX_CHECK_CUDA(cudaMemcpyToSymbol(g_timeline, &dev_ptr, sizeof(dev_ptr)));
  return 0;
}

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kMaskVal = -1e9f;

struct AttnMask {
  const uint8_t* kpad;    // [B, Sk] 1 = masked key (token id 0); may be null
  const float* maskf;     // optional float mask, value m adds m * -1e9; element (b, q, k) at b*sb + q*sq + k
  long long sb, sq;
  int causal;             // 1: key index > query index is masked (look-ahead mask)
};

// A [128 x DK] bf16 tile lives in shared memory as NP = DK / PW PANELS of [128 rows x PW columns] (PW = min(DK, 64): a panel
// row is at most one 128-byte swizzle span).  Head depth 128 = two 64-column panels, 16 KB apart.
template <int DK>
struct AttnCfg {
  static_assert(DK == 32 || DK == 64 || DK == 128, "head depth");
  static constexpr int PW = DK < 64 ? DK : 64;                 // panel width (elements)
  static constexpr int NP = DK / PW;                           // panels per tile
  static constexpr uint32_t ROW_BYTES = PW * 2;                // bytes of one panel row
  static constexpr uint32_t PANEL_BYTES = 128 * ROW_BYTES;
  static constexpr uint32_t TILE_BYTES = NP * PANEL_BYTES;     // a [128 x DK] bf16 tile
  static constexpr uint32_t SWZ = PW == 64 ? kSwz128 : kSwz64;  // swizzle span == panel row bytes
  static constexpr uint32_t SBO = 8 * ROW_BYTES;               // 8-row core-matrix group
  static constexpr uint32_t P_BYTES = 128 * 256;               // [128 x 128] bf16 as two 128-byte-wide panels
  static constexpr uint32_t KSTEP_MN = 16 * ROW_BYTES;         // k16 step of an MN-major [rows x DK] tile (16 rows down)
  // byte offset of the k-th 16-element slice along DK of a K-major tile (rows x DK): panel, then 32 bytes per slice
  static __device__ __forceinline__ uint32_t kmajor_k16(int k) { return (uint32_t)((k * 16) / PW) * PANEL_BYTES + (uint32_t)((k * 16) % PW) * 2u; }
};

// byte offset of 16-byte chunk `chunk` of row `row` of a [128 x DK] bf16 tile
template <int DK>
__device__ __forceinline__ uint32_t swz_off(uint32_t row, uint32_t chunk) {
  if (DK == 32) return swz64_off(row, chunk);
  return (chunk >> 3) * AttnCfg<DK>::PANEL_BYTES + swz128_off(row, chunk & 7u);
}
// TMA load / store of a whole [128 x DK] tile (NP boxes of PW columns); the loads complete on `bar`
template <int DK>
__device__ __forceinline__ void tile_load(uint8_t* dst, const CUtensorMap* tm, uint64_t* bar, int col, int row, int b) {
#pragma unroll
  for (int p = 0; p < AttnCfg<DK>::NP; ++p) tma_load_3d(dst + p * AttnCfg<DK>::PANEL_BYTES, tm, bar, col + p * AttnCfg<DK>::PW, row, b);
}
template <int DK>
__device__ __forceinline__ void tile_store(const CUtensorMap* tm, const uint8_t* src, int col, int row, int b) {
#pragma unroll
  for (int p = 0; p < AttnCfg<DK>::NP; ++p) tma_store_3d(tm, src + p * AttnCfg<DK>::PANEL_BYTES, col + p * AttnCfg<DK>::PW, row, b);
}

// ================================================================================================
// forward:  one CTA per (128-query tile, head, batch); TWO CTAs are resident per SM (head depth <= 64).
//   A CTA is a serial chain per kv tile (S ready -> TMEM load -> mask/max -> exponentials -> P to smem -> P.V); with one CTA
//   per SM and two query tiles per CTA the two chains ran in lockstep and the MUFU unit (16 ex2 / clock / SM, the hard floor
//   of this kernel at head depth 64) idled through every load/max/store phase, and nothing overlapped the prologue (TMEM
//   allocation, barrier init, Q load: ~3000 cycles) and the epilogue (O normalise + store: ~2400 cycles) of a 4-tile CTA
//   (profiles/r02_g_attention_timeline.txt).  Two independent CTAs per SM interleave all of that by themselves.
//   Budget per CTA: 192 threads x 168 registers, 96 KB shared memory, 256 tensor-memory columns (S 128 | O DK).
// ================================================================================================
template <int DK>
struct FwdSmem {
  using C = AttnCfg<DK>;
  static constexpr int KS = DK <= 64 ? 2 : 1;                           // K stages (head depth 128: one, to stay at 128 KB)
  static constexpr uint32_t OFF_Q = 0;                                  // 1 tile (re-used as the O_lo staging at the end)
  static constexpr uint32_t OFF_K = OFF_Q + C::TILE_BYTES;              // KS stages
  static constexpr uint32_t OFF_V = OFF_K + KS * C::TILE_BYTES;         // 1 stage (V(j+1) is only needed a whole softmax later)
  static constexpr uint32_t OFF_P = OFF_V + C::TILE_BYTES;              // P tile (re-used as the O staging at the end)
  static constexpr uint32_t OFF_BAR = OFF_P + C::P_BYTES;
  static constexpr uint32_t OFF_PADW = OFF_BAR + 128;                   // key-padding bit words: [kMaxPadTiles][4]
  static constexpr uint32_t BYTES = OFF_PADW + 64 * 4 * 4 + 1024;
};
constexpr int kMaxPadTiles = 64;       // kv tiles whose pad bits are preloaded (Sk <= 8192); longer rows load per tile

// w0 TMA, w1 MMA (+TMEM alloc), w2-5 softmax (TMEM lane quadrant = warp & 3): a thread keeps a whole 128-column S row in
// registers (single pass over TMEM).
constexpr int ATTN_FWD_THREADS = 192;

// kMaskF: an arbitrary float mask is given (its own instantiation: one more global load per score; every chunk "touched")
template <int DK, bool kMaskF>
__global__ void __launch_bounds__(ATTN_FWD_THREADS, DK <= 64 ? 2 : 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                const __grid_constant__ CUtensorMap tmOlo, int has_lo,
                float* __restrict__ lse_out, AttnMask mask, int H, int Sq, int Sk, float scale) {
  using C = AttnCfg<DK>;
  using L = FwdSmem<DK>;
  constexpr int KS = L::KS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* q_full = bars;          // [1]
  uint64_t* k_full = bars + 1;      // [2]
  uint64_t* k_empty = bars + 3;     // [2]
  uint64_t* v_full = bars + 5;      // [1]
  uint64_t* v_empty = bars + 6;     // [1]
  uint64_t* s_full = bars + 7;      // [1]
  uint64_t* p_full = bars + 8;      // [1] count 128
  uint64_t* o_full = bars + 9;      // [1]
  uint64_t* s_free = bars + 10;     // [1] count 128: the softmax warps hold the S tile in registers, TMEM S may be overwritten
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 11);

  const uint32_t warp = warp_idx(), lane = lane_idx();
  long long* const tlp = tl_begin();
  if (threadIdx.x == 64) tl_stamp(0);
  const int b = blockIdx.z, h = blockIdx.y;
  const int q0 = blockIdx.x * 128;
  const int nkv = (Sk + 127) / 128;
  // key 0 of this batch row (look-ahead mask only): the load is in flight during the prologue below, its value is
  // first needed for the tile count after the block-wide sync
  const uint8_t lead_byte = (mask.causal && mask.kpad != nullptr) ? mask.kpad[(long long)b * Sk] : (uint8_t)0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV); prefetch_tmap(&tmO);
    mbar_init(q_full, 1); mbar_init(v_full, 1); mbar_init(v_empty, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
    mbar_init(s_full, 1); mbar_init(p_full, 128); mbar_init(o_full, 1); mbar_init(s_free, 128);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_ptr);
  tcgen05_before_thread_sync();
  __syncthreads();
  tcgen05_after_thread_sync();
  const uint32_t tmem_base = *tmem_ptr;
  if (threadIdx.x == 64) tl_stamp(1);
  // kv tiles this query tile must visit.  Causal rows only see keys <= query index -- unless key 0 of this batch row is
  // padding, in which case some rows may be fully masked and (reference semantics) attend uniformly to ALL keys.
  int nt = nkv;
  if (mask.causal && lead_byte == 0) nt = min(nkv, min(q0 + 127, Sq - 1) / 128 + 1);
  const uint32_t tS = tmem_base, tO = tmem_base + 128;

  if (warp == 0) {
    if (lane == 0) {
      // ===================================== TMA producer =====================================
      mbar_arrive_expect_tx(q_full, C::TILE_BYTES);
      tile_load<DK>(smem + L::OFF_Q, &tmQ, q_full, h * DK, q0, b);
      for (int j = 0; j < nt; ++j) {
        const int st = j % KS;
        mbar_wait(&k_empty[st], ((j / KS) & 1) ^ 1);
        mbar_arrive_expect_tx(&k_full[st], C::TILE_BYTES);
        tile_load<DK>(smem + L::OFF_K + st * C::TILE_BYTES, &tmK, &k_full[st], h * DK, j * 128, b);
        mbar_wait(v_empty, (j & 1) ^ 1);
        mbar_arrive_expect_tx(v_full, C::TILE_BYTES);
        tile_load<DK>(smem + L::OFF_V, &tmV, v_full, h * DK, j * 128, b);
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =====================================
    constexpr uint32_t idescS = make_idesc_bf16(128, 128, 0, 0);   // S[q,kv] = Q (K-major) . K^T (K-major)
    constexpr uint32_t idescO = make_idesc_bf16(128, DK, 0, 1);    // O[q,dk] = P (K-major) . V (MN-major)
    const uint32_t sQ = smem_u32(smem + L::OFF_Q), sK = smem_u32(smem + L::OFF_K), sV = smem_u32(smem + L::OFF_V),
                   sP = smem_u32(smem + L::OFF_P);
    auto issue_S = [&](int j) {
      const int st = j % KS;
      mbar_wait(&k_full[st], (j / KS) & 1);
      if (j == 0) mbar_wait(q_full, 0);
      tcgen05_after_thread_sync();
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < DK / 16; ++k) {
          const uint64_t ad = make_smem_desc(sQ + C::kmajor_k16(k), 0, C::SBO, C::SWZ);
          const uint64_t bd = make_smem_desc(sK + st * C::TILE_BYTES + C::kmajor_k16(k), 0, C::SBO, C::SWZ);
          umma_f16(tS, ad, bd, idescS, k > 0 ? 1u : 0u);
        }
        umma_commit(s_full);
        umma_commit(&k_empty[st]);
      }
      __syncwarp();
    };
    auto issue_PV = [&](int j) {
      mbar_wait(p_full, j & 1);
      mbar_wait(v_full, j & 1);
      tcgen05_after_thread_sync();
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t ad = make_smem_desc(sP + (k >> 2) * 16384 + (k & 3) * 32, 0, 1024, kSwz128);
          const uint64_t bd = make_smem_desc(sV + k * C::KSTEP_MN, C::PANEL_BYTES, C::SBO, C::SWZ);
          umma_f16(tO, ad, bd, idescO, (j > 0 || k > 0) ? 1u : 0u);   // O accumulates in TMEM over the kv tiles
        }
        umma_commit(o_full);
        umma_commit(v_empty);
      }
      __syncwarp();
    };
    issue_S(0);
    // S(j+1) is issued as soon as the softmax warps have pulled S(j) into registers (s_free) -- it then runs on the tensor
    // core while they exponentiate tile j -- and P.V(j) once P(j) is in smem.
    for (int j = 0; j < nt; ++j) {
      if (j + 1 < nt) {
        mbar_wait(s_free, j & 1);
        tcgen05_after_thread_sync();
        issue_S(j + 1);
      }
      issue_PV(j);
    }
  } else {
    // ===================================== softmax warps =====================================
    const uint32_t quad = warp & 3u;             // TMEM lane quadrant
    const uint32_t row = quad * 32u + lane;
    const int q = q0 + (int)row;                 // global query index
    const int wg_tid = (int)threadIdx.x - 64;
    const uint32_t lane_sel = (quad * 32u) << 16;
    const uint32_t sPs = smem_u32(smem + L::OFF_P);
    const float* mrow = kMaskF ? mask.maskf + (long long)b * mask.sb + (long long)min(q, Sq - 1) * mask.sq : nullptr;
    const float sl2 = scale * kLog2e;
    // ---- key-padding bits of every kv tile this CTA visits: loaded ONCE into smem (a global load per tile on the
    //      critical path costs ~1000 cycles); bit e of word 4*tile+c <-> key 128*tile + 32*c + e
    uint32_t* padw = reinterpret_cast<uint32_t*>(smem + L::OFF_PADW);
    const bool pad_pre = mask.kpad != nullptr && nt <= kMaxPadTiles;
    if (pad_pre) {
      const int nwords = nt * 4;
      const int w_id = (int)warp - 2;                         // 0..3
      for (int w0 = w_id; w0 < nwords; w0 += 16) {            // 4 independent loads in flight per warp
        uint8_t v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int kv = (w0 + 4 * u) * 32 + (int)lane;
          v[u] = (w0 + 4 * u < nwords && kv < Sk) ? mask.kpad[(long long)b * Sk + kv] : (uint8_t)0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t bits = __ballot_sync(0xffffffffu, v[u] != 0);
          if (lane == 0 && w0 + 4 * u < nwords) padw[w0 + 4 * u] = bits;
        }
      }
      named_bar_sync(1, 128);
    }
    // running state, in the log2 domain (logit * log2 e).  The subtracted maximum `m_used` is STALE by design: it is only
    // raised (and O, l rescaled) when the tile maximum exceeds it by more than 8 (a factor 2^8), so P stays <= 256 and
    // the per-tile O correction is rare.
    float m_used = -INFINITY, l_run = 0.f;
    // a hidden key holds the RAW score kMaskRaw: times scale it is the reference's finite -1e9 logit (Attention.py:25-26)
    const float kMaskRaw = kMaskVal / scale;

    for (int j = 0; j < nt; ++j) {
      const int kv0 = j * 128;
      // ---- per-thread mask words of a 32-key chunk: bit e of bm = key kv0+32c+e is hidden from this query (padding or
      //      look-ahead: finite -1e9 logit), of bo = the key does not exist (-inf).  ONE code path serves every mask; a chunk
      //      with no bit set in the whole warp takes the clean body.
      auto mask_words = [&](int c, uint32_t& bm, uint32_t& bo) {
        const int k0 = kv0 + c * 32;
        bo = k0 + 32 <= Sk ? 0u : (k0 >= Sk ? 0xffffffffu : (0xffffffffu << (Sk - k0)));
        uint32_t pb;
        if (mask.kpad == nullptr) pb = 0u;
        else if (pad_pre) pb = padw[j * 4 + c];
        else {
          const int kv = k0 + (int)lane;
          pb = __ballot_sync(0xffffffffu, kv < Sk && mask.kpad[(long long)b * Sk + kv] != 0);
        }
        if (mask.causal) {                       // keys k0 + e > q are hidden
          const int dq = q - k0;
          pb |= dq >= 31 ? 0u : (dq < 0 ? 0xffffffffu : (0xfffffffeu << dq));
        }
        bm = pb;
      };
      // raw score -> masked raw score (in place)
      auto apply_mask = [&](int c, uint32_t bm, uint32_t bo, uint32_t (&rc)[32]) {
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          float t = __uint_as_float(rc[e]);
          if (bm & (1u << e)) t = kMaskRaw;
          if (kMaskF) { const int kv = kv0 + c * 32 + e; if (kv < Sk) t += mrow[kv] * kMaskRaw; }
          if (bo & (1u << e)) t = -INFINITY;
          rc[e] = __float_as_uint(t);
        }
      };
      const bool tl = (wg_tid == 0);
      if (tl) tl_stamp(8 + j * 8 + 0);
      mbar_wait(s_full, j & 1);
      tcgen05_after_thread_sync();
      if (tl) tl_stamp(8 + j * 8 + 1);
#if BX_FWD_SINGLE_PASS
      // ---- ONE round trip to tensor memory per tile: the whole 128-column row of this thread is loaded with a single
      //      tcgen05.wait::ld (a load + wait round trip costs ~250 cycles; the two-pass variant below pays eight per tile) and S is
      //      handed back at once, so S(j+1) runs under this tile's softmax.  The row is processed IN PLACE (raw score -> masked
      //      score -> probability) and leaves in groups of 8 (pack, one 16-byte shared-memory store) to stay under 168 registers.
      uint32_t r[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_x32(tS + lane_sel + c * 32, r[c]);
      tmem_ld_wait();
      tcgen05_before_thread_sync();
      mbar_arrive(s_free);
      float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      uint32_t hidden_bits = 0;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t bm, bo;
        mask_words(c, bm, bo);
        const bool touched = kMaskF || __any_sync(0xffffffffu, (bm | bo) != 0u);
        if (!kMaskF && __all_sync(0xffffffffu, bm == 0xffffffffu)) hidden_bits |= 1u << c;
        if (touched) apply_mask(c, bm, bo, r[c]);
#pragma unroll
        for (int e = 0; e < 32; ++e) mx[e & 3] = fmaxf(mx[e & 3], __uint_as_float(r[c][e]));
      }
      const float m_tile = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * sl2;     // sl2 > 0
      if (tl) tl_stamp(8 + j * 8 + 2);
      const bool need = (j > 0) && (m_tile - m_used > 8.f);
      if (j == 0) m_used = m_tile;
      if (__any_sync(0xffffffffu, need)) {
        const float alpha = need ? fast_exp2(m_used - m_tile) : 1.f;
        if (need) { m_used = m_tile; l_run *= alpha; }
        mbar_wait(o_full, (j - 1) & 1);                // P.V of the previous tile has landed in TMEM
        tcgen05_after_thread_sync();
#pragma unroll 1
        for (int c = 0; c < DK / 32; ++c) {
          uint32_t ro[32];
          tmem_ld_x32(tO + lane_sel + c * 32, ro);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) ro[e] = __float_as_uint(__uint_as_float(ro[e]) * alpha);
          tmem_st_x32(tO + lane_sel + c * 32, ro);
        }
        tmem_st_wait();
      }
      const bool real_max = __all_sync(0xffffffffu, m_used > -1e8f);
      if (j > 0) { mbar_wait(o_full, (j - 1) & 1); tcgen05_after_thread_sync(); }
      float rs4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint32_t panel = sPs + (c >> 1) * 16384;
        const bool zero = ((hidden_bits >> c) & 1u) && real_max;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t pk[4];
#pragma unroll
          for (int h2 = 0; h2 < 4; ++h2) {
            const int e = g * 8 + h2 * 2;
            const float p0 = zero ? 0.f : fast_exp2(__uint_as_float(r[c][e]) * sl2 - m_used);
            const float p1 = zero ? 0.f : fast_exp2(__uint_as_float(r[c][e + 1]) * sl2 - m_used);
            rs4[h2] += p0 + p1;
            pk[h2] = pack_bf16x2(p0, p1);
          }
          st_shared_v4(panel + swz128_off(row, (c & 1) * 4 + g), pk[0], pk[1], pk[2], pk[3]);
        }
      }
#else
      // ---- pass 1 over tensor memory: row maximum of the masked scores.  The row is NOT kept in registers (a 128-value row
      //      + packed P + masks needed more than the 168 registers two resident CTAs leave a thread: ~80 local-memory
      //      round trips per tile, and a 30 KB unrolled loop body); both passes are ROLLED loops over 32-key chunks that re-read S
      //      from tensor memory (16 KB per warp and pass, ~100 cycles of latency per chunk, hidden by the other resident CTA).
      float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      // the mask words of the tile are classified ONCE (warp-uniform bit sets): the mask words cost more than the 16 FMNMX3 of a
      // clean chunk (removing them was worth 11 % of the kernel, profiles/r02_attention_fwd_ablation.txt), so clean chunks of both
      // passes never recompute them
      uint32_t hidden_bits = 0, touched_bits = 0;    // bit c: chunk c is hidden for every row of this warp / needs the mask applied
      // most tiles need no mask at all -- every key exists, lies at or below the smallest query index of this warp (look-ahead) and
      // none is padding: three warp-uniform tests (one 16-byte shared-memory load for the padding bits) instead of four chunk
      // classifications with their shifts and votes (~300 instructions per tile)
      bool tile_clean = !kMaskF && kv0 + 128 <= Sk && (!mask.causal || kv0 + 127 <= q0 + (int)quad * 32);
      if (tile_clean && mask.kpad != nullptr) {
        if (pad_pre) {
          const uint4 w4 = *reinterpret_cast<const uint4*>(padw + j * 4);
          tile_clean = (w4.x | w4.y | w4.z | w4.w) == 0u;
        } else {
          tile_clean = false;
        }
      }
      if (!tile_clean) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t bm, bo;
          mask_words(c, bm, bo);
          if (kMaskF || __any_sync(0xffffffffu, (bm | bo) != 0u)) touched_bits |= 1u << c;
          if (!kMaskF && __all_sync(0xffffffffu, bm == 0xffffffffu)) hidden_bits |= 1u << c;
        }
      }
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        if ((hidden_bits >> c) & 1u) { mx[0] = fmaxf(mx[0], kMaskRaw); continue; }     // every existing key of the chunk holds the masked value
        uint32_t rc[32];
        tmem_ld_x32(tS + lane_sel + c * 32, rc);
        tmem_ld_wait();
        if ((touched_bits >> c) & 1u) {
          uint32_t bm, bo;
          mask_words(c, bm, bo);
          apply_mask(c, bm, bo, rc);
        }
#pragma unroll
        for (int e = 0; e < 32; ++e) mx[e & 3] = fmaxf(mx[e & 3], __uint_as_float(rc[e]));
      }
      const float m_tile = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * sl2;     // sl2 > 0
      if (tl) tl_stamp(8 + j * 8 + 2);
      // ---- lazy rescale
      const bool need = (j > 0) && (m_tile - m_used > 8.f);
      if (j == 0) m_used = m_tile;
      if (__any_sync(0xffffffffu, need)) {
        const float alpha = need ? fast_exp2(m_used - m_tile) : 1.f;
        if (need) { m_used = m_tile; l_run *= alpha; }
        mbar_wait(o_full, (j - 1) & 1);                // P.V of the previous tile has landed in TMEM
        tcgen05_after_thread_sync();
#pragma unroll 1
        for (int c = 0; c < DK / 32; ++c) {
          uint32_t ro[32];
          tmem_ld_x32(tO + lane_sel + c * 32, ro);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) ro[e] = __float_as_uint(__uint_as_float(ro[e]) * alpha);
          tmem_st_x32(tO + lane_sel + c * 32, ro);
        }
        tmem_st_wait();
      }
      // ---- pass 2: p = 2^(raw * scale * log2 e - m_used), row sum (fp32, 4 partial sums), P -> bf16 -> smem, chunk by chunk.
      //      A chunk hidden for the whole warp contributes exactly 0 (2^(-1.44e9 - m) underflows) unless some row of the
      //      warp is fully masked so far (m_used is then the -1e9 level itself and those keys count: uniform attention)
      const bool real_max = __all_sync(0xffffffffu, m_used > -1e8f);
      // the P smem tile is still being read by P.V of the previous kv tile until o_full fires (S(j) was issued BEFORE that P.V,
      // so s_full alone does not order them); it fired long ago
      if (j > 0) { mbar_wait(o_full, (j - 1) & 1); tcgen05_after_thread_sync(); }
      float rs4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t pk[16];
        const bool hid = (hidden_bits >> c) & 1u;
        if (hid && real_max) {
#pragma unroll
          for (int g = 0; g < 16; ++g) pk[g] = 0u;
        } else {
          uint32_t rc[32];
          if (hid) {
#pragma unroll
            for (int e = 0; e < 32; ++e) rc[e] = __float_as_uint(kMaskRaw);
          } else {
            tmem_ld_x32(tS + lane_sel + c * 32, rc);
            tmem_ld_wait();
          }
          if ((touched_bits >> c) & 1u) {
            uint32_t bm, bo;
            mask_words(c, bm, bo);
            apply_mask(c, bm, bo, rc);
          }
          float p[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) { p[e] = fast_exp2(__uint_as_float(rc[e]) * sl2 - m_used); rs4[e & 3] += p[e]; }
#pragma unroll
          for (int g = 0; g < 16; ++g) pk[g] = pack_bf16x2(p[2 * g], p[2 * g + 1]);
        }
        if (c == 3) {                                  // every chunk of S has been read: the next tile's S may overwrite it
          tcgen05_before_thread_sync();
          mbar_arrive(s_free);
        }
        const uint32_t panel = sPs + (c >> 1) * 16384;
#pragma unroll
        for (int g = 0; g < 4; ++g)
          st_shared_v4(panel + swz128_off(row, (c & 1) * 4 + g), pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
      }
#endif
      l_run += (rs4[0] + rs4[1]) + (rs4[2] + rs4[3]);
      if (tl) tl_stamp(8 + j * 8 + 3);
      fence_proxy_async_smem();
      tcgen05_before_thread_sync();
      mbar_arrive(p_full);
      if (tl) tl_stamp(8 + j * 8 + 4);
    }
    // finalise: O (TMEM) / l -> bf16 -> smem (the P buffer) -> TMA store ; lse = (m_used + log2 l) ln 2
    mbar_wait(o_full, (nt - 1) & 1);
    tcgen05_after_thread_sync();
    if (wg_tid == 0) tl_stamp(2);
    const float inv = 1.f / l_run;
    // Training also keeps the bf16 ROUNDING RESIDUAL of O (O_lo = bf16(O - bf16(O)), staged in the Q tile, which no MMA reads
    // any more): the backward's delta = rowsum(dO o O) cancels against dP = dO.V^T in dS = P o (dP - delta), and any component
    // shared by all value vectors (bias, LayerNorm beta) makes that cancellation amplify the 2^-9 rounding of O into
    // percent-level errors of dQ / dK (measured: tools/attn_precision.py).  delta is formed from O + O_lo.
    const uint32_t sLo = smem_u32(smem + L::OFF_Q);
#pragma unroll
    for (int c = 0; c < DK / 32; ++c) {
      uint32_t ro[32];
      tmem_ld_x32(tO + lane_sel + c * 32, ro);
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float a = __uint_as_float(ro[8 * g + 2 * e]) * inv, c2 = __uint_as_float(ro[8 * g + 2 * e + 1]) * inv;
          hi[e] = pack_bf16x2(a, c2);
          const float2 rr = unpack_bf16x2(hi[e]);
          lo[e] = pack_bf16x2(a - rr.x, c2 - rr.y);
        }
        st_shared_v4(sPs + swz_off<DK>(row, c * 4 + g), hi[0], hi[1], hi[2], hi[3]);
        if (has_lo) st_shared_v4(sLo + swz_off<DK>(row, c * 4 + g), lo[0], lo[1], lo[2], lo[3]);
      }
    }
    if (q < Sq) {
      float lse_val = (m_used + __log2f(l_run)) * 0.6931471805599453f;   // back to natural log
      // a fully-masked row: every logit is the reference's finite -1e9 (times the mask value); fp32 absorbs ln(Sk) there, and
      // raw * scale only reproduces -1e9 to ~64 ulp-units when scale is not a power of two (head depth 32, 128): snap, so that
      // the backward / attention_weights recompute exp(-1e9 - lse) = 1 exactly like the reference's own fp32 softmax
      if (m_used < -1e8f) lse_val = kMaskVal * rintf(lse_val / kMaskVal);
      lse_out[((long long)b * H + h) * Sq + q] = lse_val;
    }
    tcgen05_before_thread_sync();
    fence_proxy_async_smem();
    named_bar_sync(1, 128);
    if (wg_tid == 0) {
      tile_store<DK>(&tmO, smem + L::OFF_P, h * DK, q0, b);
      if (has_lo) tile_store<DK>(&tmOlo, smem + L::OFF_Q, h * DK, q0, b);
      tma_store_commit();
      tma_store_wait_read<0>();          // smem may be released; the write itself completes before the grid does
      tl_stamp(3);
    }
  }
  tcgen05_before_thread_sync();
  __syncthreads();
  tcgen05_after_thread_sync();
  if (warp == 1) tmem_dealloc<256>(tmem_base);
}

// ================================================================================================
// backward:  one CTA per (kv tile, head, batch); loops over query tiles.
//   S^T = K.Q^T, dP^T = V.dO^T (TMEM)  ->  P^T = exp(S^T*scale - lse), dS^T = P^T o (dP^T - delta) * scale  (smem, bf16)
//   dV += P^T.dO   dK += dS^T.Q   (TMEM accumulators over all query tiles)    dQ_tile = dS.K  -> added into global dQ (bf16: red.global.v4.bf16x2; fp32: TMA reduce-add)
// ================================================================================================
// Head depth 128: the five accumulators S^T, dP^T, dV, dK, dQ (+ P^T) would need 704 tensor-memory columns, so every
// (kv tile, head, batch) is handled by NH = 2 CTAs; both compute the full S^T / dP^T (contraction over all 128 channels) and
// each produces ONE 64-column half (DKO) of dV, dK and dQ.  Shared memory then only holds one Q/dO stage (QS = 1) and the fp32
// dQ staging is dropped (bf16 dQ only).  A slower, correct path for the reference's default configuration (4 heads of 128).
template <int DK>
struct BwdSmem {
  using C = AttnCfg<DK>;
  static constexpr int DKO = DK <= 64 ? DK : 64;                    // output columns per CTA
  static constexpr int NH = DK / DKO;                               // CTAs per (kv tile, head, batch)
  static constexpr int QS = DK <= 64 ? 2 : 1;                       // Q / dO stages
  static constexpr uint32_t OFF_K = 0;
  static constexpr uint32_t OFF_V = OFF_K + C::TILE_BYTES;
  static constexpr uint32_t OFF_Q = OFF_V + C::TILE_BYTES;          // QS stages
  static constexpr uint32_t OFF_DO = OFF_Q + QS * C::TILE_BYTES;    // QS stages
  static constexpr uint32_t OFF_DST = OFF_DO + QS * C::TILE_BYTES;  // dS^T [kv 128][q 128] bf16, two panels, DOUBLE buffered
                                                                    // (P^T lives in tensor memory, not in smem)
  static constexpr uint32_t OFF_DQ = OFF_DST + 2 * C::P_BYTES;      // fp32-dQ variant only: dQ tile fp32 [128 x DKO] as two [128 x DKO/2] boxes (TMA reduce-add staging)
  static constexpr uint32_t DQ_BYTES = DK <= 64 ? 2 * 128 * (DKO / 2) * 4 : 0;
  static constexpr uint32_t OFF_ROW = OFF_DQ + DQ_BYTES;            // lse/delta: 2 buffers x (128 + 128) floats
  static constexpr uint32_t OFF_BAR = OFF_ROW + 2 * 256 * 4;
  static constexpr uint32_t BYTES = OFF_BAR + 256 + 1024;
};

constexpr int ATTN_BWD_THREADS = 576;  // w0 TMA, w1 MMA(+TMEM alloc), w2-17 compute: 4 lane quadrants x 4 groups of 32 query columns
constexpr int ATTN_BWD_CT = 512;       // compute threads

template <int DK, bool kDqBf16>
__global__ void __launch_bounds__(ATTN_BWD_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                const __grid_constant__ CUtensorMap tmdK, const __grid_constant__ CUtensorMap tmdV,
                const __grid_constant__ CUtensorMap tmdQ, const float* __restrict__ lse, const float* __restrict__ delta,
                AttnMask mask, int H, int Sq, int Sk, float scale, __nv_bfloat16* dq16, long long ld_dq16) {
  using C = AttnCfg<DK>;
  using L = BwdSmem<DK>;
  constexpr int DKO = L::DKO, NH = L::NH, QS = L::QS;
  static_assert(kDqBf16 || DK <= 64, "head depth 128: bf16 dQ only");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* kv_full = bars;          // [1]
  uint64_t* qdo_full = bars + 1;     // [2]
  uint64_t* qdo_empty = bars + 3;    // [2]
  uint64_t* sdp_full = bars + 5;     // [1]
  uint64_t* pds_full = bars + 6;     // [1] count 256
  uint64_t* dq_full = bars + 7;      // [1]
  uint64_t* pt_free = bars + 8;      // [1] dV MMAs have finished reading P^T from tensor memory
  uint64_t* dq_free = bars + 9;      // [1] count 256: the compute warps have read the dQ tile out of tensor memory
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 10);
  float* rowbuf = reinterpret_cast<float*>(smem + L::OFF_ROW);

  const uint32_t warp = warp_idx(), lane = lane_idx();
  long long* const tlp = tl_begin();
  if (threadIdx.x == 64) tl_stamp(60);
  const int b = blockIdx.z, h = blockIdx.y, jt = blockIdx.x / NH;
  const int ho = blockIdx.x % NH;                 // which DKO-column half of dV / dK / dQ this CTA produces (head depth 128)
  const uint32_t ho_panel = (uint32_t)ho * C::PANEL_BYTES;
  const int kv0 = jt * 128;
  const int nq = (Sq + 127) / 128;
  const bool lead_pad = mask.causal && mask.kpad != nullptr && mask.kpad[(long long)b * Sk] != 0;
  const int i_start = (mask.causal && !lead_pad) ? min(jt, nq - 1) : 0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV); prefetch_tmap(&tmdO);
    prefetch_tmap(&tmdK); prefetch_tmap(&tmdV); prefetch_tmap(&tmdQ);
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&qdo_full[i], 1); mbar_init(&qdo_empty[i], 1); }
    mbar_init(sdp_full, 1); mbar_init(pds_full, ATTN_BWD_CT); mbar_init(dq_full, 1); mbar_init(pt_free, 1); mbar_init(dq_free, ATTN_BWD_CT);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr);
  tcgen05_before_thread_sync();
  __syncthreads();
  tcgen05_after_thread_sync();
  const uint32_t tmem_base = *tmem_ptr;
  // tensor memory (512 columns): S^T 128 | dP^T 128 | dV dk | dK dk | dQ dk | P^T (bf16 pairs) 64
  const uint32_t tSt = tmem_base, tdPt = tmem_base + 128, tdV = tmem_base + 256, tdK = tmem_base + 256 + DKO,
                 tdQ = tmem_base + 256 + 2 * DKO, tPt = tmem_base + 256 + 3 * DKO;

  if (warp == 0) {
    if (lane == 0) {
      // ===================================== TMA producer =====================================
      mbar_arrive_expect_tx(kv_full, 2 * C::TILE_BYTES);
      tile_load<DK>(smem + L::OFF_K, &tmK, kv_full, h * DK, kv0, b);
      tile_load<DK>(smem + L::OFF_V, &tmV, kv_full, h * DK, kv0, b);
      int n = 0;
      for (int i = i_start; i < nq; ++i, ++n) {
        const int st = n % QS;
        mbar_wait(&qdo_empty[st], ((n / QS) & 1) ^ 1);
        mbar_arrive_expect_tx(&qdo_full[st], 2 * C::TILE_BYTES);
        tile_load<DK>(smem + L::OFF_Q + st * C::TILE_BYTES, &tmQ, &qdo_full[st], h * DK, i * 128, b);
        tile_load<DK>(smem + L::OFF_DO + st * C::TILE_BYTES, &tmdO, &qdo_full[st], h * DK, i * 128, b);
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =====================================
    constexpr uint32_t idescS = make_idesc_bf16(128, 128, 0, 0);    // [kv, q] = K/V (K-major) . Q/dO^T (K-major)
    constexpr uint32_t idescKV = make_idesc_bf16(128, DKO, 0, 1);   // [kv, dk] = P^T/dS^T (K-major) . dO/Q (MN-major)
    constexpr uint32_t idescQ = make_idesc_bf16(128, DKO, 1, 1);    // [q, dk]  = dS (MN-major view of dS^T) . K (MN-major)
    const uint32_t sK = smem_u32(smem + L::OFF_K), sV = smem_u32(smem + L::OFF_V), sQ = smem_u32(smem + L::OFF_Q),
                   sdO = smem_u32(smem + L::OFF_DO), sdSt = smem_u32(smem + L::OFF_DST);
    auto issue_sdp = [&](int n) {
      const int st = n % QS;
      mbar_wait(&qdo_full[st], (n / QS) & 1);
      tcgen05_after_thread_sync();
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < DK / 16; ++k) {
          const uint64_t ad = make_smem_desc(sK + C::kmajor_k16(k), 0, C::SBO, C::SWZ);
          const uint64_t bd = make_smem_desc(sQ + st * C::TILE_BYTES + C::kmajor_k16(k), 0, C::SBO, C::SWZ);
          umma_f16(tSt, ad, bd, idescS, k > 0 ? 1u : 0u);
        }
#pragma unroll
        for (int k = 0; k < DK / 16; ++k) {
          const uint64_t ad = make_smem_desc(sV + C::kmajor_k16(k), 0, C::SBO, C::SWZ);
          const uint64_t bd = make_smem_desc(sdO + st * C::TILE_BYTES + C::kmajor_k16(k), 0, C::SBO, C::SWZ);
          umma_f16(tdPt, ad, bd, idescS, k > 0 ? 1u : 0u);
        }
        umma_commit(sdp_full);
      }
      __syncwarp();
    };
    mbar_wait(kv_full, 0);
    const int ntile = nq - i_start;
    issue_sdp(0);
    // per tile the tensor core runs   dV(n) | S^T,dP^T(n+1) | dK(n), dQ(n)   so the compute warps can start on tile n+1
    // (their S^T/dP^T are ready after the first two groups) while dK/dQ of tile n are still executing.  (Issuing S^T/dP^T(n+1)
    // even earlier -- as soon as the compute warps hold tile n's values in registers -- measured slower: 250 vs 236 us.)
    for (int n = 0; n < ntile; ++n) {
      const int st = n % QS;
      const uint32_t sdS = sdSt + (n & 1) * C::P_BYTES;
      mbar_wait(pds_full, n & 1);
      tcgen05_after_thread_sync();
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {   // dV += P^T . dO     (K = 128 queries; A = P^T in tensor memory, 8 columns per k16)
          const uint64_t bd = make_smem_desc(sdO + st * C::TILE_BYTES + ho_panel + k * C::KSTEP_MN, C::PANEL_BYTES, C::SBO, C::SWZ);
          umma_f16_ts(tdV, tPt + k * 8, bd, idescKV, (n > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(pt_free);
      }
      __syncwarp();
      if (QS == 2 && n + 1 < ntile) issue_sdp(n + 1);     // (one Q/dO stage: the next tile's operands only arrive after dK/dQ below)
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {   // dK += dS^T . Q
          const uint64_t ad = make_smem_desc(sdS + (k >> 2) * 16384 + (k & 3) * 32, 0, 1024, kSwz128);
          const uint64_t bd = make_smem_desc(sQ + st * C::TILE_BYTES + ho_panel + k * C::KSTEP_MN, C::PANEL_BYTES, C::SBO, C::SWZ);
          umma_f16(tdK, ad, bd, idescKV, (n > 0 || k > 0) ? 1u : 0u);
        }
      }
      if (n > 0) { mbar_wait(dq_free, (n - 1) & 1); tcgen05_after_thread_sync(); }   // dQ(n-1) has been read out of TMEM
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {   // dQ = dS . K        (K = 128 keys; A = MN-major view of dS^T)
          const uint64_t ad = make_smem_desc(sdS + k * 2048, 16384, 1024, kSwz128);
          const uint64_t bd = make_smem_desc(sK + ho_panel + k * C::KSTEP_MN, C::PANEL_BYTES, C::SBO, C::SWZ);
          umma_f16(tdQ, ad, bd, idescQ, k > 0 ? 1u : 0u);
        }
        umma_commit(&qdo_empty[st]);
        umma_commit(dq_full);
      }
      __syncwarp();
      if (QS == 1 && n + 1 < ntile) issue_sdp(n + 1);
    }
  } else {
    // ===================================== compute warps =====================================
    // 16 warps: warp (quad, grp) owns kv rows 32*quad.. (tensor-memory lane quadrant = warp & 3) and the 32 query columns
    // 32*grp.. of S^T / dP^T.  (Round 1 used 8 warps x 64 columns: two warps per scheduler left every latency of the
    // ld -> exp2 -> pack -> st chain exposed: issue-active 26 %, 2750 cycles of compute per tile; profiles/r02_ncu_full_attn_bwd.txt.)
    const int cw = (int)warp - 2;                 // 0..15
    const uint32_t quad = warp & 3u;
    const int grp = cw >> 2;                      // which 32 query columns / which quarter of the dk columns this warp owns
    const uint32_t row = quad * 32u + lane;       // kv row (S^T, dV, dK) or q row (dQ)
    const uint32_t lane_sel = (quad * 32u) << 16;
    const int ctid = (int)threadIdx.x - 64;       // 0..511
    const int kv = kv0 + (int)row;
    const bool kv_valid = kv < Sk;
    const bool kv_pad = kv_valid && mask.kpad != nullptr && mask.kpad[(long long)b * Sk + kv] != 0;
    const uint32_t sdSt = smem_u32(smem + L::OFF_DST);
    const float sl2 = scale * kLog2e;
    const float inv_sk = 1.f / (float)Sk;
    const uint32_t sDQ = smem_u32(smem + L::OFF_DQ);
    constexpr int HC = DKO / 4;                   // dk columns per warp in the dQ / dV / dK epilogues (16 or 8)
    // no padded / out-of-range key among this warp's rows (fixed for the whole CTA)
    const bool rows_clean = !__any_sync(0xffffffffu, kv_pad || !kv_valid);
    // ... and among ALL 128 keys of this CTA's tile (every lane looks at 4 of them)
    bool dirty = kv0 + 128 > Sk;
    if (!dirty && mask.kpad != nullptr) {
      const uint8_t* kp = mask.kpad + (long long)b * Sk + kv0 + (int)lane * 4;
      dirty = (kp[0] | kp[1] | kp[2] | kp[3]) != 0;
    }
    const bool tile_clean = !__any_sync(0xffffffffu, dirty);
    // lse (x log2 e) / delta of a query tile: thread ctid < 128 fetches lse[q], ctid 128..255 delta[q] (the other compute threads
    // only wait).  The value for the NEXT tile travels global -> shared memory by cp.async one tile ahead (no register in between)
    // and is scaled in place once it has landed.
    const bool row_loader = ctid < 256;
    const float* const row_src = (ctid < 128 ? lse : delta) + ((long long)b * H + h) * Sq;
    const float row_oob = ctid < 128 ? INFINITY : 0.f;                  // +inf lse for out-of-range queries -> p = 0
    const float row_mul = ctid < 128 ? kLog2e : scale;                  // lse in the log2 domain; delta pre-multiplied by 1/sqrt(dk)
    auto prefetch_row = [&](int i, float* buf) {
      if (!row_loader) return;
      const int qq = i * 128 + (ctid & 127);
      if (qq < Sq) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(smem_u32(buf + ctid)), "l"(row_src + qq) : "memory");
      } else {
        buf[ctid] = row_oob;
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // this warp's HC columns of a [128 x DKO] fp32 accumulator
    auto load_acc = [&](uint32_t taddr, float (&v)[HC]) {
      if (HC == 16) {
        uint32_t r[16];
        tmem_ld_x16(taddr, r); tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e < HC ? e : 0] = __uint_as_float(r[e]);
      } else {
        uint32_t r[8];
        tmem_ld_x8(taddr, r); tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e < HC ? e : 0] = __uint_as_float(r[e]);
      }
    };
    // dQ tile of query tile `qt`: row = query, this warp's quarter of the dk columns
    auto dq_epilogue = [&](int n_done, int qt0) {
      mbar_wait(dq_full, n_done & 1);
      tcgen05_after_thread_sync();
      if (ctid == 0) tl_stamp(64 + (n_done & 7) * 8 + 4);
      float v[HC];
      if (kDqBf16) {
        // bf16 dQ: straight from tensor memory to global memory.  This thread owns query row `row` and HC contiguous columns:
        // red.global.add.noftz.v4.bf16x2 adds 16 bytes per instruction, fire-and-forget -- no smem staging, no block barrier, no
        // TMA.  (One thread issuing a TMA reduce-add stalled ~1300 cycles at its next memory instruction, and the loop-top
        // barrier passed that on to all compute threads, every tile.)
        load_acc(tdQ + lane_sel + grp * HC, v);
        tcgen05_before_thread_sync();
        mbar_arrive(dq_free);                    // the dQ accumulator may be overwritten by the next tile's MMA
        const int qq = qt0 + (int)row;
        if (qq < Sq) {
          __nv_bfloat16* dst = dq16 + ((long long)b * Sq + qq) * ld_dq16 + h * DK + ho * DKO + grp * HC;
#pragma unroll
          for (int g = 0; g < HC / 8; ++g)
            asm volatile("red.global.add.noftz.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(dst + 8 * g), "r"(pack_bf16x2(v[8 * g], v[8 * g + 1])),
                         "r"(pack_bf16x2(v[8 * g + 2], v[8 * g + 3])), "r"(pack_bf16x2(v[8 * g + 4], v[8 * g + 5])),
                         "r"(pack_bf16x2(v[8 * g + 6], v[8 * g + 7]))
                         : "memory");
        }
        if (ctid == 0) tl_stamp(64 + (n_done & 7) * 8 + 5);
        return;
      }
      // fp32 dQ (tests, other callers): -> fp32 smem staging (two [128 x DKO/2] boxes) -> one TMA reduce-add per box
      // (cp.reduce.async.bulk.tensor .add.f32; rows >= Sq are clipped by TMA)
      if (ctid == 0) tma_store_wait_read<0>();   // the previous reduce has drained the staging
      named_bar_sync(2, ATTN_BWD_CT);
      const uint32_t stg = sDQ + (grp >> 1) * (L::DQ_BYTES / 2);
      load_acc(tdQ + lane_sel + grp * HC, v);
#pragma unroll
      for (int g = 0; g < HC / 4; ++g)
        st_shared_v4(stg + swz_off<DK>(row, (grp & 1) * (HC / 4) + g), __float_as_uint(v[4 * g]), __float_as_uint(v[4 * g + 1]),
                     __float_as_uint(v[4 * g + 2]), __float_as_uint(v[4 * g + 3]));
      tcgen05_before_thread_sync();
      mbar_arrive(dq_free);                      // the dQ accumulator may be overwritten by the next tile's MMA
      fence_proxy_async_smem();
      named_bar_sync(3, ATTN_BWD_CT);
      if (ctid == 0) {
        tma_reduce_add_3d(&tmdQ, smem + L::OFF_DQ, h * DK, qt0, b);
        tma_reduce_add_3d(&tmdQ, smem + L::OFF_DQ + L::DQ_BYTES / 2, h * DK + DKO / 2, qt0, b);
        tma_store_commit();
        tl_stamp(64 + (n_done & 7) * 8 + 5);
      }
    };
    prefetch_row(i_start, rowbuf);
    int n = 0;
    const int col0 = grp * 32;                    // first query column of this warp's chunk
    for (int i = i_start; i < nq; ++i, ++n) {
      const int qt0 = i * 128;
      float* rb = rowbuf + (n & 1) * 256;
      if (row_loader) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");      // this thread's element of tile n has landed ...
        rb[ctid] *= row_mul;                                      // ... scale it in place (inf / 0 stay what they are)
      }
      if (ctid == 0) tl_stamp(64 + (n & 7) * 8 + 0);
      named_bar_sync(1, ATTN_BWD_CT);                             // every thread is past tile n-1: its row buffer may be refilled
      if (i + 1 < nq) prefetch_row(i + 1, rowbuf + ((n + 1) & 1) * 256);
      const bool diag = mask.causal && (qt0 < kv0 + 128);   // tile touches the causal boundary (some q < kv)
      const bool plain = rows_clean && !diag && mask.maskf == nullptr;   // no mask logic needed in this tile
      // aligned diagonal tile of a look-ahead mask, nothing else to mask: query col is hidden from key `row` iff col < row.
      // Per warp (keys 32*quad..) and its 32-query chunk grp: grp > quad plain, grp < quad fully hidden (P = dS = 0: no loads, no
      // exponentials), grp == quad mixed.  With the WHOLE key tile clean every query of the tile sees at least its own key, so
      // no query row here is fully masked (those attend uniformly to all keys, hidden ones included: general path).
      const bool diag_fast = diag && tile_clean && mask.maskf == nullptr && qt0 == kv0;
      const int wk = (int)(row >> 5), lk = (int)(row & 31u);
      mbar_wait(sdp_full, n & 1);
      tcgen05_after_thread_sync();
      if (ctid == 0) tl_stamp(64 + (n & 7) * 8 + 2);
      const uint32_t sdS = sdSt + (n & 1) * C::P_BYTES + (grp >> 1) * 16384;   // this tile's dS^T buffer, this warp's 64-query panel
      uint32_t pw[16];                                                          // P^T row of this thread, 32 queries as bf16 pairs
      const bool hidden = diag_fast && grp < wk;
      // two sub-chunks of 16 query columns: S^T / dP^T values, probabilities and dS of a sub-chunk are the only big live arrays
      // (the compiler has 96 registers per thread with 18 warps resident)
#pragma unroll
      for (int sub = 0; sub < 2; ++sub) {
        const int c16 = col0 + sub * 16;
        uint32_t rs[16], rp[16];
        float p[16], ds[16];
        if (!hidden) {
          tmem_ld_x16(tSt + lane_sel + c16, rs);
          tmem_ld_x16(tdPt + lane_sel + c16, rp);
          tmem_ld_wait();
        }
        const uint32_t rb_l = smem_u32(rb + c16), rb_d = smem_u32(rb + 128 + c16);       // lse (log2 domain) / delta of the 16 queries
        if (hidden) {
#pragma unroll
          for (int e = 0; e < 16; ++e) { p[e] = 0.f; ds[e] = 0.f; }
        } else if (plain || diag_fast) {
          const bool mixed = diag_fast && grp == wk;          // the chunk on the diagonal: query column e is visible iff e >= lk
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4) {
            const float4 l4 = ld_shared_f4(rb_l + e4 * 16), d4 = ld_shared_f4(rb_d + e4 * 16);
            const float ll[4] = {l4.x, l4.y, l4.z, l4.w}, dd[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int e = e4 * 4 + k;
              float pe = fast_exp2(__uint_as_float(rs[e]) * sl2 - ll[k]);
              if (mixed && sub * 16 + e < lk) pe = 0.f;
              p[e] = pe;
              ds[e] = pe * (__uint_as_float(rp[e]) * scale - dd[k]);      // = p (dP - delta) / sqrt(dk)
            }
          }
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int qi = qt0 + c16 + e;
            float t2 = __uint_as_float(rs[e]) * sl2;                        // logit * log2(e)
            const bool masked = kv_pad || (diag && kv > qi);
            if (masked) t2 = kMaskVal * kLog2e;
            if (mask.maskf != nullptr && kv_valid && qi < Sq)
              t2 += mask.maskf[(long long)b * mask.sb + (long long)qi * mask.sq + kv] * (kMaskVal * kLog2e);
            const float l2 = rb[c16 + e];
            float pe = fast_exp2(t2 - l2);
            if (l2 < -1e8f) pe *= inv_sk;     // fully-masked query row: lse == -1e9 in fp32 (ln Sk absorbed) -> uniform 1/Sk
            if (!kv_valid) pe = 0.f;
            p[e] = pe;
            ds[e] = pe * (__uint_as_float(rp[e]) * scale - rb[128 + c16 + e]);
          }
        }
#pragma unroll
        for (int g = 0; g < 8; ++g) pw[sub * 8 + g] = pack_bf16x2(p[2 * g], p[2 * g + 1]);
#pragma unroll
        for (int g = 0; g < 2; ++g)
          st_shared_v4(sdS + swz128_off(row, (grp & 1) * 4 + sub * 2 + g), pack_bf16x2(ds[8 * g], ds[8 * g + 1]), pack_bf16x2(ds[8 * g + 2], ds[8 * g + 3]),
                       pack_bf16x2(ds[8 * g + 4], ds[8 * g + 5]), pack_bf16x2(ds[8 * g + 6], ds[8 * g + 7]));
      }
      // P^T -> tensor memory (A operand of the dV MMA), once the previous tile's dV MMAs have finished reading it
      if (n > 0) { mbar_wait(pt_free, (n - 1) & 1); tcgen05_after_thread_sync(); }
      tmem_st_x16(tPt + lane_sel + grp * 16, pw);
      tmem_st_wait();
      if (ctid == 0) tl_stamp(64 + (n & 7) * 8 + 3);
      fence_proxy_async_smem();
      tcgen05_before_thread_sync();
      mbar_arrive(pds_full);
      // the PREVIOUS tile's dQ is drained while the tensor core works on this tile's dV and the next tile's S^T/dP^T
      if (n > 0) dq_epilogue(n - 1, qt0 - 128);
    }
    dq_epilogue(n - 1, (nq - 1) * 128);
    // epilogue: dV, dK (TMEM) -> bf16 -> smem (the two dS^T buffers, free once the last dq_full has fired) -> TMA store
    {
      float v[HC], k[HC];
      load_acc(tdV + lane_sel + grp * HC, v);
      load_acc(tdK + lane_sel + grp * HC, k);
#pragma unroll
      for (int g = 0; g < HC / 8; ++g) {
        const uint32_t off = swz_off<DK>(row, grp * (HC / 8) + g);
        st_shared_v4(sdSt + off, pack_bf16x2(v[8 * g], v[8 * g + 1]), pack_bf16x2(v[8 * g + 2], v[8 * g + 3]),
                     pack_bf16x2(v[8 * g + 4], v[8 * g + 5]), pack_bf16x2(v[8 * g + 6], v[8 * g + 7]));
        st_shared_v4(sdSt + C::P_BYTES + off, pack_bf16x2(k[8 * g], k[8 * g + 1]), pack_bf16x2(k[8 * g + 2], k[8 * g + 3]),
                     pack_bf16x2(k[8 * g + 4], k[8 * g + 5]), pack_bf16x2(k[8 * g + 6], k[8 * g + 7]));
      }
      fence_proxy_async_smem();
      named_bar_sync(1, ATTN_BWD_CT);
      if (ctid == 0) {
        tl_stamp(61);
        tma_store_3d(&tmdV, smem + L::OFF_DST, h * DK + ho * DKO, kv0, b);      // one box: DKO <= the panel width
        tma_store_3d(&tmdK, smem + L::OFF_DST + C::P_BYTES, h * DK + ho * DKO, kv0, b);
        tma_store_commit();
        tma_store_wait_read<0>();
        tl_stamp(62);
      }
    }
  }
  tcgen05_before_thread_sync();
  __syncthreads();
  tcgen05_after_thread_sync();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ================================================================================================
// attention-weights materialisation (Decoder.py:75-76 returns them; train_step discards them): recompute
// P[b,h,q,:] = exp(logit - lse) from Q, K and the forward's lse.  Inspection/parity path, plain SIMT.
// ================================================================================================
__global__ void __launch_bounds__(128)
attn_probs_kernel(const __nv_bfloat16* __restrict__ Q, long long ld_q, const __nv_bfloat16* __restrict__ K, long long ld_k,
                  const float* __restrict__ lse, float* __restrict__ P, AttnMask mask, int H, int Sq, int Sk, int dk,
                  float scale) {
  const int q = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  __shared__ float qs[128];
  const __nv_bfloat16* qp = Q + ((long long)b * Sq + q) * ld_q + h * dk;
  for (int c = threadIdx.x; c < dk; c += blockDim.x) qs[c] = __bfloat162float(qp[c]);
  __syncthreads();
  const float l = lse[((long long)b * H + h) * Sq + q];
  float* out = P + (((long long)b * H + h) * Sq + q) * Sk;
  for (int k = threadIdx.x; k < Sk; k += blockDim.x) {
    const __nv_bfloat16* kp = K + ((long long)b * Sk + k) * ld_k + h * dk;
    float s = 0.f;
    for (int c = 0; c < dk; c += 2) {
      const float2 f = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(kp + c));
      s += qs[c] * f.x + qs[c + 1] * f.y;
    }
    float t = s * scale;
    const bool masked = (mask.kpad != nullptr && mask.kpad[(long long)b * Sk + k] != 0) || (mask.causal && k > q);
    if (masked) t = kMaskVal;
    if (mask.maskf != nullptr) t += mask.maskf[(long long)b * mask.sb + (long long)q * mask.sq + k] * kMaskVal;
    float pe = __expf(t - l);
    if (l < -1e8f) pe *= 1.f / (float)Sk;   // fully-masked row (see attn_bwd_kernel)
    out[k] = pe;
  }
}

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
template <int DK>
static int attn_fwd_launch(const void* Q, long long ld_q, const void* K, long long ld_k, const void* V, long long ld_v,
                           void* O, long long ld_o, void* O_lo, float* lse, const AttnMask& mask, int B, int H, int Sq, int Sk,
                           cudaStream_t st, long long bs_q, long long bs_kv) {
  using L = FwdSmem<DK>;
  constexpr int PW = AttnCfg<DK>::PW;      // TMA box = one panel: [128 tokens x PW columns]
  const int swz = PW * 2;
  CUtensorMap tmQ, tmK, tmV, tmO;
  int rc;
  // batch strides (elements) of Q and of K/V: by default the tensors are dense [B, S, ld]; a key/value CACHE has its own pitch
  const uint64_t sq = bs_q > 0 ? (uint64_t)bs_q : (uint64_t)Sq * ld_q, sk = bs_kv > 0 ? (uint64_t)bs_kv : (uint64_t)Sk * ld_k,
                 sv = bs_kv > 0 ? (uint64_t)bs_kv : (uint64_t)Sk * ld_v;
  if ((rc = encode_tmap_3d(&tmQ, 0, Q, (uint64_t)H * DK, Sq, B, ld_q, sq, PW, 128, 1, swz))) return rc;
  if ((rc = encode_tmap_3d(&tmK, 0, K, (uint64_t)H * DK, Sk, B, ld_k, sk, PW, 128, 1, swz))) return rc;
  if ((rc = encode_tmap_3d(&tmV, 0, V, (uint64_t)H * DK, Sk, B, ld_v, sv, PW, 128, 1, swz))) return rc;
  if ((rc = encode_tmap_3d(&tmO, 0, O, (uint64_t)H * DK, Sq, B, ld_o, (uint64_t)Sq * ld_o, PW, 128, 1, swz))) return rc;
  CUtensorMap tmOlo = tmO;
  if (O_lo != nullptr && (rc = encode_tmap_3d(&tmOlo, 0, O_lo, (uint64_t)H * DK, Sq, B, ld_o, (uint64_t)Sq * ld_o, PW, 128, 1, swz))) return rc;
  auto kern = mask.maskf != nullptr ? attn_fwd_kernel<DK, true> : attn_fwd_kernel<DK, false>;
  {
    static std::mutex mu;
    static bool attr_done[kMaxDevices] = {};
    const int dev = current_device();
    std::lock_guard<std::mutex> lock(mu);
    if (!attr_done[dev]) {
      BX_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<DK, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::BYTES));
      BX_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<DK, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::BYTES));
      // two CTAs per SM need the full shared-memory carve-out
      BX_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<DK, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
      BX_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<DK, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
      attr_done[dev] = true;
    }
  }
  dim3 grid((Sq + 127) / 128, H, B);
  void* pb = prof_on() ? prof_begin(st) : nullptr;
  kern<<<grid, ATTN_FWD_THREADS, L::BYTES, st>>>(tmQ, tmK, tmV, tmO, tmOlo, O_lo != nullptr ? 1 : 0, lse, mask, H, Sq, Sk, 1.f / sqrtf((float)DK));
  if (pb) prof_end(pb, 1, 4.0 * B * H * (double)Sq * Sk * DK, st);     // dense count, as the reference computes it
  BX_POST_LAUNCH();
  return 0;
}

template <int DK, bool kDqBf16>
static int attn_bwd_launch(const void* Q, long long ld_q, const void* K, long long ld_k, const void* V, long long ld_v,
                           const void* dO, long long ld_do, const float* lse, const float* delta, void* dq_acc,
                           long long ld_dq, void* dK, long long ld_dk, void* dV, long long ld_dv, const AttnMask& mask,
                           int B, int H, int Sq, int Sk, cudaStream_t st) {
  using L = BwdSmem<DK>;
  constexpr int PW = AttnCfg<DK>::PW;
  const int swz = PW * 2;
  CUtensorMap tmQ, tmK, tmV, tmdO, tmdK, tmdV, tmdQ;
  int rc;
  if ((rc = encode_tmap_3d(&tmQ, 0, Q, (uint64_t)H * DK, Sq, B, ld_q, (uint64_t)Sq * ld_q, PW, 128, 1, swz))) return rc;
  if ((rc = encode_tmap_3d(&tmK, 0, K, (uint64_t)H * DK, Sk, B, ld_k, (uint64_t)Sk * ld_k, PW, 128, 1, swz))) return rc;
  if ((rc = encode_tmap_3d(&tmV, 0, V, (uint64_t)H * DK, Sk, B, ld_v, (uint64_t)Sk * ld_v, PW, 128, 1, swz))) return rc;
  if ((rc = encode_tmap_3d(&tmdO, 0, dO, (uint64_t)H * DK, Sq, B, ld_do, (uint64_t)Sq * ld_do, PW, 128, 1, swz))) return rc;
  if ((rc = encode_tmap_3d(&tmdK, 0, dK, (uint64_t)H * DK, Sk, B, ld_dk, (uint64_t)Sk * ld_dk, PW, 128, 1, swz))) return rc;
  if ((rc = encode_tmap_3d(&tmdV, 0, dV, (uint64_t)H * DK, Sk, B, ld_dv, (uint64_t)Sk * ld_dv, PW, 128, 1, swz))) return rc;
  if (kDqBf16) rc = encode_tmap_3d(&tmdQ, 0, dq_acc, (uint64_t)H * DK, Sq, B, ld_dq, (uint64_t)Sq * ld_dq, PW, 128, 1, swz);   // (unused: red.global)
  else rc = encode_tmap_3d(&tmdQ, 1, dq_acc, (uint64_t)H * DK, Sq, B, ld_dq, (uint64_t)Sq * ld_dq, DK / 2, 128, 1, swz);
  if (rc) return rc;
  auto kern = attn_bwd_kernel<DK, kDqBf16>;
  {
    static std::mutex mu;
    static bool attr_done[kMaxDevices] = {};
    const int dev = current_device();
    std::lock_guard<std::mutex> lock(mu);
    if (!attr_done[dev]) {
      BX_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::BYTES));
      attr_done[dev] = true;
    }
  }
  dim3 grid(((Sk + 127) / 128) * L::NH, H, B);
  void* pb = prof_on() ? prof_begin(st) : nullptr;
  kern<<<grid, ATTN_BWD_THREADS, L::BYTES, st>>>(tmQ, tmK, tmV, tmdO, tmdK, tmdV, tmdQ, lse, delta, mask, H, Sq, Sk,
                                                 1.f / sqrtf((float)DK), (__nv_bfloat16*)dq_acc, ld_dq);
  if (pb) prof_end(pb, 2, 8.0 * B * H * (double)Sq * Sk * DK, st);     // 2x the forward's 4.B.H.Sq.Sk.dk
  BX_POST_LAUNCH();
  return 0;
}

static int check_attn_args(int B, int H, int Sq, int Sk, int dk, int causal) {
  BX_REQUIRE(dk == 32 || dk == 64 || dk == 128, "attention: head depth %d unsupported (32, 64 or 128)", dk);
  BX_REQUIRE(B > 0 && H > 0 && Sq > 0 && Sk > 0 && B < 65536 && H < 65536, "attention: bad shape");
  BX_REQUIRE(!causal || Sq == Sk, "attention: causal (look-ahead) mask needs Sq == Sk");
  return 0;
}

int attn_fwd(const void* Q, long long ld_q, const void* K, long long ld_k, const void* V, long long ld_v, void* O,
             long long ld_o, float* lse, const uint8_t* kpad, int causal, const float* maskf, long long mask_sb,
             long long mask_sq, int B, int H, int Sq, int Sk, int dk, cudaStream_t st, void* O_lo, long long bs_q, long long bs_kv) {
  if (int rc = check_attn_args(B, H, Sq, Sk, dk, causal)) return rc;
  AttnMask m{kpad, maskf, mask_sb, mask_sq, causal};
  if (dk == 128) return attn_fwd_launch<128>(Q, ld_q, K, ld_k, V, ld_v, O, ld_o, O_lo, lse, m, B, H, Sq, Sk, st, bs_q, bs_kv);
  return dk == 64 ? attn_fwd_launch<64>(Q, ld_q, K, ld_k, V, ld_v, O, ld_o, O_lo, lse, m, B, H, Sq, Sk, st, bs_q, bs_kv)
                  : attn_fwd_launch<32>(Q, ld_q, K, ld_k, V, ld_v, O, ld_o, O_lo, lse, m, B, H, Sq, Sk, st, bs_q, bs_kv);
}

int attn_bwd(const void* Q, long long ld_q, const void* K, long long ld_k, const void* V, long long ld_v, const void* dO,
             long long ld_do, const float* lse, const float* delta, void* dq_acc, long long ld_dq, int dq_bf16, void* dK,
             long long ld_dk, void* dV, long long ld_dv, const uint8_t* kpad, int causal, const float* maskf,
             long long mask_sb, long long mask_sq, int B, int H, int Sq, int Sk, int dk, cudaStream_t st) {
  if (int rc = check_attn_args(B, H, Sq, Sk, dk, causal)) return rc;
  BX_REQUIRE(ld_dq % 8 == 0 && (reinterpret_cast<uintptr_t>(dq_acc) & 15) == 0, "attention: dq accumulator alignment");
  AttnMask m{kpad, maskf, mask_sb, mask_sq, causal};
#define BX_BWD_GO(D, BF) \
  return attn_bwd_launch<D, BF>(Q, ld_q, K, ld_k, V, ld_v, dO, ld_do, lse, delta, dq_acc, ld_dq, dK, ld_dk, dV, ld_dv, m, B, H, Sq, Sk, st)
  if (dk == 128) {
    BX_REQUIRE(dq_bf16, "attention backward at head depth 128 accumulates dQ in bf16 only (dq_is_bf16 = 1)");
    BX_BWD_GO(128, true);
  }
  if (dk == 64) { if (dq_bf16) BX_BWD_GO(64, true); else BX_BWD_GO(64, false); }
  if (dq_bf16) BX_BWD_GO(32, true); else BX_BWD_GO(32, false);
#undef BX_BWD_GO
}

int attn_probs(const void* Q, long long ld_q, const void* K, long long ld_k, const float* lse, float* P,
               const uint8_t* kpad, int causal, const float* maskf, long long mask_sb, long long mask_sq, int B, int H,
               int Sq, int Sk, int dk, cudaStream_t st) {
  BX_REQUIRE(dk <= 128 && dk % 2 == 0, "attn_probs: dk %d unsupported", dk);
  AttnMask m{kpad, maskf, mask_sb, mask_sq, causal};
  attn_probs_kernel<<<dim3(Sq, H, B), 128, 0, st>>>((const __nv_bfloat16*)Q, ld_q, (const __nv_bfloat16*)K, ld_k, lse, P, m,
                                                    H, Sq, Sk, dk, 1.f / sqrtf((float)dk));
  BX_POST_LAUNCH();
  return 0;
}

}  // namespace bx
