// Whole-model orchestration of the hot path: Transformer.call (Transformer.py:21-32) -> Encoder.call
// (Encoder.py:48-60) / Decoder.call (Decoder.py:59-79) -> final_layer, the loss (train.py:83-88) and the
// hand-written reverse pass that replaces tape.gradient (train.py:137).  One C call launches every kernel of a
// phase on the caller's stream; nothing here allocates, synchronises or touches the host.
//
// PARAMETER ARENA (fp32; grads / Adam m / Adam v / bf16 shadow use the same offsets).  Order follows keras'
// trainable_variables order at block granularity; Q/K/V kernels of one MultiHeadAttention are PACKED side by side
// so the three Dense layers (Attention.py:62-64) run as one GEMM:
//   enc_emb [Vin,d]
//   per encoder layer: Wqkv [d,3d] bqkv [3d] Wo [d,d] bo [d] W1 [d,f] b1 [f] W2 [f,d] b2 [d] ln1g ln1b ln2g ln2b [d]
//   dec_emb [Vtgt,d]
//   per decoder layer: Wqkv bqkv Wo bo | Wq [d,d] bq [d] Wkv [d,2d] bkv [2d] Wo2 bo2 | W1 b1 W2 b2 | ln1g ln1b ln2g ln2b ln3g ln3b
//   Wf [d,Vp] bf [Vp]                 (Vp = target vocab rounded up to 8 columns for 16-byte TMA row pitch)
// b200xf_model_variable() exposes the keras-order per-variable (strided) views, e.g. wq.kernel = Wqkv[:, 0:d].
#include <stdlib.h>
#include <string.h>

#include <string>
#include <utility>
#include <vector>

#include "../../include/b200xf.h"
#include "common.cuh"
#include "kernels.h"

namespace {

using bf16 = __nv_bfloat16;

inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

struct MhaSelf { int64_t Wqkv, bqkv, Wo, bo; };
struct MhaCross { int64_t Wq, bq, Wkv, bkv, Wo, bo; };
struct Ffn { int64_t W1, b1, W2, b2; };
struct EncLayerP { MhaSelf mha; Ffn ffn; int64_t ln1g, ln1b, ln2g, ln2b; };
struct DecLayerP { MhaSelf mha1; MhaCross mha2; Ffn ffn; int64_t ln1g, ln1b, ln2g, ln2b, ln3g, ln3b; };

struct VarView { std::string name; int64_t off, rows, cols, ld; };

struct Layout {
  int L, d, H, f, Vin, Vtgt, Vp;
  int64_t enc_emb, dec_emb, Wf, bf, total;
  int64_t stage_begin[3], stage_end[3];   // 0 final layer, 1 decoder (+emb), 2 encoder (+emb)
  std::vector<std::pair<int64_t, int64_t>> enc_layer_range, dec_layer_range;   // arena range of every layer (gradient buckets)
  std::pair<int64_t, int64_t> enc_emb_range, dec_emb_range;
  std::vector<EncLayerP> enc;
  std::vector<DecLayerP> dec;
  std::vector<VarView> vars;
};

struct Bump {
  int64_t cur = 0;
  int64_t take(int64_t n) { int64_t o = cur; cur = align_up(cur + n, 64); return o; }   // 256-byte aligned (fp32)
};

Layout make_layout(const b200xf_model_config& c) {
  Layout lo;
  lo.L = c.num_layers; lo.d = c.d_model; lo.H = c.num_heads; lo.f = c.dff;
  lo.Vin = c.input_vocab_size; lo.Vtgt = c.target_vocab_size; lo.Vp = (int)align_up(c.target_vocab_size, 8);
  const int64_t d = lo.d, f = lo.f;
  Bump b;
  auto self = [&](MhaSelf& m) { m.Wqkv = b.take(d * 3 * d); m.bqkv = b.take(3 * d); m.Wo = b.take(d * d); m.bo = b.take(d); };
  auto ffn = [&](Ffn& m) { m.W1 = b.take(d * f); m.b1 = b.take(f); m.W2 = b.take(f * d); m.b2 = b.take(d); };
  lo.stage_begin[2] = b.cur;
  lo.enc_emb = b.take((int64_t)lo.Vin * d);
  lo.enc_emb_range = {lo.enc_emb, b.cur};
  lo.enc.resize(lo.L);
  for (auto& e : lo.enc) {
    const int64_t begin = b.cur;
    self(e.mha); ffn(e.ffn);
    e.ln1g = b.take(d); e.ln1b = b.take(d); e.ln2g = b.take(d); e.ln2b = b.take(d);
    lo.enc_layer_range.push_back({begin, b.cur});
  }
  lo.stage_end[2] = b.cur;
  lo.stage_begin[1] = b.cur;
  lo.dec_emb = b.take((int64_t)lo.Vtgt * d);
  lo.dec_emb_range = {lo.dec_emb, b.cur};
  lo.dec.resize(lo.L);
  for (auto& e : lo.dec) {
    const int64_t begin = b.cur;
    self(e.mha1);
    e.mha2.Wq = b.take(d * d); e.mha2.bq = b.take(d); e.mha2.Wkv = b.take(d * 2 * d); e.mha2.bkv = b.take(2 * d);
    e.mha2.Wo = b.take(d * d); e.mha2.bo = b.take(d);
    ffn(e.ffn);
    e.ln1g = b.take(d); e.ln1b = b.take(d); e.ln2g = b.take(d); e.ln2b = b.take(d); e.ln3g = b.take(d); e.ln3b = b.take(d);
    lo.dec_layer_range.push_back({begin, b.cur});
  }
  lo.stage_end[1] = b.cur;
  lo.stage_begin[0] = b.cur;
  lo.Wf = b.take(d * (int64_t)lo.Vp);
  lo.bf = b.take(lo.Vp);
  lo.stage_end[0] = b.cur;
  lo.total = b.cur;

  // keras-order variable views
  auto& v = lo.vars;
  auto dense = [&](const std::string& p, int64_t k_off, int64_t rows, int64_t cols, int64_t ld, int64_t b_off) {
    v.push_back({p + "/kernel", k_off, rows, cols, ld});
    v.push_back({p + "/bias", b_off, cols, 0, 1});
  };
  auto self_vars = [&](const std::string& p, const MhaSelf& m) {
    dense(p + "/wq", m.Wqkv, d, d, 3 * d, m.bqkv);
    dense(p + "/wk", m.Wqkv + d, d, d, 3 * d, m.bqkv + d);
    dense(p + "/wv", m.Wqkv + 2 * d, d, d, 3 * d, m.bqkv + 2 * d);
    dense(p + "/dense", m.Wo, d, d, d, m.bo);
  };
  auto ffn_vars = [&](const std::string& p, const Ffn& m) {
    dense(p + "/dense0", m.W1, d, f, f, m.b1);
    dense(p + "/dense1", m.W2, f, d, d, m.b2);
  };
  auto ln_vars = [&](const std::string& p, int64_t g, int64_t be) {
    v.push_back({p + "/gamma", g, d, 0, 1});
    v.push_back({p + "/beta", be, d, 0, 1});
  };
  v.push_back({"encoder/embedding", lo.enc_emb, lo.Vin, d, d});
  for (int i = 0; i < lo.L; ++i) {
    const std::string p = "encoder/layer" + std::to_string(i);
    self_vars(p + "/mha", lo.enc[i].mha);
    ffn_vars(p + "/ffn", lo.enc[i].ffn);
    ln_vars(p + "/layernorm1", lo.enc[i].ln1g, lo.enc[i].ln1b);
    ln_vars(p + "/layernorm2", lo.enc[i].ln2g, lo.enc[i].ln2b);
  }
  v.push_back({"decoder/embedding", lo.dec_emb, lo.Vtgt, d, d});
  for (int i = 0; i < lo.L; ++i) {
    const std::string p = "decoder/layer" + std::to_string(i);
    const DecLayerP& e = lo.dec[i];
    self_vars(p + "/mha1", e.mha1);
    dense(p + "/mha2/wq", e.mha2.Wq, d, d, d, e.mha2.bq);
    dense(p + "/mha2/wk", e.mha2.Wkv, d, d, 2 * d, e.mha2.bkv);
    dense(p + "/mha2/wv", e.mha2.Wkv + d, d, d, 2 * d, e.mha2.bkv + d);
    dense(p + "/mha2/dense", e.mha2.Wo, d, d, d, e.mha2.bo);
    ffn_vars(p + "/ffn", e.ffn);
    ln_vars(p + "/layernorm1", e.ln1g, e.ln1b);
    ln_vars(p + "/layernorm2", e.ln2g, e.ln2b);
    ln_vars(p + "/layernorm3", e.ln3g, e.ln3b);
  }
  dense("final_layer", lo.Wf, d, lo.Vtgt, lo.Vp, lo.bf);
  return lo;
}

// ---- workspace (activations kept for backward + gradient scratch), carved from one caller-owned buffer ----
struct LnSave { bf16* z; float *mean, *rstd; uint8_t* keep; };   // z = x + dropout(y) (bf16) and the row statistics
struct EncAct { bf16* qkv; bf16* attn_o; bf16* attn_o_lo; float* lse; LnSave ln1; bf16* out1_16; bf16* h; uint64_t* hbits; LnSave ln2; };
struct DecAct {
  bf16* qkv1; bf16* attn1_o; bf16* attn1_o_lo; float* lse1; LnSave ln1; bf16* out1_16;   // *_o_lo: bf16 rounding residual of the attention output
  bf16* q2; bf16* kv2; bf16* attn2_o; bf16* attn2_o_lo; float* lse2; LnSave ln2; bf16* out2_16;
  bf16* h; uint64_t* hbits; LnSave ln3;
};
struct Work {
  std::vector<bf16*> enc_x16, dec_x16;    // [L+1] layer inputs/outputs (index 0 = embedding output); the streams are bf16
  bf16* y16;                              // [Mmax, d] sublayer output on its way into the residual LayerNorm
  std::vector<EncAct> enc;
  std::vector<DecAct> dec;
  bf16* logits;                           // [Mt, Vp] bf16, overwritten in place by dlogits
  uint8_t *kpad_in, *kpad_tar;
  long long* dec_hist;                    // [max_batch, max_tgt_len] decoder input tokens seen so far (incremental decoding)
  float* dec_lse;                         // [max_batch, H] scratch lse of a one-query attention
  // backward scratch
  bf16 *gR0, *gR1;                        // [Mmax, d] residual-path gradient (dz of the last LayerNorm backward), ping-pong
  bf16* gD;                               // [Mmax, d] dgrad output of the last sublayer; the stream gradient is gR + gD
  float* d_enc_out;                       // [Ms, d] fp32: accumulates cross-attention K/V gradients of all decoder layers
  // Buffers the weight-gradient GEMMs (side stream) read while the main stream moves on are RINGS: the main stream waits for
  // the last side reader of a slot before it overwrites that slot (see BwdCtx)
  bf16* dy16[3];                          // [Mmax,d] gradient of a sublayer output (one slot per LayerNorm of a decoder layer)
  bf16* dO16;                             // [Mmax,d] gradient of the attention output (main stream only)
  bf16* dqkv[2];                          // [Mmax,3d] packed self-attention Q/K/V gradient
  bf16* dkv2[2];                          // [Ms,2d]  cross-attention K/V gradient
  bf16* dh[2];                            // [Mmax,f] FFN hidden gradient
  bf16* dq16[2];                          // [Mmax,d] cross-attention dQ (bf16, accumulated by red.global adds)
  float* delta;                           // [B,H,Smax]
  int64_t bytes;
};

struct ByteBump {
  uint8_t* base; int64_t cur = 0;
  template <typename T> T* take(int64_t n) {
    T* p = base ? reinterpret_cast<T*>(base + cur) : nullptr;
    cur = align_up(cur + n * (int64_t)sizeof(T), 1024);
    return p;
  }
};

Work make_work(const b200xf_model_config& c, const Layout& lo, void* base) {
  Work w;
  ByteBump b{reinterpret_cast<uint8_t*>(base)};
  const int64_t d = lo.d, f = lo.f, H = lo.H, L = lo.L;
  const int64_t Ms = (int64_t)c.max_batch * c.max_src_len, Mt = (int64_t)c.max_batch * c.max_tgt_len;
  const int64_t Mx = Ms > Mt ? Ms : Mt;
  const int64_t Smax = c.max_src_len > c.max_tgt_len ? c.max_src_len : c.max_tgt_len;
  auto ln = [&](int64_t M) {
    LnSave s; s.z = b.take<bf16>(M * d); s.mean = b.take<float>(M); s.rstd = b.take<float>(M); s.keep = b.take<uint8_t>(M * (d / 8));
    return s;
  };
  w.enc_x16.resize(L + 1); w.dec_x16.resize(L + 1);
  for (int i = 0; i <= L; ++i) w.enc_x16[i] = b.take<bf16>(Ms * d);
  for (int i = 0; i <= L; ++i) w.dec_x16[i] = b.take<bf16>(Mt * d);
  w.y16 = b.take<bf16>(Mx * d);
  w.enc.resize(L); w.dec.resize(L);
  for (auto& a : w.enc) {
    a.qkv = b.take<bf16>(Ms * 3 * d); a.attn_o = b.take<bf16>(Ms * d); a.attn_o_lo = b.take<bf16>(Ms * d); a.lse = b.take<float>(c.max_batch * H * c.max_src_len);
    a.ln1 = ln(Ms); a.out1_16 = b.take<bf16>(Ms * d); a.h = b.take<bf16>(Ms * f); a.hbits = b.take<uint64_t>(Ms * ((f + 63) / 64)); a.ln2 = ln(Ms);
  }
  for (auto& a : w.dec) {
    a.qkv1 = b.take<bf16>(Mt * 3 * d); a.attn1_o = b.take<bf16>(Mt * d); a.attn1_o_lo = b.take<bf16>(Mt * d); a.lse1 = b.take<float>(c.max_batch * H * c.max_tgt_len);
    a.ln1 = ln(Mt); a.out1_16 = b.take<bf16>(Mt * d);
    a.q2 = b.take<bf16>(Mt * d); a.kv2 = b.take<bf16>(Ms * 2 * d); a.attn2_o = b.take<bf16>(Mt * d); a.attn2_o_lo = b.take<bf16>(Mt * d);
    a.lse2 = b.take<float>(c.max_batch * H * c.max_tgt_len);
    a.ln2 = ln(Mt); a.out2_16 = b.take<bf16>(Mt * d); a.h = b.take<bf16>(Mt * f); a.hbits = b.take<uint64_t>(Mt * ((f + 63) / 64)); a.ln3 = ln(Mt);
  }
  w.logits = b.take<bf16>(Mt * lo.Vp);
  w.kpad_in = b.take<uint8_t>(Ms); w.kpad_tar = b.take<uint8_t>(Mt);
  w.dec_hist = b.take<long long>(Mt); w.dec_lse = b.take<float>(c.max_batch * H);
  w.gR0 = b.take<bf16>(Mx * d); w.gR1 = b.take<bf16>(Mx * d); w.gD = b.take<bf16>(Mx * d); w.d_enc_out = b.take<float>(Ms * d);
  for (auto& p : w.dy16) p = b.take<bf16>(Mx * d);
  w.dO16 = b.take<bf16>(Mx * d);
  for (auto& p : w.dqkv) p = b.take<bf16>(Mx * 3 * d);
  for (auto& p : w.dkv2) p = b.take<bf16>(Ms * 2 * d);
  for (auto& p : w.dh) p = b.take<bf16>(Mx * f);
  for (auto& p : w.dq16) p = b.take<bf16>(Mt * d);
  w.delta = b.take<float>(c.max_batch * H * Smax);
  w.bytes = b.cur;
  return w;
}

int validate(const b200xf_model_config& c) {
  BX_REQUIRE(c.num_layers >= 1 && c.d_model >= 32 && c.num_heads >= 1 && c.dff >= 8, "model: bad sizes");
  BX_REQUIRE(c.d_model % c.num_heads == 0, "model: d_model %% num_heads != 0 (Attention.py:42)");
  const int dk = c.d_model / c.num_heads;
  BX_REQUIRE(dk == 32 || dk == 64 || dk == 128, "model: head depth %d unsupported (d_model/num_heads must be 32, 64 or 128)", dk);
  BX_REQUIRE(c.d_model % 8 == 0 && c.dff % 8 == 0 && c.d_model <= 1024, "model: d_model, dff must be multiples of 8; d_model <= 1024");
  BX_REQUIRE(c.input_vocab_size >= 2 && c.target_vocab_size >= 2, "model: vocab too small");
  BX_REQUIRE(c.max_batch >= 1 && c.max_src_len >= 1 && c.max_tgt_len >= 1, "model: bad workspace maxima");
  BX_REQUIRE(c.pos_rows >= c.max_src_len && c.pos_rows >= c.max_tgt_len, "model: positional table shorter than max length");
  BX_REQUIRE(c.dropout_rate >= 0.f && c.dropout_rate < 1.f, "model: dropout rate out of range");
  return 0;
}

}  // namespace

struct b200xf_model {
  b200xf_model_config cfg;
  Layout lo;
  Work w;
  float *P, *G, *Am, *Av;
  bf16* Sh;
  const float *pos_in, *pos_tgt;
  // state of the last forward
  int B = 0, S = 0, T = 0, training = 0;
  uint64_t seed = 0;
  const unsigned long long* seed_dev = nullptr;   // when set, kernels read the dropout seed from the device (graph replays)
  const int64_t *inp = nullptr, *tar = nullptr;
  int tar_stride = 0, tar_offset = 0;
  const float *enc_mask = nullptr, *la_mask = nullptr, *dec_mask = nullptr;
  int64_t la_bs = 0;
  bool have_loss_grad = false;
  bool logits_valid = false;   // w.logits holds the last forward's bf16 logits (not yet turned into dlogits)
  // second stream for work off the critical path (weight-gradient GEMMs, bias-gradient column sums, the cross-attention K/V
  // projections): forked from / joined to the caller's stream with events inside every C call, so the caller still sees ONE
  // stream -- and a CUDA-graph capture of the caller's stream records the fork as parallel branches
  cudaStream_t side = nullptr;
  std::vector<cudaEvent_t> events;
  int next_event = 0;
  int concurrency = 0;         // 0: everything on the caller's stream
  int fuse_ln = 1;             // LayerNorm in the out-projection / FFN dense1 GEMM epilogue (d_model 128 / 256 / 512)
  const int* t_valid = nullptr;   // device int: target positions >= *t_valid are shape-bucketing filler (NULL: none)
  int dec_B = 0, dec_S = 0, dec_len = 0;   // incremental decoding (b200xf_model_decode_begin / _step): batch, source length, tokens cached
  // state carried between the parts of a layer-granular backward (b200xf_model_backward_part): the gradient of the stream
  // tensor entering the next layer (see StreamGrad) and which of the two dz buffers is written next
  const void* bw_r = nullptr; bool bw_r_f32 = false; const void* bw_dg = nullptr;
  int bw_dz = 0;
};

namespace {

using M = b200xf_model;

#define RC(expr)                 \
  do {                           \
    int _rc = (expr);            \
    if (_rc) return _rc;         \
  } while (0)

// dropout site ids (any distinct constants)
inline uint32_t site_enc_embed() { return 1; }
inline uint32_t site_dec_embed() { return 2; }
inline uint32_t site_enc(int l, int k) { return 16 + l * 8 + k; }
inline uint32_t site_dec(int l, int k) { return 4096 + l * 8 + k; }

// y[M,N] = x[M,K] . W[K,N] + b        (keras Dense; W in keras [in,out] layout, row pitch ldw)
int linear_fwd(const bf16* x, int64_t ldx, const bf16* W, int64_t ldw, const float* b, void* y, int64_t ldy, bool y_f32,
               bool relu, int64_t Mr, int64_t N, int64_t K, cudaStream_t st, uint64_t* relu_bits = nullptr) {
  const bool bits = relu && relu_bits != nullptr && !y_f32 && N % 64 == 0;    // compact (y > 0) gate, 1 bit per element
  return bx::gemm_bf16(x, ldx, 0, W, ldw, 1, y, ldy, y_f32 ? 1 : 0, b, bits ? relu_bits : nullptr, bits ? N / 64 : 0,
                       (relu ? 1 : 0) | (bits ? 8 : 0), 1, Mr, N, K, st);
}
// dx[M,K] (+)= dy[M,N] . W^T          (W [K,N] row-major is exactly the K-major B operand of this product)
int linear_dgrad(const bf16* dy, int64_t ld_dy, const bf16* W, int64_t ldw, void* dx, int64_t ld_dx, bool dx_f32, bool accum,
                 const bf16* relu_aux, int64_t ld_aux, int64_t Mr, int64_t N, int64_t K, cudaStream_t st,
                 const uint64_t* relu_bits = nullptr, float* dx_colsum = nullptr) {
  // dx_colsum: += column sums of the (bf16) dx this call writes = the bias gradient of the layer below, from the epilogue's staging tile
  if (relu_bits != nullptr && !dx_f32 && K % 64 == 0)     // ReLU gate from the compact mask the forward epilogue stored
    return bx::gemm_bf16(dy, ld_dy, 0, W, ldw, 0, dx, ld_dx, 0, nullptr, relu_bits, K / 64, (accum ? 2 : 0) | 16, 1, Mr, K, N, st, dx_colsum);
  const int flags = (accum ? 2 : 0) | (relu_aux ? 4 : 0);
  return bx::gemm_bf16(dy, ld_dy, 0, W, ldw, 0, dx, ld_dx, dx_f32 ? 1 : 0, nullptr, relu_aux, ld_aux, flags, 1, Mr, K, N, st, dx_colsum);
}
// dW[K,N] += x[M,K]^T . dy[M,N] ; db[N] += colsum(dy)     (split-K over the token dimension, TMA reduce-add)
int linear_wgrad(const bf16* x, int64_t ldx, const bf16* dy, int64_t ld_dy, float* dW, int64_t ldw, float* db, int64_t Mr,
                 int64_t N, int64_t K, cudaStream_t st) {
  RC(bx::gemm_bf16(x, ldx, 1, dy, ld_dy, 1, dW, ldw, 1, nullptr, nullptr, 0, 2, /*splits: auto*/ 0, K, N, Mr, st));
  if (db) RC(bx::colsum_bf16(dy, ld_dy, db, Mr, (int)N, st));
  return 0;
}

struct MaskSpec { const uint8_t* kpad; int causal; const float* maskf; int64_t sb, sq; };

MaskSpec enc_self_mask(const M* m) {
  if (m->enc_mask) return {nullptr, 0, m->enc_mask, (int64_t)m->S, 0};
  return {m->w.kpad_in, 0, nullptr, 0, 0};
}
MaskSpec dec_self_mask(const M* m) {
  if (m->la_mask) return {nullptr, 0, m->la_mask, m->la_bs, (int64_t)m->T};
  return {m->w.kpad_tar, 1, nullptr, 0, 0};
}
MaskSpec dec_cross_mask(const M* m) {
  if (m->dec_mask) return {nullptr, 0, m->dec_mask, (int64_t)m->S, 0};
  return {m->w.kpad_in, 0, nullptr, 0, 0};
}

// out = LayerNorm(x + dropout(y)); y is the sublayer output in w.y16
int residual_ln(M* m, const bf16* x, int64_t g_off, int64_t b_off, bf16* out16, const LnSave& s, int64_t rows, uint32_t site,
                cudaStream_t st) {
  const float rate = m->training ? m->cfg.dropout_rate : 0.f;
  return bx::residual_ln_fwd(m->w.y16, x, m->P + g_off, m->P + b_off, s.z, out16, s.mean, s.rstd, s.keep, rows, m->lo.d, m->cfg.ln_eps, rate,
                             m->seed, site, st, m->seed_dev);
}

// ---- fork / join of the side stream ------------------------------------------------------------------------------------
// Par(m, st): `main` is the caller's stream; `side` is the model's second stream while concurrency is on (and per-launch
// profiling is off: its events would time overlapped kernels), else `main` itself -- every helper below then degenerates to
// plain in-order launches.  mark() records "side work so far"; wait(mark) makes the main stream wait for it.
struct Par {
  M* m;
  cudaStream_t main, side;
  bool forked;
  Par(M* m_, cudaStream_t st) : m(m_), main(st), side(st), forked(false) {}
  cudaEvent_t next_event() {
    cudaEvent_t e = m->events[m->next_event];
    m->next_event = (m->next_event + 1) % (int)m->events.size();
    return e;
  }
  int fork() {        // side stream joins the main stream's current position
    if (m->concurrency == 0 || m->side == nullptr || bx::prof_on()) return 0;
    cudaEvent_t e = next_event();
    BX_CHECK_CUDA(cudaEventRecord(e, main));
    BX_CHECK_CUDA(cudaStreamWaitEvent(m->side, e, 0));
    side = m->side;
    forked = true;
    return 0;
  }
  // the side stream additionally waits for what the MAIN stream has launched so far (its operands are produced there)
  int side_after_main() {
    if (!forked) return 0;
    cudaEvent_t e = next_event();
    BX_CHECK_CUDA(cudaEventRecord(e, main));
    BX_CHECK_CUDA(cudaStreamWaitEvent(side, e, 0));
    return 0;
  }
  cudaEvent_t mark() {                       // completion of the side work launched so far
    if (!forked) return nullptr;
    cudaEvent_t e = next_event();
    if (cudaEventRecord(e, side) != cudaSuccess) return nullptr;
    return e;
  }
  int wait(cudaEvent_t e) {                  // main stream waits for a mark
    if (forked && e != nullptr) BX_CHECK_CUDA(cudaStreamWaitEvent(main, e, 0));
    return 0;
  }
  int join() {                               // main stream waits for ALL side work; the side stream is released
    if (!forked) return 0;
    cudaEvent_t e = next_event();
    BX_CHECK_CUDA(cudaEventRecord(e, side));
    BX_CHECK_CUDA(cudaStreamWaitEvent(main, e, 0));
    forked = false;
    side = main;
    return 0;
  }
};

// out = LayerNorm(x + dropout(A . W + b)): ONE kernel (LayerNorm in the GEMM epilogue, gemm_ln.cu) when d_model is 128 / 256 /
// 512, else the Dense GEMM followed by the stand-alone dropout + residual + LayerNorm kernel
int dense_residual_ln(M* m, const bf16* A, int64_t lda, int64_t K, int64_t w_off, int64_t bias_off, const bf16* x, int64_t g_off, int64_t b_off,
                      bf16* out16, const LnSave& s, int64_t rows, uint32_t site, cudaStream_t st) {
  const int d = m->lo.d;
  const float rate = m->training ? m->cfg.dropout_rate : 0.f;
  if (m->fuse_ln && bx::gemm_ln_supported(rows, d, K))
    return bx::gemm_ln(A, lda, m->Sh + w_off, d, m->P + bias_off, x, m->P + g_off, m->P + b_off, s.z, out16, s.mean, s.rstd, s.keep, rows, d, K,
                       m->cfg.ln_eps, rate, m->seed, site, st, m->seed_dev);
  RC(linear_fwd(A, lda, m->Sh + w_off, d, m->P + bias_off, m->w.y16, d, false, false, rows, d, K, st));
  return residual_ln(m, x, g_off, b_off, out16, s, rows, site, st);
}

int forward_impl(M* m, float* logits_out, int64_t ld_logits, cudaStream_t st) {
  const Layout& lo = m->lo;
  Work& w = m->w;
  const int d = lo.d, H = lo.H, f = lo.f, L = lo.L, dk = d / H;
  const int B = m->B, S = m->S, T = m->T;
  const int64_t Ms = (int64_t)B * S, Mt = (int64_t)B * T;
  const float rate = m->training ? m->cfg.dropout_rate : 0.f;
  const bf16* Sh = m->Sh;

  RC(bx::pad_mask(reinterpret_cast<const long long*>(m->inp), S, 0, w.kpad_in, B, S, st));
  RC(bx::pad_mask(reinterpret_cast<const long long*>(m->tar), m->tar_stride, m->tar_offset, w.kpad_tar, B, T, st));

  // ---------------- encoder (Encoder.py:48-60) ----------------
  RC(bx::embed_fwd(reinterpret_cast<const long long*>(m->inp), S, 0, m->P + lo.enc_emb, m->pos_in, nullptr, w.enc_x16[0], Ms, S, d,
                   lo.Vin, rate, m->seed, site_enc_embed(), st, m->seed_dev));
  const MaskSpec em = enc_self_mask(m);
  for (int l = 0; l < L; ++l) {
    const EncLayerP& p = lo.enc[l];
    EncAct& a = w.enc[l];
    // MultiHeadAttention (Attention.py:59-78): packed QKV projection -> flash attention -> output projection
    RC(linear_fwd(w.enc_x16[l], d, Sh + p.mha.Wqkv, 3 * d, m->P + p.mha.bqkv, a.qkv, 3 * d, false, false, Ms, 3 * d, d, st));
    RC(bx::attn_fwd(a.qkv, 3 * d, a.qkv + d, 3 * d, a.qkv + 2 * d, 3 * d, a.attn_o, d, a.lse, em.kpad, em.causal, em.maskf, em.sb,
                    em.sq, B, H, S, S, dk, st, m->training ? a.attn_o_lo : nullptr));
    RC(dense_residual_ln(m, a.attn_o, d, d, p.mha.Wo, p.mha.bo, w.enc_x16[l], p.ln1g, p.ln1b, a.out1_16, a.ln1, Ms, site_enc(l, 0), st));   // Encoder.py:22-23
    // point_wise_feed_forward_network (point_ffn.py:3-7)
    bx::prof_set_gemm_cat(3);
    RC(linear_fwd(a.out1_16, d, Sh + p.ffn.W1, f, m->P + p.ffn.b1, a.h, f, false, true, Ms, f, d, st, a.hbits));
    RC(dense_residual_ln(m, a.h, f, f, p.ffn.W2, p.ffn.b2, a.out1_16, p.ln2g, p.ln2b, w.enc_x16[l + 1], a.ln2, Ms, site_enc(l, 1), st));   // Encoder.py:26-27
    bx::prof_set_gemm_cat(0);
  }
  const bf16* enc_out16 = w.enc_x16[L];

  // ---------------- decoder (Decoder.py:59-79) ----------------
  // the cross-attention K/V projections of ALL decoder layers depend only on the encoder output: they run on the side
  // stream while the main stream works through the decoder (each layer's mha2 waits for its own projection)
  Par par(m, st);
  RC(par.fork());
  std::vector<cudaEvent_t> kv_ready(L, nullptr);
  for (int l = 0; l < L; ++l) {
    const DecLayerP& p = lo.dec[l];
    RC(linear_fwd(enc_out16, d, Sh + p.mha2.Wkv, 2 * d, m->P + p.mha2.bkv, w.dec[l].kv2, 2 * d, false, false, Ms, 2 * d, d, par.side));
    kv_ready[l] = par.mark();
  }
  RC(bx::embed_fwd(reinterpret_cast<const long long*>(m->tar), m->tar_stride, m->tar_offset, m->P + lo.dec_emb, m->pos_tgt, nullptr,
                   w.dec_x16[0], Mt, T, d, lo.Vtgt, rate, m->seed, site_dec_embed(), st, m->seed_dev));
  const MaskSpec sm = dec_self_mask(m), cm = dec_cross_mask(m);
  for (int l = 0; l < L; ++l) {
    const DecLayerP& p = lo.dec[l];
    DecAct& a = w.dec[l];
    // mha1: look-ahead-masked self attention (Decoder.py:29-31)
    RC(linear_fwd(w.dec_x16[l], d, Sh + p.mha1.Wqkv, 3 * d, m->P + p.mha1.bqkv, a.qkv1, 3 * d, false, false, Mt, 3 * d, d, st));
    RC(bx::attn_fwd(a.qkv1, 3 * d, a.qkv1 + d, 3 * d, a.qkv1 + 2 * d, 3 * d, a.attn1_o, d, a.lse1, sm.kpad, sm.causal, sm.maskf, sm.sb,
                    sm.sq, B, H, T, T, dk, st, m->training ? a.attn1_o_lo : nullptr));
    RC(dense_residual_ln(m, a.attn1_o, d, d, p.mha1.Wo, p.mha1.bo, w.dec_x16[l], p.ln1g, p.ln1b, a.out1_16, a.ln1, Mt, site_dec(l, 0), st));
    // mha2: q from the decoder stream, k = v = enc_output, encoder padding mask (Decoder.py:33-36)
    RC(linear_fwd(a.out1_16, d, Sh + p.mha2.Wq, d, m->P + p.mha2.bq, a.q2, d, false, false, Mt, d, d, st));
    RC(par.wait(kv_ready[l]));
    RC(bx::attn_fwd(a.q2, d, a.kv2, 2 * d, a.kv2 + d, 2 * d, a.attn2_o, d, a.lse2, cm.kpad, cm.causal, cm.maskf, cm.sb, cm.sq, B, H, T, S,
                    dk, st, m->training ? a.attn2_o_lo : nullptr));
    RC(dense_residual_ln(m, a.attn2_o, d, d, p.mha2.Wo, p.mha2.bo, a.out1_16, p.ln2g, p.ln2b, a.out2_16, a.ln2, Mt, site_dec(l, 1), st));
    // ffn (Decoder.py:38-40)
    bx::prof_set_gemm_cat(3);
    RC(linear_fwd(a.out2_16, d, Sh + p.ffn.W1, f, m->P + p.ffn.b1, a.h, f, false, true, Mt, f, d, st, a.hbits));
    RC(dense_residual_ln(m, a.h, f, f, p.ffn.W2, p.ffn.b2, a.out2_16, p.ln3g, p.ln3b, w.dec_x16[l + 1], a.ln3, Mt, site_dec(l, 2), st));
    bx::prof_set_gemm_cat(0);
  }
  RC(par.join());
  // ---------------- final_layer (Transformer.py:30) ----------------
  if (logits_out != nullptr)
    RC(linear_fwd(w.dec_x16[L], d, Sh + lo.Wf, lo.Vp, m->P + lo.bf, logits_out, ld_logits, true, false, Mt, lo.Vtgt, d, st));
  else
    RC(linear_fwd(w.dec_x16[L], d, Sh + lo.Wf, lo.Vp, m->P + lo.bf, w.logits, lo.Vp, false, false, Mt, lo.Vtgt, d, st));
  return 0;
}

// a ring slot: the buffer and the mark of its last side-stream reader
template <typename T, int N>
struct Ring {
  T* buf[N];
  cudaEvent_t busy[N];
  int n = 0;
  void init(T* const* b) { for (int i = 0; i < N; ++i) { buf[i] = b[i]; busy[i] = nullptr; } }
  // next slot, safe to overwrite on the main stream
  int acquire(Par& par, T** out) {
    const int i = n++ % N;
    RC(par.wait(busy[i]));
    busy[i] = nullptr;
    *out = buf[i];
    return 0;
  }
  void read_by_side(const T* p, cudaEvent_t e) {
    for (int i = 0; i < N; ++i) if (buf[i] == p) busy[i] = e;
  }
};

// The gradient of a stream tensor travels as a pair: `r` (residual path: dz of the LayerNorm backward that consumed it,
// or the fp32 d_enc_out) plus `dg` (dgrad output of the sublayer that consumed it, may be NULL); the next LayerNorm
// backward (or the embedding backward) sums the two in fp32.
struct StreamGrad { const void* r; bool r_f32; const bf16* dg; };

// One backward stage.  Critical path on the main stream: LayerNorm backward -> dgrad GEMMs -> attention backward -> ...;
// everything that only feeds the gradient arena (wgrad GEMMs, bias column sums) and the cross-attention K/V branch into
// d_enc_out runs on the side stream, so their CTAs fill the tail waves of the main stream's GEMMs and the HBM-bound kernels
// of the main stream overlap tensor-bound GEMMs of the side stream.
struct BwdCtx {
  M* m;
  Par par;
  Ring<bf16, 3> dy;
  Ring<bf16, 2> dqkv, dkv2, dh, dq16;
  BwdCtx(M* m_, cudaStream_t st) : m(m_), par(m_, st) {
    Work& w = m->w;
    dy.init(w.dy16); dqkv.init(w.dqkv); dkv2.init(w.dkv2); dh.init(w.dh); dq16.init(w.dq16);
  }
  cudaStream_t st() const { return par.main; }

  // LayerNorm + dropout + residual backward: g (gradient of the LN output) -> dz (gradient of z = of the residual input,
  // bf16) and *dy_out (gradient of the sublayer output, bf16, in a dy ring slot).
  // `bias_off`: bias of the sublayer's last Dense -- its gradient (column sums of dy) falls out of this kernel for free
  int residual_ln_back(const StreamGrad& g, const LnSave& s, int64_t g_off, int64_t b_off, int64_t bias_off, bf16* dz, const bf16** dy_out,
                       int64_t rows, uint32_t site) {
    const float rate = m->training ? m->cfg.dropout_rate : 0.f;
    bf16* buf = nullptr;
    RC(dy.acquire(par, &buf));
    *dy_out = buf;
    return bx::residual_ln_bwd(g.r, g.r_f32 ? 1 : 0, g.dg, s.z, s.mean, s.rstd, m->P + g_off, dz, buf, m->G + g_off, m->G + b_off,
                               m->G + bias_off, s.keep, rows, m->lo.d, rate, m->seed, site, st(), m->seed_dev);
  }

  // side stream: dW += x^T . dy (+ db += colsum(dy)); x is a saved activation, dy was just produced on the main stream
  int wgrad_side(const bf16* x, int64_t ldx, const bf16* dyv, int64_t ld_dy, int64_t w_off, int64_t ldw, int64_t b_off, int64_t rows, int64_t N,
                 int64_t K) {
    RC(par.side_after_main());
    return linear_wgrad(x, ldx, dyv, ld_dy, m->G + w_off, ldw, b_off >= 0 ? m->G + b_off : nullptr, rows, N, K, par.side);
  }

  // FFN backward: dyv = d(ffn output); x16 = ffn input; writes d(input) (sublayer path only) to w.gD
  int ffn_back(const Ffn& p, const bf16* x16, const bf16* h, const uint64_t* hbits, const bf16* dyv, int64_t rows) {
    const int d = m->lo.d, f = m->lo.f;
    Work& w = m->w;
    struct CatGuard { CatGuard() { bx::prof_set_gemm_cat(3); } ~CatGuard() { bx::prof_set_gemm_cat(0); } } cat_guard;
    RC(wgrad_side(h, f, dyv, d, p.W2, d, -1, rows, d, f));                                     // db2: from the LN backward
    dy.read_by_side(dyv, par.mark());
    bf16* dhb = nullptr;
    RC(dh.acquire(par, &dhb));
    static const bool fuse_db = [] { const char* e = getenv("B200XF_FUSE_COLSUM"); return !(e && e[0] == '0'); }();
    RC(linear_dgrad(dyv, d, m->Sh + p.W2, d, dhb, f, false, false, h, f, rows, d, f, st(), hbits, fuse_db ? m->G + p.b1 : nullptr));   // dh = (dy . W2^T) o [h > 0]; db1 += colsum(dh)
    RC(wgrad_side(x16, d, dhb, f, p.W1, f, fuse_db ? -1 : p.b1, rows, f, d));                  // db1: from the GEMM above
    dh.read_by_side(dhb, par.mark());
    RC(linear_dgrad(dhb, f, m->Sh + p.W1, f, w.gD, d, false, false, nullptr, 0, rows, f, d, st()));
    return 0;
  }

  // self-attention block backward (packed QKV).  dyv = d(out-projection output); writes d(x) (sublayer path) to w.gD
  int self_attn_back(const MhaSelf& p, const bf16* x16, const bf16* qkv, const bf16* attn_o, const bf16* attn_o_lo, const float* lse,
                     const MaskSpec& mk, const bf16* dyv, int S) {
    const int d = m->lo.d, H = m->lo.H, dk = d / H, B = m->B;
    const int64_t rows = (int64_t)B * S;
    Work& w = m->w;
    RC(wgrad_side(attn_o, d, dyv, d, p.Wo, d, -1, rows, d, d));                                 // dbo: from the LN backward
    dy.read_by_side(dyv, par.mark());
    RC(linear_dgrad(dyv, d, m->Sh + p.Wo, d, w.dO16, d, false, false, nullptr, 0, rows, d, d, st()));
    RC(bx::attn_delta(attn_o, d, w.dO16, d, w.delta, B, S, H, dk, st(), m->training ? attn_o_lo : nullptr));
    bf16* g = nullptr;
    RC(dqkv.acquire(par, &g));
    // dQ is accumulated (bf16 red.global adds) straight into the first d columns of the packed QKV gradient: zero them first
    BX_CHECK_CUDA(cudaMemset2DAsync(g, (size_t)3 * d * sizeof(bf16), 0, (size_t)d * sizeof(bf16), rows, st()));
    RC(bx::attn_bwd(qkv, 3 * d, qkv + d, 3 * d, qkv + 2 * d, 3 * d, w.dO16, d, lse, w.delta, g, 3 * d, 1, g + d, 3 * d, g + 2 * d, 3 * d,
                    mk.kpad, mk.causal, mk.maskf, mk.sb, mk.sq, B, H, S, S, dk, st()));
    RC(wgrad_side(x16, d, g, 3 * d, p.Wqkv, 3 * d, p.bqkv, rows, 3 * d, d));
    dqkv.read_by_side(g, par.mark());
    RC(linear_dgrad(g, 3 * d, m->Sh + p.Wqkv, 3 * d, w.gD, d, false, false, nullptr, 0, rows, 3 * d, d, st()));
    return 0;
  }
};

int backward_final(M* m, cudaStream_t st) {
  const Layout& lo = m->lo;
  Work& w = m->w;
  const int d = lo.d, L = lo.L;
  const int64_t Mt = (int64_t)m->B * m->T;
  BwdCtx c(m, st);
  RC(c.par.fork());
  // dlogits (bf16, in w.logits) -> final_layer grads (side stream) and d(dec_output) in gD (main stream)
  RC(c.wgrad_side(w.dec_x16[L], d, w.logits, lo.Vp, lo.Wf, lo.Vp, lo.bf, Mt, lo.Vtgt, d));
  RC(linear_dgrad(w.logits, lo.Vp, m->Sh + lo.Wf, lo.Vp, w.gD, d, false, false, nullptr, 0, Mt, lo.Vtgt, d, st));
  RC(c.par.join());
  return 0;
}

// Backward in PARTS, each finishing one contiguous range of the gradient arena (data-parallel buckets: the range is
// all-reduced while the next part runs): the decoder layers l_hi..l_lo (part of the chain that starts at the vocabulary
// dgrad's output in gD), the decoder embedding, the encoder layers, the encoder embedding.
int backward_decoder_layers(M* m, int l_hi, int l_lo, cudaStream_t st) {
  const Layout& lo = m->lo;
  Work& w = m->w;
  const int d = lo.d, H = lo.H, L = lo.L, dk = d / H;
  const int B = m->B, S = m->S, T = m->T;
  const int64_t Ms = (int64_t)B * S, Mt = (int64_t)B * T;
  const MaskSpec sm = dec_self_mask(m), cm = dec_cross_mask(m);
  if (l_hi == L - 1) {
    BX_CHECK_CUDA(cudaMemsetAsync(w.d_enc_out, 0, Ms * d * sizeof(float), st));
    m->bw_r = w.gD; m->bw_r_f32 = false; m->bw_dg = nullptr; m->bw_dz = 0;     // the vocabulary dgrad wrote the stream gradient to gD
  }
  BwdCtx c(m, st);
  RC(c.par.fork());
  StreamGrad g{m->bw_r, m->bw_r_f32, reinterpret_cast<const bf16*>(m->bw_dg)};
  bf16* dz = m->bw_dz == 0 ? w.gR0 : w.gR1;
  bf16* dz_other = m->bw_dz == 0 ? w.gR1 : w.gR0;
  const bf16* dy = nullptr;
  auto advance = [&]() { g = StreamGrad{dz, false, w.gD}; std::swap(dz, dz_other); m->bw_dz ^= 1; };
  for (int l = l_hi; l >= l_lo; --l) {
    const DecLayerP& p = lo.dec[l];
    DecAct& a = w.dec[l];
    // out3 = LN3(out2 + drop(ffn(out2)))
    RC(c.residual_ln_back(g, a.ln3, p.ln3g, p.ln3b, p.ffn.b2, dz, &dy, Mt, site_dec(l, 2)));
    RC(c.ffn_back(p.ffn, a.out2_16, a.h, a.hbits, dy, Mt));
    advance();
    // out2 = LN2(out1 + drop(mha2(enc_out, enc_out, out1)))
    RC(c.residual_ln_back(g, a.ln2, p.ln2g, p.ln2b, p.mha2.bo, dz, &dy, Mt, site_dec(l, 1)));
    RC(c.wgrad_side(a.attn2_o, d, dy, d, p.mha2.Wo, d, -1, Mt, d, d));
    c.dy.read_by_side(dy, c.par.mark());
    RC(linear_dgrad(dy, d, m->Sh + p.mha2.Wo, d, w.dO16, d, false, false, nullptr, 0, Mt, d, d, st));
    RC(bx::attn_delta(a.attn2_o, d, w.dO16, d, w.delta, B, T, H, dk, st, m->training ? a.attn2_o_lo : nullptr));
    bf16 *dq = nullptr, *dkv = nullptr;
    RC(c.dq16.acquire(c.par, &dq));
    RC(c.dkv2.acquire(c.par, &dkv));
    BX_CHECK_CUDA(cudaMemsetAsync(dq, 0, Mt * d * sizeof(bf16), st));
    // dK/dV of the encoder memory go to dkv [Ms, 2d]; dQ (bf16 red.global adds) to dq
    RC(bx::attn_bwd(a.q2, d, a.kv2, 2 * d, a.kv2 + d, 2 * d, w.dO16, d, a.lse2, w.delta, dq, d, 1, dkv, 2 * d, dkv + d, 2 * d,
                    cm.kpad, cm.causal, cm.maskf, cm.sb, cm.sq, B, H, T, S, dk, st));
    // the whole K/V branch (weight gradient and the fp32 accumulation into d_enc_out) is off the decoder's critical path
    RC(c.wgrad_side(w.enc_x16[L], d, dkv, 2 * d, p.mha2.Wkv, 2 * d, p.mha2.bkv, Ms, 2 * d, d));
    RC(linear_dgrad(dkv, 2 * d, m->Sh + p.mha2.Wkv, 2 * d, w.d_enc_out, d, true, true, nullptr, 0, Ms, 2 * d, d, c.par.side));
    RC(c.wgrad_side(a.out1_16, d, dq, d, p.mha2.Wq, d, p.mha2.bq, Mt, d, d));
    {
      cudaEvent_t e = c.par.mark();
      c.dkv2.read_by_side(dkv, e);
      c.dq16.read_by_side(dq, e);
    }
    RC(linear_dgrad(dq, d, m->Sh + p.mha2.Wq, d, w.gD, d, false, false, nullptr, 0, Mt, d, d, st));
    advance();
    // out1 = LN1(x + drop(mha1(x, x, x)))
    RC(c.residual_ln_back(g, a.ln1, p.ln1g, p.ln1b, p.mha1.bo, dz, &dy, Mt, site_dec(l, 0)));
    RC(c.self_attn_back(p.mha1, w.dec_x16[l], a.qkv1, a.attn1_o, a.attn1_o_lo, a.lse1, sm, dy, T));
    advance();
  }
  m->bw_r = g.r; m->bw_r_f32 = g.r_f32; m->bw_dg = g.dg;
  RC(c.par.join());
  return 0;
}

int backward_decoder_embed(M* m, cudaStream_t st) {
  const Layout& lo = m->lo;
  const int d = lo.d, T = m->T;
  const int64_t Mt = (int64_t)m->B * T;
  const float rate = m->training ? m->cfg.dropout_rate : 0.f;
  return bx::embed_bwd(reinterpret_cast<const long long*>(m->tar), m->tar_stride, m->tar_offset, m->bw_r, m->bw_dg, m->G + lo.dec_emb, Mt, T, d,
                       lo.Vtgt, rate, m->seed, site_dec_embed(), st, m->seed_dev);
}

int backward_encoder_layers(M* m, int l_hi, int l_lo, cudaStream_t st) {
  const Layout& lo = m->lo;
  Work& w = m->w;
  const int L = lo.L;
  const int B = m->B, S = m->S;
  const int64_t Ms = (int64_t)B * S;
  const MaskSpec em = enc_self_mask(m);
  if (l_hi == L - 1) {      // fp32 sum of the cross-attention K/V gradients of all decoder layers
    m->bw_r = w.d_enc_out; m->bw_r_f32 = true; m->bw_dg = nullptr; m->bw_dz = 0;
  }
  BwdCtx c(m, st);
  RC(c.par.fork());
  StreamGrad g{m->bw_r, m->bw_r_f32, reinterpret_cast<const bf16*>(m->bw_dg)};
  bf16* dz = m->bw_dz == 0 ? w.gR0 : w.gR1;
  bf16* dz_other = m->bw_dz == 0 ? w.gR1 : w.gR0;
  const bf16* dy = nullptr;
  auto advance = [&]() { g = StreamGrad{dz, false, w.gD}; std::swap(dz, dz_other); m->bw_dz ^= 1; };
  for (int l = l_hi; l >= l_lo; --l) {
    const EncLayerP& p = lo.enc[l];
    EncAct& a = w.enc[l];
    RC(c.residual_ln_back(g, a.ln2, p.ln2g, p.ln2b, p.ffn.b2, dz, &dy, Ms, site_enc(l, 1)));
    RC(c.ffn_back(p.ffn, a.out1_16, a.h, a.hbits, dy, Ms));
    advance();
    RC(c.residual_ln_back(g, a.ln1, p.ln1g, p.ln1b, p.mha.bo, dz, &dy, Ms, site_enc(l, 0)));
    RC(c.self_attn_back(p.mha, w.enc_x16[l], a.qkv, a.attn_o, a.attn_o_lo, a.lse, em, dy, S));
    advance();
  }
  m->bw_r = g.r; m->bw_r_f32 = g.r_f32; m->bw_dg = g.dg;
  RC(c.par.join());
  return 0;
}

int backward_encoder_embed(M* m, cudaStream_t st) {
  const Layout& lo = m->lo;
  const int d = lo.d, S = m->S;
  const int64_t Ms = (int64_t)m->B * S;
  const float rate = m->training ? m->cfg.dropout_rate : 0.f;
  return bx::embed_bwd(reinterpret_cast<const long long*>(m->inp), S, 0, m->bw_r, m->bw_dg, m->G + lo.enc_emb, Ms, S, d, lo.Vin, rate, m->seed,
                       site_enc_embed(), st, m->seed_dev);
}

int backward_decoder(M* m, cudaStream_t st) {
  RC(backward_decoder_layers(m, m->lo.L - 1, 0, st));
  return backward_decoder_embed(m, st);
}
int backward_encoder(M* m, cudaStream_t st) {
  RC(backward_encoder_layers(m, m->lo.L - 1, 0, st));
  return backward_encoder_embed(m, st);
}

// ---- incremental (KV-cached) decoding: train.py:91-121 re-runs encoder and decoder for every generated token; here the encoder and
// the cross-attention K/V projections run once (decode_begin) and each step processes ONE new target position per sequence:
// its Q/K/V row is written straight into the self-attention cache (the layer's qkv1 buffer viewed as [B, max_tgt_len, 3d]: the
// GEMM's output pitch is the cache's batch pitch), attention reads the cached keys 0..pos through batch-strided TMA maps.
int decode_begin(M* m, cudaStream_t st) {
  const Layout& lo = m->lo;
  Work& w = m->w;
  const int d = lo.d, H = lo.H, f = lo.f, L = lo.L, dk = d / H;
  const int B = m->B, S = m->S;
  const int64_t Ms = (int64_t)B * S;
  const bf16* Sh = m->Sh;
  RC(bx::pad_mask(reinterpret_cast<const long long*>(m->inp), S, 0, w.kpad_in, B, S, st));
  RC(bx::embed_fwd(reinterpret_cast<const long long*>(m->inp), S, 0, m->P + lo.enc_emb, m->pos_in, nullptr, w.enc_x16[0], Ms, S, d, lo.Vin, 0.f, 0, 0, st));
  for (int l = 0; l < L; ++l) {
    const EncLayerP& p = lo.enc[l];
    EncAct& a = w.enc[l];
    RC(linear_fwd(w.enc_x16[l], d, Sh + p.mha.Wqkv, 3 * d, m->P + p.mha.bqkv, a.qkv, 3 * d, false, false, Ms, 3 * d, d, st));
    RC(bx::attn_fwd(a.qkv, 3 * d, a.qkv + d, 3 * d, a.qkv + 2 * d, 3 * d, a.attn_o, d, a.lse, w.kpad_in, 0, nullptr, 0, 0, B, H, S, S, dk, st));
    RC(dense_residual_ln(m, a.attn_o, d, d, p.mha.Wo, p.mha.bo, w.enc_x16[l], p.ln1g, p.ln1b, a.out1_16, a.ln1, Ms, 0, st));
    RC(linear_fwd(a.out1_16, d, Sh + p.ffn.W1, f, m->P + p.ffn.b1, a.h, f, false, true, Ms, f, d, st));
    RC(dense_residual_ln(m, a.h, f, f, p.ffn.W2, p.ffn.b2, a.out1_16, p.ln2g, p.ln2b, w.enc_x16[l + 1], a.ln2, Ms, 0, st));
  }
  for (int l = 0; l < L; ++l) {
    const DecLayerP& p = lo.dec[l];
    RC(linear_fwd(w.enc_x16[L], d, Sh + p.mha2.Wkv, 2 * d, m->P + p.mha2.bkv, w.dec[l].kv2, 2 * d, false, false, Ms, 2 * d, d, st));
  }
  return 0;
}

int decode_step(M* m, const long long* tokens, int pos, float* logits_out, int64_t ld_logits, cudaStream_t st) {
  const Layout& lo = m->lo;
  Work& w = m->w;
  const int d = lo.d, H = lo.H, f = lo.f, L = lo.L, dk = d / H;
  const int B = m->dec_B, S = m->dec_S;
  const int64_t Tc = m->cfg.max_tgt_len;              // cache pitch (positions per sequence)
  const bf16* Sh = m->Sh;
  // token history -> key-padding bytes of the cached target positions (Decoder.py:29-31: look-ahead mask OR target padding mask;
  // every cached key is <= the query position, so only the padding part is left)
  BX_CHECK_CUDA(cudaMemcpy2DAsync(w.dec_hist + pos, (size_t)Tc * sizeof(long long), tokens, sizeof(long long), sizeof(long long), B,
                                  cudaMemcpyDeviceToDevice, st));
  RC(bx::pad_mask(w.dec_hist, (int)Tc, 0, w.kpad_tar, B, pos + 1, st));
  // embedding of the new token at position `pos` (Decoder.py:65-69)
  RC(bx::embed_fwd(tokens, 1, 0, m->P + lo.dec_emb, m->pos_tgt + (int64_t)pos * d, nullptr, w.dec_x16[0], B, 1, d, lo.Vtgt, 0.f, 0, 0, st));
  for (int l = 0; l < L; ++l) {
    const DecLayerP& p = lo.dec[l];
    DecAct& a = w.dec[l];
    bf16* cache = a.qkv1;                              // [B, Tc, 3d]
    RC(linear_fwd(w.dec_x16[l], d, Sh + p.mha1.Wqkv, 3 * d, m->P + p.mha1.bqkv, cache + (int64_t)pos * 3 * d, Tc * 3 * d, false, false, B, 3 * d, d, st));
    RC(bx::attn_fwd(cache + (int64_t)pos * 3 * d, 3 * d, cache + d, 3 * d, cache + 2 * d, 3 * d, a.attn1_o, d, w.dec_lse, w.kpad_tar, 0, nullptr, 0, 0,
                    B, H, 1, pos + 1, dk, st, nullptr, Tc * 3 * d, Tc * 3 * d));
    RC(dense_residual_ln(m, a.attn1_o, d, d, p.mha1.Wo, p.mha1.bo, w.dec_x16[l], p.ln1g, p.ln1b, a.out1_16, a.ln1, B, 0, st));
    RC(linear_fwd(a.out1_16, d, Sh + p.mha2.Wq, d, m->P + p.mha2.bq, a.q2, d, false, false, B, d, d, st));
    RC(bx::attn_fwd(a.q2, d, a.kv2, 2 * d, a.kv2 + d, 2 * d, a.attn2_o, d, w.dec_lse, w.kpad_in, 0, nullptr, 0, 0, B, H, 1, S, dk, st));
    RC(dense_residual_ln(m, a.attn2_o, d, d, p.mha2.Wo, p.mha2.bo, a.out1_16, p.ln2g, p.ln2b, a.out2_16, a.ln2, B, 0, st));
    RC(linear_fwd(a.out2_16, d, Sh + p.ffn.W1, f, m->P + p.ffn.b1, a.h, f, false, true, B, f, d, st));
    RC(dense_residual_ln(m, a.h, f, f, p.ffn.W2, p.ffn.b2, a.out2_16, p.ln3g, p.ln3b, w.dec_x16[l + 1], a.ln3, B, 0, st));
  }
  return linear_fwd(w.dec_x16[L], d, Sh + lo.Wf, lo.Vp, m->P + lo.bf, logits_out, ld_logits, true, false, B, lo.Vtgt, d, st);
}

}  // namespace

// ====================================================================================================
// C ABI
// ====================================================================================================
extern "C" {

int64_t b200xf_model_param_count(const b200xf_model_config* cfg) {
  if (validate(*cfg)) return -1;
  return make_layout(*cfg).total;
}
int64_t b200xf_model_workspace_bytes(const b200xf_model_config* cfg) {
  if (validate(*cfg)) return -1;
  Layout lo = make_layout(*cfg);
  return make_work(*cfg, lo, nullptr).bytes;
}
int b200xf_model_variable(const b200xf_model_config* cfg, int index, char* name, int name_cap, int64_t* offset, int64_t* rows,
                          int64_t* cols, int64_t* ld) {
  RC(validate(*cfg));
  Layout lo = make_layout(*cfg);
  if (index < 0 || index >= (int)lo.vars.size()) return 1;
  const VarView& v = lo.vars[index];
  if (name && name_cap > 0) { strncpy(name, v.name.c_str(), name_cap - 1); name[name_cap - 1] = 0; }
  *offset = v.off; *rows = v.rows; *cols = v.cols; *ld = v.ld;
  return 0;
}
int b200xf_model_grad_range(const b200xf_model_config* cfg, int stage, int64_t* begin, int64_t* end) {
  RC(validate(*cfg));
  BX_REQUIRE(stage >= 0 && stage < 3, "grad_range: stage must be 0..2");
  Layout lo = make_layout(*cfg);
  *begin = lo.stage_begin[stage]; *end = lo.stage_end[stage];
  return 0;
}
int b200xf_model_create(const b200xf_model_config* cfg, float* params, float* grads, float* adam_m, float* adam_v, void* shadow,
                        const float* pos_in, const float* pos_tgt, void* workspace, b200xf_model** out) {
  RC(validate(*cfg));
  BX_REQUIRE(params && grads && shadow && pos_in && pos_tgt && workspace && out, "model_create: null buffer");
  BX_REQUIRE((reinterpret_cast<uintptr_t>(params) & 255) == 0 && (reinterpret_cast<uintptr_t>(grads) & 255) == 0 &&
                 (reinterpret_cast<uintptr_t>(shadow) & 255) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 1023) == 0,
             "model_create: arenas must be 256-byte aligned, workspace 1024-byte aligned");
  b200xf_model* m = new b200xf_model();
  m->cfg = *cfg;
  m->lo = make_layout(*cfg);
  m->w = make_work(*cfg, m->lo, workspace);
  m->P = params; m->G = grads; m->Am = adam_m; m->Av = adam_v; m->Sh = reinterpret_cast<bf16*>(shadow);
  m->pos_in = pos_in; m->pos_tgt = pos_tgt;
  // second stream + event pool (host-side objects of the CURRENT device; created here so that nothing is created while a
  // CUDA-graph capture is in progress)
  // measured on B200 (profiles/r02_*bench_conc*.json): the forked wgrad / K-V branches do NOT pay at transformer-base shapes -- every
  // GEMM is a persistent kernel that fills all SMs (one CTA per SM, ~210 KB of shared memory), so a second stream's CTAs only ever
  // enter at the tail of a wave, and the HBM-bound kernels cannot co-reside with a GEMM CTA; the cross-stream dependencies cost
  // more than the tails return (22.1 vs 21.8 ms/step).  Off by default; B200XF_CONCURRENCY=1 or b200xf_model_set_concurrency.
  const char* env = getenv("B200XF_CONCURRENCY");
  m->concurrency = (env != nullptr && env[0] == '1') ? 1 : 0;
  env = getenv("B200XF_FUSE_LN");
  m->fuse_ln = (env != nullptr && env[0] == '0') ? 0 : 1;
  if (cudaStreamCreateWithFlags(&m->side, cudaStreamNonBlocking) != cudaSuccess) { m->side = nullptr; m->concurrency = 0; (void)cudaGetLastError(); }
  m->events.resize(1024, nullptr);
  for (auto& e : m->events)
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { m->concurrency = 0; (void)cudaGetLastError(); break; }
  *out = m;
  return 0;
}
void b200xf_model_destroy(b200xf_model* m) {
  if (m == nullptr) return;
  for (auto e : m->events) if (e != nullptr) cudaEventDestroy(e);
  if (m->side != nullptr) cudaStreamDestroy(m->side);
  delete m;
}
void b200xf_model_set_concurrency(b200xf_model* m, int on) { m->concurrency = (on != 0 && m->side != nullptr) ? 1 : 0; }

void b200xf_model_set_seed_ptr(b200xf_model* m, const uint64_t* seed_dev) {
  m->seed_dev = reinterpret_cast<const unsigned long long*>(seed_dev);
}

int b200xf_model_sync_shadow(b200xf_model* m, b200xf_stream_t stream) {
  return bx::cast_f32_bf16(m->P, m->Sh, m->lo.total, reinterpret_cast<cudaStream_t>(stream));
}

int b200xf_model_forward(b200xf_model* m, const int64_t* inp, const int64_t* tar, int tar_stride, int tar_offset, int B, int S, int T,
                         int training, uint64_t seed, const float* enc_mask, const float* look_ahead_mask, int64_t la_batch_stride,
                         const float* dec_mask, float* logits_out, int64_t ld_logits, b200xf_stream_t stream) {
  BX_REQUIRE(B >= 1 && B <= m->cfg.max_batch && S >= 1 && S <= m->cfg.max_src_len && T >= 1 && T <= m->cfg.max_tgt_len,
             "forward: shape (B %d, S %d, T %d) exceeds the workspace (%d, %d, %d)", B, S, T, m->cfg.max_batch, m->cfg.max_src_len,
             m->cfg.max_tgt_len);
  BX_REQUIRE(tar_stride >= tar_offset + T, "forward: target stride/offset inconsistent");
  BX_REQUIRE(logits_out == nullptr || ld_logits % 4 == 0, "forward: logits leading dimension must be a multiple of 4");
  m->B = B; m->S = S; m->T = T; m->training = training; m->seed = seed;
  m->inp = inp; m->tar = tar; m->tar_stride = tar_stride; m->tar_offset = tar_offset;
  m->enc_mask = enc_mask; m->la_mask = look_ahead_mask; m->la_bs = la_batch_stride; m->dec_mask = dec_mask;
  m->have_loss_grad = false;
  m->logits_valid = logits_out == nullptr;
  return forward_impl(m, logits_out, ld_logits, reinterpret_cast<cudaStream_t>(stream));
}

int b200xf_model_loss(b200xf_model* m, const int64_t* tar, int tar_stride, int label_offset, float inv_batch, double* stats, int with_grad,
                      b200xf_stream_t stream) {
  BX_REQUIRE(m->B > 0 && m->logits_valid, "loss: needs a preceding forward with logits_out == NULL (and no loss-with-grad since)");
  const int64_t Mt = (int64_t)m->B * m->T;
  RC(bx::xent(m->w.logits, 0, m->lo.Vp, reinterpret_cast<const long long*>(tar), tar_stride, label_offset, m->T,
              with_grad ? m->w.logits : nullptr, m->lo.Vp, nullptr, stats, Mt, m->lo.Vtgt, inv_batch, reinterpret_cast<cudaStream_t>(stream),
              m->t_valid));
  m->have_loss_grad = with_grad != 0;
  if (with_grad) m->logits_valid = false;
  return 0;
}

int b200xf_model_backward(b200xf_model* m, int stage, b200xf_stream_t stream) {
  BX_REQUIRE(m->have_loss_grad, "backward: call b200xf_model_loss(with_grad=1) after a training forward first");
  BX_REQUIRE(stage >= -1 && stage <= 2, "backward: stage must be -1..2");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (stage == -1 || stage == 0) RC(backward_final(m, st));
  if (stage == -1 || stage == 1) RC(backward_decoder(m, st));
  if (stage == -1 || stage == 2) RC(backward_encoder(m, st));
  return 0;
}

/* parts, in backward-completion order: 0 final layer | 1 decoder layer `layer` (L-1 first) | 2 decoder embedding |
 * 3 encoder layer `layer` (L-1 first) | 4 encoder embedding */
int b200xf_model_backward_part(b200xf_model* m, int part, int layer, b200xf_stream_t stream) {
  BX_REQUIRE(m->have_loss_grad, "backward: call b200xf_model_loss(with_grad=1) after a training forward first");
  BX_REQUIRE(part >= 0 && part <= 4 && layer >= 0 && layer < m->lo.L, "backward_part: bad part %d / layer %d", part, layer);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (part) {
    case 0: return backward_final(m, st);
    case 1: return backward_decoder_layers(m, layer, layer, st);
    case 2: return backward_decoder_embed(m, st);
    case 3: return backward_encoder_layers(m, layer, layer, st);
    default: return backward_encoder_embed(m, st);
  }
}
int b200xf_model_grad_range_part(const b200xf_model_config* cfg, int part, int layer, int64_t* begin, int64_t* end) {
  RC(validate(*cfg));
  Layout lo = make_layout(*cfg);
  BX_REQUIRE(part >= 0 && part <= 4 && layer >= 0 && layer < lo.L, "grad_range_part: bad part %d / layer %d", part, layer);
  std::pair<int64_t, int64_t> r;
  switch (part) {
    case 0: r = {lo.stage_begin[0], lo.stage_end[0]}; break;
    case 1: r = lo.dec_layer_range[layer]; break;
    case 2: r = lo.dec_emb_range; break;
    case 3: r = lo.enc_layer_range[layer]; break;
    default: r = lo.enc_emb_range; break;
  }
  *begin = r.first; *end = r.second;
  return 0;
}

void b200xf_model_set_valid_target_len_ptr(b200xf_model* m, const int32_t* t_valid_dev) { m->t_valid = t_valid_dev; }

int b200xf_model_decode_begin(b200xf_model* m, const int64_t* inp, int B, int S, b200xf_stream_t stream) {
  BX_REQUIRE(B >= 1 && B <= m->cfg.max_batch && S >= 1 && S <= m->cfg.max_src_len, "decode_begin: shape (B %d, S %d) exceeds the workspace", B, S);
  m->B = B; m->S = S; m->T = 1; m->training = 0; m->seed = 0; m->inp = inp; m->tar = nullptr;
  m->enc_mask = m->la_mask = m->dec_mask = nullptr;
  m->have_loss_grad = false; m->logits_valid = false;
  m->dec_B = B; m->dec_S = S; m->dec_len = 0;
  return decode_begin(m, reinterpret_cast<cudaStream_t>(stream));
}
int b200xf_model_decode_step(b200xf_model* m, const int64_t* tokens, float* logits_out, int64_t ld_logits, b200xf_stream_t stream) {
  BX_REQUIRE(m->dec_B > 0, "decode_step: call b200xf_model_decode_begin first");
  BX_REQUIRE(m->dec_len < m->cfg.max_tgt_len, "decode_step: the cache holds max_tgt_len = %d positions", m->cfg.max_tgt_len);
  BX_REQUIRE(logits_out != nullptr && ld_logits % 4 == 0 && ld_logits >= m->lo.Vtgt, "decode_step: bad logits buffer");
  const int pos = m->dec_len++;
  return decode_step(m, reinterpret_cast<const long long*>(tokens), pos, logits_out, ld_logits, reinterpret_cast<cudaStream_t>(stream));
}

int b200xf_model_attention_weights(b200xf_model* m, int layer, int block, float* P, b200xf_stream_t stream) {
  BX_REQUIRE(m->B > 0, "attention_weights: no forward has run");
  BX_REQUIRE(layer >= 0 && layer < m->lo.L && (block == 1 || block == 2), "attention_weights: bad layer/block");
  const int d = m->lo.d, H = m->lo.H, dk = d / H;
  const DecAct& a = m->w.dec[layer];
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (block == 1) {
    const MaskSpec mk = dec_self_mask(m);
    return bx::attn_probs(a.qkv1, 3 * d, a.qkv1 + d, 3 * d, a.lse1, P, mk.kpad, mk.causal, mk.maskf, mk.sb, mk.sq, m->B, H, m->T, m->T, dk, st);
  }
  const MaskSpec mk = dec_cross_mask(m);
  return bx::attn_probs(a.q2, d, a.kv2, 2 * d, a.lse2, P, mk.kpad, mk.causal, mk.maskf, mk.sb, mk.sq, m->B, H, m->T, m->S, dk, st);
}

int b200xf_model_tap(b200xf_model* m, const char* name, float* out, b200xf_stream_t stream) {
  BX_REQUIRE(m->B > 0, "tap: no forward has run");
  const int d = m->lo.d, L = m->lo.L;
  const bf16* src = nullptr;
  int64_t rows = 0;
  int idx = -1;
  if (!strcmp(name, "enc_embed")) { src = m->w.enc_x16[0]; rows = (int64_t)m->B * m->S; }
  else if (!strcmp(name, "dec_embed")) { src = m->w.dec_x16[0]; rows = (int64_t)m->B * m->T; }
  else if (sscanf(name, "enc_layer%d", &idx) == 1 && idx >= 0 && idx < L) { src = m->w.enc_x16[idx + 1]; rows = (int64_t)m->B * m->S; }
  else if (sscanf(name, "dec_layer%d", &idx) == 1 && idx >= 0 && idx < L) { src = m->w.dec_x16[idx + 1]; rows = (int64_t)m->B * m->T; }
  BX_REQUIRE(src != nullptr, "tap: unknown activation '%s'", name);
  return bx::cast_bf16_f32(src, out, rows * d, reinterpret_cast<cudaStream_t>(stream));   // the streams are bf16 in HBM
}

}  // extern "C"
