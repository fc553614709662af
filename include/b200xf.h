/* b200xf -- C ABI of the B200-native Transformer training hot path (drop-in for kuetuofa/Transformer's
 * TensorFlow-2 runtime on the path Transformer.call -> Train.train_step -> distributed_train).
 *
 * Conventions:
 *   - extern "C", plain pointers and sizes only; device pointers are `void*`/typed pointers into CUDA memory
 *     owned by the CALLER (the Python host allocates them as torch CUDA tensors); no entry point allocates or
 *     synchronises unless its name says so; every launch goes to the `stream` argument (a cudaStream_t).
 *   - return 0 on success, non-zero on error; b200xf_last_error() returns a thread-local message.
 *   - bf16 = raw 16-bit bfloat16 storage; matrices are row-major with a leading dimension in ELEMENTS.
 *
 * Each entry point names the reference interface (file:line in kuetuofa/Transformer) whose TF op(s) it replaces.
 * The reference has no FFI of its own (pure Python on TensorFlow); INTEGRATION.md shows the ctypes binding.
 */
#ifndef B200XF_H_
#define B200XF_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* b200xf_stream_t; /* cudaStream_t */

const char* b200xf_last_error(void);
int b200xf_version(void);
/* number of kernels this library has launched on this process so far (bench.py's gpu_launches) */
long long b200xf_launch_count(void);
/* optional CUDA-event timing of every tensor-core kernel launch (used by bench.py for the live roofline numbers):
 * categories 0 = GEMM, 1 = attention forward, 2 = attention backward; after synchronising the stream,
 * b200xf_prof_collect fills out[cat*3 + {0,1,2}] = {total ms, total algorithmic flops, launches} and resets. */
/* debug: CTA (0,0,0) of the attention kernels writes clock64 stamps into this device buffer (>= 128 int64; NULL = off) */
int b200xf_debug_set_timeline(long long* dev_ptr);
void b200xf_prof_enable(int on);
int b200xf_prof_collect(double* out, int ncat);

/* ---- dense contractions: keras Dense (Attention.py:62-64,77; point_ffn.py:5-6; Transformer.py:30) and their
 *      reverse-mode gradients (train.py:137).  D[M,N] (+)= epi(A[M,K] . B[K,N]).
 *      a_major / b_major: 0 = K contiguous in memory (A stored [M,K], B stored [N,K]);
 *                         1 = M resp. N contiguous   (A stored [K,M], B stored [K,N] -- the keras kernel layout).
 *      flags: 1 = ReLU, 2 = accumulate into fp32 D (TMA reduce-add; required for splits > 1),
 *             4 = multiply by (aux > 0) (ReLU backward, aux bf16 [M, ld_aux]).                                   */
/*      splits: K-split count for ACCUM GEMMs (<= 0: chosen by the library).  Tiles are 128 x BN per CTA or, for
 *      M > 128 and N > 128, 256 x 256 per CTA PAIR (tcgen05 cta_group::2); b200xf_gemm_force_cta_group(1|2)
 *      overrides the choice (0 restores it) -- used by tests and benchmarks.                                   */
void b200xf_gemm_force_cta_group(int cg);
/* GEMM grids are PERSISTENT by default (one CTA or CTA pair per SM, static tile schedule).  While NCCL collectives run on the same GPU
 * (the data-parallel backward, distributed_train.py:137-166) their CTAs occupy SMs, and a persistent CTA that finds its SM taken starts a
 * whole wave late; n > 0 launches ~n short-lived units of consecutive tiles per SM instead and lets the hardware block scheduler place
 * them on whatever SMs are free.  n = 0 restores the persistent schedule.  Process-global, takes effect for subsequent launches. */
void b200xf_gemm_set_chunks_per_unit(int n);
/* K3 / K4 fused (Attention.py:77 + Encoder.py:22-23 / Decoder.py:30-31,35-36 ; point_ffn.py:6 + Encoder.py:26-27 / Decoder.py:39-40):
 *      z = xres + dropout(A[M,K] . W[K,N] + bias)   out = LayerNorm(z) * gamma + beta        -- ONE kernel, LayerNorm in the epilogue
 *      A bf16 (row pitch lda), W bf16 keras [in,out] layout (row pitch ldw), N = d_model in {128, 256, 512} (a CTA pair keeps whole
 *      rows in tensor memory); xres / z / out bf16 [M,N] contiguous, mean / rstd fp32 [M], keep_bits uint8 [M, N/8] (written when
 *      rate > 0 and non-NULL; identical to b200xf_residual_ln_fwd's mask for the same (seed, site)).                             */
int b200xf_gemm_bias_dropout_residual_ln(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, const void* xres_bf16,
                                         const float* gamma, const float* beta, void* z_bf16, void* out_bf16, float* mean, float* rstd,
                                         void* keep_bits, int64_t M, int64_t N, int64_t K, float eps, float rate, uint64_t seed, uint32_t site,
                                         b200xf_stream_t stream);
int b200xf_gemm_bf16(const void* A, int64_t lda, int a_major, const void* B, int64_t ldb, int b_major, void* D,
                     int64_t ldd, int d_is_f32, const float* bias, const void* aux, int64_t ld_aux, int flags,
                     int splits, int64_t M, int64_t N, int64_t K, b200xf_stream_t stream);
/* Same product with a bf16 D, plus  colsum[n] += sum_m D[m,n]  (fp32 atomics, D as rounded to bf16) taken from the epilogue's staging tile.
 * This is how the backward gets a Dense bias gradient (reverse mode of `+ b`, point_ffn.py:5 under train.py:137) without re-reading
 * dy from HBM: the dgrad GEMM that PRODUCES dy for the layer below also sums its columns.  colsum is NOT zeroed here.           */
int b200xf_gemm_bf16_colsum(const void* A, int64_t lda, int a_major, const void* B, int64_t ldb, int b_major, void* D,
                            int64_t ldd, const float* bias, const void* aux, int64_t ld_aux, int flags, int64_t M, int64_t N,
                            int64_t K, float* colsum, b200xf_stream_t stream);

/* ---- scaled_dot_product_attention + split_heads/merge (Attention.py:3-34, 52-57, 74-76).
 *      Q/K/V/O: bf16 token-major [B*S, ld], head h = columns [h*dk, (h+1)*dk).  lse: fp32 [B,H,Sq].
 *      kpad: uint8 [B,Sk], 1 = key token id 0 (create_padding_mask, positionalencoding.py:25-30) or NULL;
 *      causal: look-ahead mask (positionalencoding.py:32-34); maskf: optional float mask, element (b,q,k) at
 *      b*mask_sb + q*mask_sq + k, contributes maskf * -1e9 (Attention.py:25-26).                                  */
int b200xf_attn_fwd(const void* Q, int64_t ld_q, const void* K, int64_t ld_k, const void* V, int64_t ld_v, void* O,
                    int64_t ld_o, float* lse, const uint8_t* kpad, int causal, const float* maskf, int64_t mask_sb,
                    int64_t mask_sq, int B, int H, int Sq, int Sk, int dk, b200xf_stream_t stream);
/* Training forward: additionally writes O_lo (bf16, same layout / pitch as O) = bf16(O - bf16(O)), the rounding residual of the
 * attention output.  b200xf_attn_delta_residual then forms delta = rowsum(dO o (O + O_lo)): dS = P o (dP - delta) cancels
 * catastrophically whenever the value vectors share a common component, and delta from the bf16-rounded O alone then puts
 * percent-level errors into dQ / dK (tools/attn_precision.py; reverse-mode of Attention.py:29-32). */
int b200xf_attn_fwd_residual(const void* Q, int64_t ld_q, const void* K, int64_t ld_k, const void* V, int64_t ld_v, void* O,
                             void* O_lo, int64_t ld_o, float* lse, const uint8_t* kpad, int causal, const float* maskf, int64_t mask_sb,
                             int64_t mask_sq, int B, int H, int Sq, int Sk, int dk, b200xf_stream_t stream);
int b200xf_attn_delta_residual(const void* O, const void* O_lo, int64_t ld_o, const void* dO, int64_t ld_do, float* delta, int B, int S,
                               int H, int dk, b200xf_stream_t stream);
/*      delta: fp32 [B,H,Sq] = rowsum(dO o O) (b200xf_attn_delta); dq_acc: [B*Sq, ld_dq] ZEROED by the caller, fp32
 *      (every key tile's contribution reduce-added by TMA, cp.reduce.async.bulk.tensor) or, with dq_is_bf16, bf16
 *      (added with red.global.add.v4.bf16x2 directly into the tensor the QKV weight-gradient GEMM reads);
 *      dK/dV: bf16 token-major like K/V.                                                                        */
int b200xf_attn_bwd(const void* Q, int64_t ld_q, const void* K, int64_t ld_k, const void* V, int64_t ld_v,
                    const void* dO, int64_t ld_do, const float* lse, const float* delta, void* dq_acc, int64_t ld_dq,
                    int dq_is_bf16, void* dK, int64_t ld_dk, void* dV, int64_t ld_dv, const uint8_t* kpad, int causal,
                    const float* maskf, int64_t mask_sb, int64_t mask_sq, int B, int H, int Sq, int Sk, int dk,
                    b200xf_stream_t stream);
int b200xf_attn_delta(const void* O, int64_t ld_o, const void* dO, int64_t ld_do, float* delta, int B, int S, int H,
                      int dk, b200xf_stream_t stream);
/*      attention_weights (Decoder.py:75-76): P fp32 [B,H,Sq,Sk] recomputed from Q, K and lse                     */
int b200xf_attn_probs(const void* Q, int64_t ld_q, const void* K, int64_t ld_k, const float* lse, float* P,
                      const uint8_t* kpad, int causal, const float* maskf, int64_t mask_sb, int64_t mask_sq, int B,
                      int H, int Sq, int Sk, int dk, b200xf_stream_t stream);

/* ---- Embedding * sqrt(d) + pos_encoding[:, :S] + Dropout (Encoder.py:51-55, Decoder.py:65-69) and its gradient.
 *      row r = (b, s), rows = B*S; its token id is ids[b*id_stride + id_offset + s] (tar[:, :-1] without a copy). */
int b200xf_embed_fwd(const int64_t* ids, int id_stride, int id_offset, const float* table, const float* pos,
                     float* out_f32, void* out_bf16, int64_t rows, int S, int d, int64_t V, float rate, uint64_t seed,
                     uint32_t site, b200xf_stream_t stream);
/*      out_f32 may be NULL (the training path keeps only the bf16 stream).
 *      backward: the gradient of the embedding output is g1 + g2 (bf16 [rows, d]; g2 may be NULL).                 */
int b200xf_embed_bwd(const int64_t* ids, int id_stride, int id_offset, const void* g1_bf16, const void* g2_bf16, float* dtable,
                     int64_t rows, int S, int d, int64_t V, float rate, uint64_t seed, uint32_t site, b200xf_stream_t stream);

/* ---- Dropout + residual add + LayerNormalization(eps) (Encoder.py:22-23,26-27; Decoder.py:30-31,35-36,39-40).
 *      y (sublayer output), x (residual input), z = x + dropout(y) (kept for backward) and out = LN(z) are bf16
 *      [rows, d] streams in HBM; statistics and normalisation are fp32 on the unrounded z.                          */
int b200xf_residual_ln_fwd(const void* y_bf16, const void* x_bf16, const float* gamma, const float* beta, void* z_bf16,
                           void* out_bf16, float* mean, float* rstd, void* keep_bits, int64_t rows, int d, float eps,
                           float rate, uint64_t seed, uint32_t site, b200xf_stream_t stream);
/*      keep_bits (uint8 [rows, d/8], may be NULL): the dropout keep mask, bit e of byte j = element 8j+e, written
 *      when rate > 0 so that the backward pass reads it back instead of re-running the counter-based RNG.         */
/*      backward: incoming gradient of `out` = g1 (fp32 if g1_is_f32 else bf16) + g2 (bf16, may be NULL), summed in fp32.
 *      dz (bf16) = gradient of z = of the residual input; dy_sub (bf16, may be NULL when rate == 0: equals dz) =
 *      gradient of the sublayer output; dgamma/dbeta += ; dbias (may be NULL) += column sums of dy_sub: the bias
 *      gradient of the sublayer's last Dense, for free                                                              */
int b200xf_residual_ln_bwd(const void* g1, int g1_is_f32, const void* g2_bf16, const void* z_bf16, const float* mean,
                           const float* rstd, const float* gamma, void* dz_bf16, void* dy_sub_bf16, float* dgamma,
                           float* dbeta, float* dbias, const void* keep_bits, int64_t rows, int d, float rate,
                           uint64_t seed, uint32_t site, b200xf_stream_t stream);
/*      keep_bits: the mask the forward stored, or NULL to regenerate it from (seed, site).                          */

/* ---- Train.loss_function + SparseCategoricalAccuracy (train.py:83-88, 141) and d(loss)/d(logits).
 *      row r = (b, t) reads label labels[b*label_stride + label_offset + t] (tar[:, 1:] without a copy).
 *      stats: double[4] += {sum of masked per-token losses, #argmax==label (all positions), #positions, #non-pad}.
 *      dlogits (bf16 [rows, ld_d], may alias bf16 logits, may be NULL) = (softmax - onehot) * (label!=0) * inv_batch */
int b200xf_xent(const void* logits, int logits_f32, int64_t ld, const int64_t* labels, int label_stride,
                int label_offset, int T, void* dlogits, int64_t ld_d, float* row_loss, double* stats, int64_t rows,
                int V, float inv_batch, b200xf_stream_t stream);

/* ---- bias gradients: out[n] += sum_m X[m,n] (X bf16) */
int b200xf_colsum_bf16(const void* X, int64_t ld, float* out, int64_t rows, int N, b200xf_stream_t stream);

/* ---- keras Adam, TF form (train.py:65-66,138): m,v update; p -= alpha*m/(sqrt(v)+eps); shadow = bf16(p) (may be NULL) */
int b200xf_adam_step(float* p, float* m, float* v, const float* g, void* shadow_bf16, int64_t n, float alpha,
                     float beta1, float beta2, float eps, b200xf_stream_t stream);

/*      Graph-replayable form: the per-step scalars live on the device.  b200xf_step_params (one thread) reads and
 *      increments state[0] = keras optimizer.iterations and writes seed = base_seed + iterations, lr =
 *      CustomSchedule(iterations) (train.py:21-34; const_lr >= 0 overrides) and the TF-form alpha;
 *      b200xf_adam_step_dev reads alpha from the device.                                                        */
int b200xf_step_params(int64_t* state, uint64_t base_seed, float d_model, float warmup_steps, float const_lr, float beta1,
                       float beta2, uint64_t* seed_out, float* alpha_out, float* lr_out, b200xf_stream_t stream);
int b200xf_adam_step_dev(float* p, float* m, float* v, const float* g, void* shadow_bf16, int64_t n, const float* alpha_dev,
                         float beta1, float beta2, float eps, b200xf_stream_t stream);

/* ---- casts / masks */
int b200xf_cast_f32_bf16(const float* in, void* out, int64_t n, b200xf_stream_t stream);
int b200xf_cast_bf16_f32(const void* in, float* out, int64_t n, b200xf_stream_t stream);
int b200xf_cast_f32_bf16_strided(const float* in, void* out, int64_t rows, int d, int64_t ld_out,
                                 b200xf_stream_t stream);
/*      create_padding_mask (positionalencoding.py:25-30) in compact form: mask[b*S+s] = (ids[b*id_stride+id_offset+s] == 0) */
int b200xf_pad_mask(const int64_t* ids, int id_stride, int id_offset, uint8_t* mask, int64_t rows, int S,
                    b200xf_stream_t stream);

/* ---- whole-model entry points (Transformer.call, Transformer.py:21-32; Train.train_step / test_step,
 *      train.py:124-156).  See transformer_b200/csrc/model.cu for the parameter-arena layout.                     */
typedef struct b200xf_model_config {
  int32_t num_layers, d_model, num_heads, dff;
  int32_t input_vocab_size, target_vocab_size; /* logical V (keras Embedding / final Dense sizes) */
  int32_t max_batch, max_src_len, max_tgt_len; /* workspace is sized for these */
  int32_t pos_rows;                            /* rows of the two positional-encoding tables */
  float dropout_rate;
  float ln_eps;
} b200xf_model_config;

typedef struct b200xf_model b200xf_model; /* opaque */

/* bytes of (a) one parameter arena in fp32 elements and (b) the activation workspace for `cfg` */
int64_t b200xf_model_param_count(const b200xf_model_config* cfg);
int64_t b200xf_model_workspace_bytes(const b200xf_model_config* cfg);
/* offset (in elements) and shape of trainable variable `index` in keras order; returns 0 while index is valid.
 * name: e.g. "encoder/layer0/mha/wq/kernel"; rows/cols/ld describe a (possibly strided) 2-D view, cols==0 for 1-D */
int b200xf_model_variable(const b200xf_model_config* cfg, int index, char* name, int name_cap, int64_t* offset,
                          int64_t* rows, int64_t* cols, int64_t* ld);
/* params/grads/adam_m/adam_v: fp32 arenas of b200xf_model_param_count elements; shadow: bf16 arena, same length;
 * pos_in/pos_tgt: fp32 [pos_rows, d_model]; workspace: b200xf_model_workspace_bytes bytes. */
int b200xf_model_create(const b200xf_model_config* cfg, float* params, float* grads, float* adam_m, float* adam_v,
                        void* shadow, const float* pos_in, const float* pos_tgt, void* workspace,
                        b200xf_model** out);
void b200xf_model_destroy(b200xf_model* m);
/* Inside one forward / backward call the library forks work that is off the critical path (weight-gradient GEMMs, bias
 * column sums, the cross-attention K/V projections) onto a second stream it owns and joins it back before returning, so the
 * caller still orders everything with ONE stream (and may capture that stream into a CUDA graph).  OFF by default (measured: no gain
 * at transformer-base shapes, see model.cu); on = 1 or environment B200XF_CONCURRENCY=1 at model creation enables it. */
void b200xf_model_set_concurrency(b200xf_model* m, int on);
/* when seed_dev != NULL the dropout seed of forward/backward is read from this device word (b200xf_step_params output)
 * instead of the `seed` argument -- required for CUDA-graph replays */
void b200xf_model_set_seed_ptr(b200xf_model* m, const uint64_t* seed_dev);
/* refresh the bf16 shadow from the fp32 params (after loading / initialising weights) */
int b200xf_model_sync_shadow(b200xf_model* m, b200xf_stream_t stream);

/* Forward.  inp int64 [B,S]; tar int64 [B, tar_stride] of which columns [tar_offset, tar_offset+T) are the decoder
 * input.  training != 0 enables dropout (seeded by `seed`) and keeps activations for b200xf_model_backward.
 * Optional float masks of the six-argument API (NULL = derive the canonical masks from the ids):
 *   enc_mask/dec_mask (B,1,1,S) ; look_ahead_mask (B,1,T,T) or (1,1,T,T) (la_batch_stride 0).
 * logits_out: NULL, or fp32 [B*T, ld_logits] to materialise final_layer outputs (Transformer.py:30).             */
int b200xf_model_forward(b200xf_model* m, const int64_t* inp, const int64_t* tar, int tar_stride, int tar_offset,
                         int B, int S, int T, int training, uint64_t seed, const float* enc_mask,
                         const float* look_ahead_mask, int64_t la_batch_stride, const float* dec_mask,
                         float* logits_out, int64_t ld_logits, b200xf_stream_t stream);
/* loss/accuracy of the last forward against labels tar[:, label_offset : label_offset+T] (train.py:83-88,141);
 * stats: device double[4] (see b200xf_xent), accumulated.  with_grad != 0 also leaves dlogits for backward.       */
int b200xf_model_loss(b200xf_model* m, const int64_t* tar, int tar_stride, int label_offset, float inv_batch,
                      double* stats, int with_grad, b200xf_stream_t stream);
/* Backward of the last training forward+loss into the (caller-zeroed) grads arena.  stage: -1 = everything;
 * otherwise 0 = final layer, 1 = decoder stack + target embedding, 2 = encoder stack + input embedding, so a
 * data-parallel caller can all-reduce finished arena ranges while later stages run (distributed_train.py:59). */
int b200xf_model_backward(b200xf_model* m, int stage, b200xf_stream_t stream);
/* [begin,end) element range of the gradient arena that `stage` finishes */
int b200xf_model_grad_range(const b200xf_model_config* cfg, int stage, int64_t* begin, int64_t* end);
/* The same backward in finer PARTS, each finishing one contiguous range of the gradient arena -- the data-parallel gradient buckets
 * (distributed_train.py:137-166: one SUM all-reduce per bucket, overlapped with the parts that follow).  Order of completion:
 *   part 0 final layer | part 1 decoder layer `layer` (L-1 down to 0) | part 2 decoder embedding |
 *   part 3 encoder layer `layer` (L-1 down to 0) | part 4 encoder embedding          (`layer` is ignored by parts 0, 2, 4) */
int b200xf_model_backward_part(b200xf_model* m, int part, int layer, b200xf_stream_t stream);
int b200xf_model_grad_range_part(const b200xf_model_config* cfg, int part, int layer, int64_t* begin, int64_t* end);
/* Shape bucketing: the host may pad a ragged batch (utils.py:154-159 padded_batch) up to a bucket length with pad id 0 so that one
 * captured CUDA graph serves a range of lengths.  Padded keys are masked and padded target positions carry label 0 (no loss, no
 * gradient) by the reference's own rules; the one thing that would change is the accuracy metric, which counts ALL positions
 * (train.py:141): positions t >= *t_valid_dev of b200xf_model_loss are therefore left out of the counters.  The length lives on
 * the device (a captured graph of the bucket shape is replayed for every real length); NULL = no filler. */
void b200xf_model_set_valid_target_len_ptr(b200xf_model* m, const int32_t* t_valid_dev);
/* Greedy inference with a key/value cache (train.py:91-121 recomputes encoder and decoder for every generated token):
 * decode_begin runs the encoder (training = 0) on inp int64 [B,S] and the cross-attention K/V projections of every decoder layer;
 * each decode_step feeds ONE new decoder-input token per sequence (tokens int64 [B], device), appends its self-attention K/V to
 * the cache and returns the logits of that position (logits_out fp32 [B, ld_logits]).  At most max_tgt_len steps. */
int b200xf_model_decode_begin(b200xf_model* m, const int64_t* inp, int B, int S, b200xf_stream_t stream);
int b200xf_model_decode_step(b200xf_model* m, const int64_t* tokens, float* logits_out, int64_t ld_logits, b200xf_stream_t stream);
/* attention_weights['decoder_layer{layer+1}_block{block}'] of the last forward -> P fp32 [B,H,T,Sk] */
int b200xf_model_attention_weights(b200xf_model* m, int layer, int block, float* P, b200xf_stream_t stream);
/* debug/parity taps: copy an internal activation of the last forward (fp32 [rows, d_model]) -- names:
 * "enc_embed","enc_layer{i}","dec_embed","dec_layer{i}" */
int b200xf_model_tap(b200xf_model* m, const char* name, float* out, b200xf_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B200XF_H_ */
