// extern "C" surface of libb200xf.so (declared in include/b200xf.h): thin forwards to the bx:: launchers.
#include "../../include/b200xf.h"

#include "common.cuh"
#include "kernels.h"

#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

const char* b200xf_last_error(void) { return bx::last_error(); }
int b200xf_version(void) { return 100; }
long long b200xf_launch_count(void) { return bx::launch_count(); }
void b200xf_prof_enable(int on) { bx::prof_enable(on); }
int b200xf_debug_set_timeline(long long* dev_ptr) { return bx::set_timeline(dev_ptr); }
void b200xf_gemm_force_cta_group(int cg) { bx::gemm_force_cta_group(cg); }
void b200xf_gemm_set_chunks_per_unit(int n) { bx::gemm_set_chunks_per_unit(n); }
int b200xf_prof_collect(double* out, int ncat) { return bx::prof_collect(out, ncat); }

int b200xf_gemm_bf16(const void* A, int64_t lda, int a_major, const void* B, int64_t ldb, int b_major, void* D,
                     int64_t ldd, int d_is_f32, const float* bias, const void* aux, int64_t ld_aux, int flags,
                     int splits, int64_t M, int64_t N, int64_t K, b200xf_stream_t stream) {
  return bx::gemm_bf16(A, lda, a_major, B, ldb, b_major, D, ldd, d_is_f32, bias, aux, ld_aux, flags, splits, M, N, K,
                       ST(stream));
}
int b200xf_gemm_bf16_colsum(const void* A, int64_t lda, int a_major, const void* B, int64_t ldb, int b_major, void* D,
                            int64_t ldd, const float* bias, const void* aux, int64_t ld_aux, int flags, int64_t M, int64_t N,
                            int64_t K, float* colsum, b200xf_stream_t stream) {
  return bx::gemm_bf16(A, lda, a_major, B, ldb, b_major, D, ldd, 0, bias, aux, ld_aux, flags, 1, M, N, K, ST(stream), colsum);
}
int b200xf_gemm_bias_dropout_residual_ln(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, const void* xres_bf16,
                                         const float* gamma, const float* beta, void* z_bf16, void* out_bf16, float* mean, float* rstd,
                                         void* keep_bits, int64_t M, int64_t N, int64_t K, float eps, float rate, uint64_t seed, uint32_t site,
                                         b200xf_stream_t stream) {
  return bx::gemm_ln(A, lda, W, ldw, bias, xres_bf16, gamma, beta, z_bf16, out_bf16, mean, rstd, keep_bits, M, N, K, eps, rate, seed, site,
                     ST(stream));
}
int b200xf_attn_fwd(const void* Q, int64_t ld_q, const void* K, int64_t ld_k, const void* V, int64_t ld_v, void* O,
                    int64_t ld_o, float* lse, const uint8_t* kpad, int causal, const float* maskf, int64_t mask_sb,
                    int64_t mask_sq, int B, int H, int Sq, int Sk, int dk, b200xf_stream_t stream) {
  return bx::attn_fwd(Q, ld_q, K, ld_k, V, ld_v, O, ld_o, lse, kpad, causal, maskf, mask_sb, mask_sq, B, H, Sq, Sk, dk,
                      ST(stream));
}
int b200xf_attn_fwd_residual(const void* Q, int64_t ld_q, const void* K, int64_t ld_k, const void* V, int64_t ld_v, void* O,
                             void* O_lo, int64_t ld_o, float* lse, const uint8_t* kpad, int causal, const float* maskf, int64_t mask_sb,
                             int64_t mask_sq, int B, int H, int Sq, int Sk, int dk, b200xf_stream_t stream) {
  return bx::attn_fwd(Q, ld_q, K, ld_k, V, ld_v, O, ld_o, lse, kpad, causal, maskf, mask_sb, mask_sq, B, H, Sq, Sk, dk,
                      ST(stream), O_lo);
}
int b200xf_attn_delta_residual(const void* O, const void* O_lo, int64_t ld_o, const void* dO, int64_t ld_do, float* delta, int B, int S,
                               int H, int dk, b200xf_stream_t stream) {
  return bx::attn_delta(O, ld_o, dO, ld_do, delta, B, S, H, dk, ST(stream), O_lo);
}
int b200xf_attn_bwd(const void* Q, int64_t ld_q, const void* K, int64_t ld_k, const void* V, int64_t ld_v,
                    const void* dO, int64_t ld_do, const float* lse, const float* delta, void* dq_acc, int64_t ld_dq,
                    int dq_is_bf16, void* dK, int64_t ld_dk, void* dV, int64_t ld_dv, const uint8_t* kpad, int causal,
                    const float* maskf, int64_t mask_sb, int64_t mask_sq, int B, int H, int Sq, int Sk, int dk,
                    b200xf_stream_t stream) {
  return bx::attn_bwd(Q, ld_q, K, ld_k, V, ld_v, dO, ld_do, lse, delta, dq_acc, ld_dq, dq_is_bf16, dK, ld_dk, dV, ld_dv, kpad, causal,
                      maskf, mask_sb, mask_sq, B, H, Sq, Sk, dk, ST(stream));
}
int b200xf_attn_delta(const void* O, int64_t ld_o, const void* dO, int64_t ld_do, float* delta, int B, int S, int H,
                      int dk, b200xf_stream_t stream) {
  return bx::attn_delta(O, ld_o, dO, ld_do, delta, B, S, H, dk, ST(stream));
}
int b200xf_attn_probs(const void* Q, int64_t ld_q, const void* K, int64_t ld_k, const float* lse, float* P,
                      const uint8_t* kpad, int causal, const float* maskf, int64_t mask_sb, int64_t mask_sq, int B,
                      int H, int Sq, int Sk, int dk, b200xf_stream_t stream) {
  return bx::attn_probs(Q, ld_q, K, ld_k, lse, P, kpad, causal, maskf, mask_sb, mask_sq, B, H, Sq, Sk, dk, ST(stream));
}
int b200xf_embed_fwd(const int64_t* ids, int id_stride, int id_offset, const float* table, const float* pos,
                     float* out_f32, void* out_bf16, int64_t rows, int S, int d, int64_t V, float rate, uint64_t seed,
                     uint32_t site, b200xf_stream_t stream) {
  return bx::embed_fwd(reinterpret_cast<const long long*>(ids), id_stride, id_offset, table, pos, out_f32, out_bf16, rows,
                       S, d, V, rate, seed, site, ST(stream));
}
int b200xf_embed_bwd(const int64_t* ids, int id_stride, int id_offset, const void* g1_bf16, const void* g2_bf16, float* dtable,
                     int64_t rows, int S, int d, int64_t V, float rate, uint64_t seed, uint32_t site, b200xf_stream_t stream) {
  return bx::embed_bwd(reinterpret_cast<const long long*>(ids), id_stride, id_offset, g1_bf16, g2_bf16, dtable, rows, S, d, V, rate,
                       seed, site, ST(stream));
}
int b200xf_residual_ln_fwd(const void* y_bf16, const void* x_bf16, const float* gamma, const float* beta, void* z_bf16,
                           void* out_bf16, float* mean, float* rstd, void* keep_bits, int64_t rows, int d, float eps, float rate,
                           uint64_t seed, uint32_t site, b200xf_stream_t stream) {
  return bx::residual_ln_fwd(y_bf16, x_bf16, gamma, beta, z_bf16, out_bf16, mean, rstd, keep_bits, rows, d, eps, rate, seed, site,
                             ST(stream));
}
int b200xf_residual_ln_bwd(const void* g1, int g1_is_f32, const void* g2_bf16, const void* z_bf16, const float* mean,
                           const float* rstd, const float* gamma, void* dz_bf16, void* dy_sub_bf16, float* dgamma,
                           float* dbeta, float* dbias, const void* keep_bits, int64_t rows, int d, float rate, uint64_t seed,
                           uint32_t site, b200xf_stream_t stream) {
  return bx::residual_ln_bwd(g1, g1_is_f32, g2_bf16, z_bf16, mean, rstd, gamma, dz_bf16, dy_sub_bf16, dgamma, dbeta, dbias, keep_bits,
                             rows, d, rate, seed, site, ST(stream));
}
int b200xf_xent(const void* logits, int logits_f32, int64_t ld, const int64_t* labels, int label_stride,
                int label_offset, int T, void* dlogits, int64_t ld_d, float* row_loss, double* stats, int64_t rows,
                int V, float inv_batch, b200xf_stream_t stream) {
  return bx::xent(logits, logits_f32, ld, reinterpret_cast<const long long*>(labels), label_stride, label_offset, T,
                  dlogits, ld_d, row_loss, stats, rows, V, inv_batch, ST(stream));
}
int b200xf_colsum_bf16(const void* X, int64_t ld, float* out, int64_t rows, int N, b200xf_stream_t stream) {
  return bx::colsum_bf16(X, ld, out, rows, N, ST(stream));
}
int b200xf_adam_step(float* p, float* m, float* v, const float* g, void* shadow_bf16, int64_t n, float alpha,
                     float beta1, float beta2, float eps, b200xf_stream_t stream) {
  return bx::adam_step(p, m, v, g, shadow_bf16, n, alpha, beta1, beta2, eps, ST(stream));
}
int b200xf_adam_step_dev(float* p, float* m, float* v, const float* g, void* shadow_bf16, int64_t n, const float* alpha_dev,
                         float beta1, float beta2, float eps, b200xf_stream_t stream) {
  return bx::adam_step(p, m, v, g, shadow_bf16, n, 0.f, beta1, beta2, eps, ST(stream), alpha_dev);
}
int b200xf_step_params(int64_t* state, uint64_t base_seed, float d_model, float warmup_steps, float const_lr, float beta1,
                       float beta2, uint64_t* seed_out, float* alpha_out, float* lr_out, b200xf_stream_t stream) {
  return bx::step_params(reinterpret_cast<long long*>(state), base_seed, d_model, warmup_steps, const_lr, beta1, beta2,
                         reinterpret_cast<unsigned long long*>(seed_out), alpha_out, lr_out, ST(stream));
}
int b200xf_cast_f32_bf16(const float* in, void* out, int64_t n, b200xf_stream_t stream) {
  return bx::cast_f32_bf16(in, out, n, ST(stream));
}
int b200xf_cast_bf16_f32(const void* in, float* out, int64_t n, b200xf_stream_t stream) {
  return bx::cast_bf16_f32(in, out, n, ST(stream));
}
int b200xf_cast_f32_bf16_strided(const float* in, void* out, int64_t rows, int d, int64_t ld_out,
                                 b200xf_stream_t stream) {
  return bx::cast_f32_bf16_strided(in, out, rows, d, ld_out, ST(stream));
}
int b200xf_pad_mask(const int64_t* ids, int id_stride, int id_offset, uint8_t* mask, int64_t rows, int S,
                    b200xf_stream_t stream) {
  return bx::pad_mask(reinterpret_cast<const long long*>(ids), id_stride, id_offset, mask, rows, S, ST(stream));
}

}  // extern "C"
