// Internal C++ declarations of the kernel launchers (definitions in gemm.cu / attention.cu / elementwise.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bx {

int gemm_bf16(const void* A, long long lda, int a_major, const void* B, long long ldb, int b_major, void* D,
              long long ldd, int d_is_f32, const float* bias, const void* aux, long long ld_aux, int flags,
              int splits, long long M, long long N, long long K, cudaStream_t stream, float* colsum = nullptr);

// K3/K4 fused: z = xres + dropout(A.W + bias); out = LayerNorm(z) * gamma + beta, in the GEMM epilogue (gemm_ln.cu); N = 128 / 256 / 512
bool gemm_ln_supported(long long M, long long N, long long K);
int gemm_ln(const void* A, long long lda, const void* W, long long ldw, const float* bias, const void* xres, const float* gamma,
            const float* beta, void* z, void* out, float* mean, float* rstd, void* keep_bits, long long M, long long N, long long K, float eps,
            float rate, unsigned long long seed, unsigned site, cudaStream_t stream, const unsigned long long* seed_dev = nullptr);

void gemm_set_chunks_per_unit(int n);   // 0: persistent GEMM grids (default); k > 0: ~k chunks of tiles per SM, scheduled by the hardware
void gemm_force_cta_group(int cg);   // 0 = heuristic, 1 / 2 = force single-CTA / CTA-pair tiles (tests, benchmarks)

int set_timeline(long long* dev_ptr);   // debug: register a device buffer of >= 128 int64 clock stamps (nullptr disables)

int attn_fwd(const void* Q, long long ld_q, const void* K, long long ld_k, const void* V, long long ld_v, void* O,
             long long ld_o, float* lse, const uint8_t* kpad, int causal, const float* maskf, long long mask_sb,
             long long mask_sq, int B, int H, int Sq, int Sk, int dk, cudaStream_t st,
             void* O_lo = nullptr,    // O_lo (optional, same layout as O): bf16 rounding residual of O for the backward's delta
             long long bs_q = 0, long long bs_kv = 0);   // batch strides of Q and of K/V in elements (0: dense [B, S, ld]); a K/V cache has its own
int attn_bwd(const void* Q, long long ld_q, const void* K, long long ld_k, const void* V, long long ld_v, const void* dO,
             long long ld_do, const float* lse, const float* delta, void* dq_acc, long long ld_dq, int dq_bf16, void* dK,
             long long ld_dk, void* dV, long long ld_dv, const uint8_t* kpad, int causal, const float* maskf,
             long long mask_sb, long long mask_sq, int B, int H, int Sq, int Sk, int dk, cudaStream_t st);
int attn_probs(const void* Q, long long ld_q, const void* K, long long ld_k, const float* lse, float* P,
               const uint8_t* kpad, int causal, const float* maskf, long long mask_sb, long long mask_sq, int B, int H,
               int Sq, int Sk, int dk, cudaStream_t st);
int attn_delta(const void* O, long long ld_o, const void* dO, long long ld_do, float* delta, int B, int S, int H, int dk,
               cudaStream_t st, const void* O_lo = nullptr);   // delta = rowsum(dO o (O + O_lo))

// ids of row r = (b, s) are read at ids[b * id_stride + id_offset + s]  (tar[:, :-1] without a copy)
int embed_fwd(const long long* ids, int id_stride, int id_offset, const float* table, const float* pos, float* out_f32,
              void* out_bf16, long long rows, int S, int d, long long V, float rate, unsigned long long seed, unsigned site,
              cudaStream_t st, const unsigned long long* seed_dev = nullptr);   // seed_dev != NULL overrides `seed` (graph replays)
int embed_bwd(const long long* ids, int id_stride, int id_offset, const void* g1, const void* g2, float* dtable, long long rows,
              int S, int d, long long V, float rate, unsigned long long seed, unsigned site, cudaStream_t st,
              const unsigned long long* seed_dev = nullptr);   // gradient of the embedding output = g1 (+ g2), bf16 streams
// y, x, z, out: bf16 [rows, d]
// keep_bits: uint8 [rows, d/8] dropout keep mask (bit e of byte c8 = element 8*c8+e); written by fwd when rate > 0 and
// non-NULL, read back by bwd (NULL there: the mask is regenerated from (seed, site))
int residual_ln_fwd(const void* y, const void* x, const float* gamma, const float* beta, void* z, void* out, float* mean,
                    float* rstd, void* keep_bits, long long rows, int d, float eps, float rate, unsigned long long seed,
                    unsigned site, cudaStream_t st, const unsigned long long* seed_dev = nullptr);
// incoming gradient = g1 (fp32 if g1_f32 else bf16) + g2 (bf16, may be NULL); dz, dy_sub bf16 (dy_sub NULL: not written)
int residual_ln_bwd(const void* g1, int g1_f32, const void* g2, const void* z, const float* mean, const float* rstd,
                    const float* gamma, void* dz, void* dy_sub, float* dgamma, float* dbeta, float* dbias, const void* keep_bits,
                    long long rows, int d, float rate, unsigned long long seed, unsigned site, cudaStream_t st,
                    const unsigned long long* seed_dev = nullptr);
int xent(const void* logits, int logits_f32, long long ld, const long long* labels, int label_stride, int label_offset,
         int T_len, void* dlogits, long long ld_d, float* row_loss, double* stats, long long rows, int V, float inv_batch,
         cudaStream_t st, const int* t_valid = nullptr);   // device int (or NULL): positions t >= *t_valid are shape-bucketing filler: no accuracy counters
int colsum_bf16(const void* X, long long ld, float* out, long long rows, int N, cudaStream_t st);
int adam_step(float* p, float* m, float* v, const float* g, void* shadow, long long n, float alpha, float b1, float b2,
              float eps, cudaStream_t st, const float* alpha_dev = nullptr);
int step_params(long long* state, unsigned long long base_seed, float d_model, float warmup, float const_lr, float b1, float b2,
                unsigned long long* seed_out, float* alpha_out, float* lr_out, cudaStream_t st);
int cast_f32_bf16(const float* in, void* out, long long n, cudaStream_t st);
int cast_bf16_f32(const void* in, float* out, long long n, cudaStream_t st);
int cast_f32_bf16_strided(const float* in, void* out, long long rows, int d, long long ld_out, cudaStream_t st);
int pad_mask(const long long* ids, int id_stride, int id_offset, uint8_t* mask, long long rows, int S, cudaStream_t st);

}  // namespace bx
