// HBM-bound kernels of the training step: embedding (+pos-enc, dropout), dropout+residual+LayerNorm fwd/bwd,
// softmax-cross-entropy (+ its gradient, accuracy counters), column sums (bias grads), fused Adam (+bf16 shadow),
// casts and masks.  All vectorised 16-byte accesses, fp32 math.  Reference call sites are cited per kernel.
#include <mutex>

#include "common.cuh"

namespace bx {

// ------------------------------------------------------------------------------------------------
// counter-based dropout RNG (Philox-4x32, 7 rounds) keyed by (seed, site, element); the LayerNorm kernels store the keep mask
// (1 bit per element) in forward and read it back in backward, the embedding kernels regenerate it
// keras Dropout semantics: keep w.p. 1-rate, scale kept by 1/(1-rate).
// ------------------------------------------------------------------------------------------------
// keep-scale factors for the 4 consecutive elements starting at flat element index `idx4*4`
__device__ __forceinline__ float4 dropout_scale4(unsigned long long seed, uint32_t site, unsigned long long idx4,
                                                 float rate) {
  const uint4 r = philox4x32(make_uint4((uint32_t)idx4, (uint32_t)(idx4 >> 32), site, 0u),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float inv = 1.f / (1.f - rate);
  const float s = 2.3283064365386963e-10f;  // 2^-32
  return make_float4((r.x * s >= rate) ? inv : 0.f, (r.y * s >= rate) ? inv : 0.f, (r.z * s >= rate) ? inv : 0.f,
                     (r.w * s >= rate) ? inv : 0.f);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ------------------------------------------------------------------------------------------------
// 8-wide bf16 helpers: the activation and gradient streams between sublayers are bf16 in HBM (half the traffic of
// fp32 streams); all arithmetic below is fp32 in registers.
// ------------------------------------------------------------------------------------------------
struct F8 { float v[8]; };
__device__ __forceinline__ F8 unpack8(const uint4 u) {
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  F8 f;
  f.v[0] = a.x; f.v[1] = a.y; f.v[2] = b.x; f.v[3] = b.y; f.v[4] = c.x; f.v[5] = c.y; f.v[6] = d.x; f.v[7] = d.y;
  return f;
}
__device__ __forceinline__ uint4 pack8(const F8& f) {
  return make_uint4(pack_bf16x2(f.v[0], f.v[1]), pack_bf16x2(f.v[2], f.v[3]), pack_bf16x2(f.v[4], f.v[5]), pack_bf16x2(f.v[6], f.v[7]));
}
// ------------------------------------------------------------------------------------------------
// K6  embedding * sqrt(d) + pos_encoding[:, :S] (+ dropout)      Encoder.py:51-55, Decoder.py:65-69
//   ids int64 [rows] (rows = B*S), table fp32 [V, d], pos fp32 [P, d]  ->  x bf16 [rows, d] (+ optional fp32 copy)
// ------------------------------------------------------------------------------------------------
__global__ void embed_fwd_kernel(const long long* __restrict__ ids, const float* __restrict__ table,
                                 const float* __restrict__ pos, float* __restrict__ out_f32,
                                 __nv_bfloat16* __restrict__ out_bf16, long long rows, int S, int id_stride,
                                 int id_offset, int d, long long V, float scale, float rate, unsigned long long seed,
                                 uint32_t site, const unsigned long long* __restrict__ seed_dev) {
  if (seed_dev != nullptr) seed = *seed_dev;      // CUDA-graph replays: the per-step seed lives on the device
  const int d4 = d >> 2;
  const long long total = rows * d4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / d4;
    const int c4 = (int)(i - r * d4);
    long long id = ids[(r / S) * id_stride + id_offset + (r % S)];
    id = id < 0 ? 0 : (id >= V ? V - 1 : id);
    const float4 e = __ldg(reinterpret_cast<const float4*>(table + id * d) + c4);
    const float4 p = __ldg(reinterpret_cast<const float4*>(pos + (r % S) * (long long)d) + c4);
    float4 o = make_float4(e.x * scale + p.x, e.y * scale + p.y, e.z * scale + p.z, e.w * scale + p.w);
    if (rate > 0.f) {
      const float4 k = dropout_scale4(seed, site, (unsigned long long)i, rate);
      o.x *= k.x; o.y *= k.y; o.z *= k.z; o.w *= k.w;
    }
    if (out_f32 != nullptr) reinterpret_cast<float4*>(out_f32)[i] = o;
    reinterpret_cast<uint2*>(out_bf16)[i] = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
  }
}

// backward: dTable[id] += (g1 + g2) * sqrt(d) * dropmask   (dense-equivalent IndexedSlices sum, Appendix A.8)
//   the gradient of the embedding output arrives as two bf16 streams (residual path + first sublayer's dgrad)
__global__ void embed_bwd_kernel(const long long* __restrict__ ids, const __nv_bfloat16* __restrict__ g1,
                                 const __nv_bfloat16* __restrict__ g2, float* __restrict__ dtable, long long rows, int S,
                                 int id_stride, int id_offset, int d, long long V, float scale, float rate,
                                 unsigned long long seed, uint32_t site, const unsigned long long* __restrict__ seed_dev) {
  if (seed_dev != nullptr) seed = *seed_dev;
  const int d4 = d >> 2;
  const long long total = rows * d4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / d4;
    const int c4 = (int)(i - r * d4);
    long long id = ids[(r / S) * id_stride + id_offset + (r % S)];
    id = id < 0 ? 0 : (id >= V ? V - 1 : id);
    const uint2 a = reinterpret_cast<const uint2*>(g1)[i];
    const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y);
    float4 g = make_float4(a0.x, a0.y, a1.x, a1.y);
    if (g2 != nullptr) {
      const uint2 b = reinterpret_cast<const uint2*>(g2)[i];
      const float2 b0 = unpack_bf16x2(b.x), b1 = unpack_bf16x2(b.y);
      g.x += b0.x; g.y += b0.y; g.z += b1.x; g.w += b1.y;
    }
    if (rate > 0.f) {
      const float4 k = dropout_scale4(seed, site, (unsigned long long)i, rate);
      g.x *= k.x; g.y *= k.y; g.z *= k.z; g.w *= k.w;
    }
    float* dst = dtable + id * d + c4 * 4;
    atomicAdd(dst + 0, g.x * scale);
    atomicAdd(dst + 1, g.y * scale);
    atomicAdd(dst + 2, g.z * scale);
    atomicAdd(dst + 3, g.w * scale);
  }
}

// ------------------------------------------------------------------------------------------------
// K3/K4 epilogue (stand-alone form) + K5:  z = x + dropout(y);  out = LayerNorm(z)   (eps 1e-6, biased var)
//   Encoder.py:22-23,26-27  Decoder.py:30-31,35-36,39-40 ; LN formula SURVEY Appendix A.4
//   y (sublayer output), x (residual input), z (kept for backward) and out are bf16 [rows, d]; statistics and the
//   normalisation use the unrounded fp32 z.  One warp per row, NV = ceil(d/256) 16-byte vectors per lane; the next
//   row's loads are issued before the current row is reduced (software pipeline).
// ------------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256)
residual_ln_fwd_kernel(const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ x,
                       const float* __restrict__ gamma, const float* __restrict__ beta, __nv_bfloat16* __restrict__ z_out,
                       __nv_bfloat16* __restrict__ out, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                       uint8_t* __restrict__ keep_out, long long rows, int d, float eps, float rate, unsigned long long seed, uint32_t site,
                       const unsigned long long* __restrict__ seed_dev) {
  if (seed_dev != nullptr) seed = *seed_dev;
  const int lane = threadIdx.x & 31;
  const long long warp_global = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int d8 = d >> 3;
  const float inv_keep = 1.f / (1.f - rate);
  uint4 ny[NV], nx[NV];
  auto fetch = [&](long long r) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c8 = lane + 32 * i;
      if (c8 < d8) {
        ny[i] = reinterpret_cast<const uint4*>(y)[r * d8 + c8];
        nx[i] = reinterpret_cast<const uint4*>(x)[r * d8 + c8];
      } else { ny[i] = make_uint4(0, 0, 0, 0); nx[i] = make_uint4(0, 0, 0, 0); }
    }
  };
  if (warp_global < rows) fetch(warp_global);
  for (long long r = warp_global; r < rows; r += nwarps) {
    F8 z[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c8 = lane + 32 * i;
      F8 yv = unpack8(ny[i]);
      const F8 xv = unpack8(nx[i]);
      if (rate > 0.f && c8 < d8) {
        const uint32_t keep = dropout_keep8(seed, site, (unsigned long long)(r * d8 + c8), rate);
        if (keep_out != nullptr) keep_out[r * d8 + c8] = (uint8_t)keep;     // the backward pass reads the bits back: no RNG there
#pragma unroll
        for (int e = 0; e < 8; ++e) yv.v[e] = ((keep >> e) & 1u) ? yv.v[e] * inv_keep : 0.f;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) { z[i].v[e] = xv.v[e] + yv.v[e]; s += z[i].v[e]; }
    }
    if (r + nwarps < rows) fetch(r + nwarps);
    const float mean = warp_sum(s) / d;
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (lane + 32 * i < d8) {
#pragma unroll
        for (int e = 0; e < 8; ++e) { const float a = z[i].v[e] - mean; v += a * a; }
      }
    }
    const float rstd = rsqrtf(warp_sum(v) / d + eps);
    if (lane == 0) { mean_out[r] = mean; rstd_out[r] = rstd; }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c8 = lane + 32 * i;
      if (c8 < d8) {
        const long long idx = r * d8 + c8;
        reinterpret_cast<uint4*>(z_out)[idx] = pack8(z[i]);
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * c8), g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * c8 + 1);
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * c8), b1 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * c8 + 1);
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        F8 o;
#pragma unroll
        for (int e = 0; e < 8; ++e) o.v[e] = (z[i].v[e] - mean) * rstd * g[e] + b[e];
        reinterpret_cast<uint4*>(out)[idx] = pack8(o);
      }
    }
  }
}

constexpr int kLnBwdStages = 3;
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* gptr) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(gptr) : "memory");
}
__device__ __forceinline__ uint4 lds_u4(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}

// backward of out = LN(z), z = x + dropout(y); the incoming gradient is g1 (+ g2): the residual-path gradient and the
// previous sublayer's dgrad output arrive as separate streams and are summed here in fp32 (no read-modify-write
// accumulation in a GEMM epilogue, no intermediate bf16 rounding of the sum):
//   dz = rstd * (g*dy - mean(g*dy) - xhat * mean(g*dy*xhat))            (bf16, gradient to the residual input x)
//   dy_sub = dz * dropmask/(1-rate)                                      (bf16, gradient to the sublayer output;
//                                                                         NULL when rate == 0: it equals dz)
//   dgamma += sum_rows dy*xhat ; dbeta += sum_rows dy ; dbias += sum_rows dy_sub   (block partials -> atomicAdd)
// kRegen: the dropout mask is regenerated with Philox (no stored bits); that variant needs the registers of a 256-thread block
template <int NV, bool kG1F32, bool kRegen, bool kFull>
__global__ void __launch_bounds__((NV <= 2 && !kRegen) ? 512 : 256, 1)
residual_ln_bwd_kernel(const void* __restrict__ g1v, const __nv_bfloat16* __restrict__ g2, const __nv_bfloat16* __restrict__ z,
                       const float* __restrict__ mean_in, const float* __restrict__ rstd_in, const float* __restrict__ gamma,
                       __nv_bfloat16* __restrict__ dz, __nv_bfloat16* __restrict__ dy_sub, float* __restrict__ dgamma,
                       float* __restrict__ dbeta, float* __restrict__ dbias, const uint8_t* __restrict__ keep_in, long long rows, int d, float rate,
                       unsigned long long seed, uint32_t site, const unsigned long long* __restrict__ seed_dev) {
  if (seed_dev != nullptr) seed = *seed_dev;
  extern __shared__ float red[];  // the input ring; at the end the warps' dgamma / dbeta / dbias partials [nw][3][d]
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const long long warp_global = (long long)blockIdx.x * nw + wib;
  const long long nwarps = (long long)gridDim.x * nw;
  const int d8 = d >> 3;
  F8 ag[NV], ab[NV], ac[NV];      // gamma is re-read per row (L1-resident) instead of living in 8*NV registers
#pragma unroll
  for (int i = 0; i < NV; ++i) {
#pragma unroll
    for (int e = 0; e < 8; ++e) { ag[i].v[e] = 0.f; ab[i].v[e] = 0.f; ac[i].v[e] = 0.f; }
  }
  // Input pipeline: every warp owns a ring of kLnBwdStages row slots in shared memory, filled by cp.async (16 bytes per lane and vector)
  // two rows ahead of the one being reduced -- a lane reads back exactly the bytes it copied, so no barrier is involved.  With the
  // prefetch in REGISTERS (one row ahead, 26 registers) a 512-thread block kept ~49 KB in flight per SM and the kernel ran at ~0.6 of
  // the HBM copy rate; the ring holds 16 warps x 2 rows x 3 KB.
  constexpr int G1B = kG1F32 ? 4 : 2;                       // bytes per element of g1
  const int g1_bytes = d * G1B, slot_bytes = g1_bytes + 4 * d;     // g1 | g2 | z
  uint8_t* const ring = reinterpret_cast<uint8_t*>(red) + (size_t)wib * kLnBwdStages * slot_bytes;
  const uint32_t ring_s = smem_u32(ring);
  // running 16-byte-vector offsets of the rows in flight (row * d8 + lane): one 64-bit add per row instead of a 64-bit multiply per access
  const long long vstride = nwarps * d8;
  auto issue = [&](long long r, long long voff, int slot) {
    if (r < rows) {
      const uint32_t sbase = ring_s + slot * slot_bytes;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c8 = lane + 32 * i;
        if (c8 < d8) {
          const long long idx = voff + 32 * i;
          if (kG1F32) {
            cp_async16(sbase + c8 * 32, reinterpret_cast<const uint4*>(g1v) + 2 * idx);
            cp_async16(sbase + c8 * 32 + 16, reinterpret_cast<const uint4*>(g1v) + 2 * idx + 1);
          } else {
            cp_async16(sbase + c8 * 16, reinterpret_cast<const uint4*>(g1v) + idx);
          }
          if (g2 != nullptr) cp_async16(sbase + g1_bytes + c8 * 16, reinterpret_cast<const uint4*>(g2) + idx);
          cp_async16(sbase + g1_bytes + 2 * d + c8 * 16, reinterpret_cast<const uint4*>(z) + idx);
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");    // (an empty group past the last row keeps the wait count uniform)
  };
  const float inv_keep = 1.f / (1.f - rate);
  uint32_t nkeep[NV], nk[NV];
  float nmean = 0.f, nrstd = 0.f;
  auto fetch_small = [&](long long r, long long voff) {     // one row ahead in registers: keep byte(s), mean, rstd
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c8 = lane + 32 * i;
      nkeep[i] = (!kRegen && keep_in != nullptr && c8 < d8) ? keep_in[voff + 32 * i] : 0u;
    }
    nmean = mean_in[r]; nrstd = rstd_in[r];
  };
  long long voff = warp_global * d8 + lane;                 // of the row being reduced
  issue(warp_global, voff, 0);
  issue(warp_global + nwarps, voff + vstride, 1);
  if (warp_global < rows) fetch_small(warp_global, voff);
  int slot = 0;
  for (long long r = warp_global; r < rows; r += nwarps, voff += vstride) {
    issue(r + 2 * nwarps, voff + 2 * vstride, slot == 0 ? 2 : slot - 1);        // (the slot read in the previous iteration)
    asm volatile("cp.async.wait_group 2;" ::: "memory");    // this row's group has landed
    F8 gd[NV], xh[NV];              // gamma * dy and xhat of this row
    const float rstd = nrstd, nmr = -nmean * nrstd;
    float s1 = 0.f, s2 = 0.f;
    const uint32_t sbase = ring_s + slot * slot_bytes;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c8 = lane + 32 * i;
      const bool live = kFull || c8 < d8;      // kFull: d = 256 NV, every lane vector is inside the row (no per-element selects)
      const int cg = live ? c8 : 0;
      F8 dy;
      if (kG1F32) {
        const float4 lo = ld_shared_f4(sbase + cg * 32), hi = ld_shared_f4(sbase + cg * 32 + 16);
        dy.v[0] = lo.x; dy.v[1] = lo.y; dy.v[2] = lo.z; dy.v[3] = lo.w; dy.v[4] = hi.x; dy.v[5] = hi.y; dy.v[6] = hi.z; dy.v[7] = hi.w;
      } else {
        dy = unpack8(lds_u4(sbase + cg * 16));
      }
      F8 b;
      if (g2 != nullptr) {
        b = unpack8(lds_u4(sbase + g1_bytes + cg * 16));
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) b.v[e] = 0.f;
      }
      const F8 zv = unpack8(lds_u4(sbase + g1_bytes + 2 * d + cg * 16));
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * cg), g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * cg + 1);
      const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      nk[i] = nkeep[i];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float dv = dy.v[e] + b.v[e];
        xh[i].v[e] = live ? fmaf(zv.v[e], rstd, nmr) : 0.f;
        gd[i].v[e] = live ? gm[e] * dv : 0.f;
        s1 += gd[i].v[e];
        s2 += gd[i].v[e] * xh[i].v[e];
        ag[i].v[e] += live ? dv * xh[i].v[e] : 0.f;
        ab[i].v[e] += live ? dv : 0.f;
      }
    }
    slot = slot == 2 ? 0 : slot + 1;
    if (r + nwarps < rows) fetch_small(r + nwarps, voff + vstride);
    const float inv_d = 1.f / (float)d;
    const float nrm1 = -rstd * warp_sum(s1) * inv_d, nrm2 = -rstd * warp_sum(s2) * inv_d;     // dz = rstd gd - rstd m1 - xhat rstd m2
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c8 = lane + 32 * i;
      if (c8 < d8) {
        const long long idx = voff + 32 * i;
        F8 o;
#pragma unroll
        for (int e = 0; e < 8; ++e) o.v[e] = fmaf(gd[i].v[e], rstd, fmaf(xh[i].v[e], nrm2, nrm1));
        reinterpret_cast<uint4*>(dz)[idx] = pack8(o);
        if (rate > 0.f) {
          const uint32_t keep = kRegen ? dropout_keep8(seed, site, (unsigned long long)idx, rate) : nk[i];
#pragma unroll
          for (int e = 0; e < 8; ++e) o.v[e] = ((keep >> e) & 1u) ? o.v[e] * inv_keep : 0.f;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) ac[i].v[e] += o.v[e];       // bias grad of the sublayer's last Dense
        if (dy_sub != nullptr) reinterpret_cast<uint4*>(dy_sub)[idx] = pack8(o);
      }
    }
  }
  // column partials: every warp parks its 3 x d sums in shared memory (the ring is dead now), then each thread adds up one column over
  // the warps and issues ONE global atomic per column and block (shared-memory atomics here cost 1/6 of the kernel: 16 warps, same addresses)
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  float* const part = red + (size_t)wib * 3 * d;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (lane + 32 * i) * 8;
    if (c < d) {
      *reinterpret_cast<float4*>(part + c) = make_float4(ag[i].v[0], ag[i].v[1], ag[i].v[2], ag[i].v[3]);
      *reinterpret_cast<float4*>(part + c + 4) = make_float4(ag[i].v[4], ag[i].v[5], ag[i].v[6], ag[i].v[7]);
      *reinterpret_cast<float4*>(part + d + c) = make_float4(ab[i].v[0], ab[i].v[1], ab[i].v[2], ab[i].v[3]);
      *reinterpret_cast<float4*>(part + d + c + 4) = make_float4(ab[i].v[4], ab[i].v[5], ab[i].v[6], ab[i].v[7]);
      *reinterpret_cast<float4*>(part + 2 * d + c) = make_float4(ac[i].v[0], ac[i].v[1], ac[i].v[2], ac[i].v[3]);
      *reinterpret_cast<float4*>(part + 2 * d + c + 4) = make_float4(ac[i].v[4], ac[i].v[5], ac[i].v[6], ac[i].v[7]);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 3 * d; c += blockDim.x) {
    float v = 0.f;
    for (int w = 0; w < nw; ++w) v += red[(size_t)w * 3 * d + c];
    if (c < d) atomicAdd(dgamma + c, v);
    else if (c < 2 * d) atomicAdd(dbeta + c - d, v);
    else if (dbias != nullptr) atomicAdd(dbias + c - 2 * d, v);
  }
}

// ------------------------------------------------------------------------------------------------
// K7 (stand-alone form): masked sparse softmax cross-entropy over materialised logits + its gradient
//   train.py:83-88 (loss = sum_{real!=0} (lse - logit[real]) / batch_size), train.py:141 (accuracy over ALL
//   positions incl. pads), gradient (softmax - onehot) * mask / batch_size  (SURVEY Appendix A.7).
//   one CTA per row; logits bf16 or fp32 [rows, ld], V valid columns; dlogits bf16 [rows, ld_d] (may alias
//   a bf16 logits buffer).  stats: double[0] loss_sum, [1] correct, [2] positions, [3] nonpad
// ------------------------------------------------------------------------------------------------
template <typename T> struct LogitVec;
template <> struct LogitVec<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void load(const float* p, float (&x)[8]) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
  }
};
template <> struct LogitVec<__nv_bfloat16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&x)[8]) {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    const float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y), c = unpack_bf16x2(v.z), d = unpack_bf16x2(v.w);
    x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y; x[4] = c.x; x[5] = c.y; x[6] = d.x; x[7] = d.y;
  }
};

template <typename T>
__global__ void __launch_bounds__(256)
xent_kernel(const T* logits, long long ld, const long long* __restrict__ labels, int label_stride,
            int label_offset, int T_len, __nv_bfloat16* dlogits, long long ld_d, float* __restrict__ row_loss,
            double* __restrict__ stats, int V, float inv_batch, const int* __restrict__ t_valid_dev) {   // dlogits may alias logits (no __restrict__)
  // one CTA per row, 16-byte vector loads.  labels for row r = (b, t) live at labels[b*label_stride + label_offset + t].
  // The kernel is instruction-issue bound (ncu: 23 instructions per logit before this version, issue slots 73 % busy next to 59 % MUFU
  // and 62 % DRAM), so the inner loops carry nothing per ELEMENT beyond unpack / max / exp / add: the ragged last vector is peeled, the
  // running maximum is updated branch-free, the arg-max is tracked per VECTOR (its element is looked up once, after the reduction), the
  // label's logit is fetched by one thread up front and the label's gradient correction is a per-vector test.
  constexpr int VEC = LogitVec<T>::N;
  constexpr float L2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
  const long long r = blockIdx.x;
  const T* row = logits + r * ld;
  const long long b = r / T_len, t = r % T_len;
  const long long raw_label = labels[b * label_stride + label_offset + t];
  const bool nonpad = raw_label != 0;
  const int label = (raw_label < 0 || raw_label >= V) ? 0 : (int)raw_label;
  __shared__ float s_m[8], s_s[8];
  __shared__ int s_i[8];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int nfull = V / VEC;                        // vectors entirely inside the row
  const int nvec = (V + VEC - 1) / VEC;             // ld is a multiple of VEC, so every vector lies inside the allocation
  float lab_logit = 0.f;
  if (threadIdx.x == 0) lab_logit = (float)row[label];        // before pass 2 may overwrite the (aliased) row in place
  // ---- pass 1: online max / sum-exp, and the first VECTOR attaining the maximum
  float m = -INFINITY, s = 0.f;
  int bi = 0x7fffffff;
  auto fold = [&](const float (&x)[8], int i) {
    float vmax = x[0];
#pragma unroll
    for (int e = 1; e < VEC; ++e) vmax = fmaxf(vmax, x[e]);
    bi = vmax > m ? i : bi;
    const float mn = fmaxf(m, vmax), mb = mn * L2E;
    float acc = s * fast_exp2((m - mn) * L2E);       // (m = -inf, s = 0 at the start: 0 * 2^-inf = 0)
#pragma unroll
    for (int e = 0; e < VEC; ++e) acc += fast_exp2(fmaf(x[e], L2E, -mb));
    s = acc; m = mn;
  };
  for (int i = threadIdx.x; i < nfull; i += blockDim.x) {
    float x[8];
    LogitVec<T>::load(row + (long long)i * VEC, x);
    fold(x, i);
  }
  if (nfull < nvec && (nfull % (int)blockDim.x) == (int)threadIdx.x) {        // the ragged last vector: columns >= V are not logits
    float x[8];
    LogitVec<T>::load(row + (long long)nfull * VEC, x);
#pragma unroll
    for (int e = 0; e < VEC; ++e) if (nfull * VEC + e >= V) x[e] = -INFINITY;
    fold(x, nfull);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    const int a2 = __shfl_xor_sync(0xffffffffu, bi, o);
    const float mn = fmaxf(m, m2);
    s = (m == -INFINITY ? 0.f : s * fast_exp2((m - mn) * L2E)) + (m2 == -INFINITY ? 0.f : s2 * fast_exp2((m2 - mn) * L2E));
    bi = (m2 > m || (m2 == m && a2 < bi)) ? a2 : bi;
    m = mn;
  }
  if (lane == 0) { s_m[wib] = m; s_s[wib] = s; s_i[wib] = bi; }
  __syncthreads();
  m = s_m[0]; s = s_s[0]; bi = s_i[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
    const float m2 = s_m[w], s2 = s_s[w];
    const int a2 = s_i[w];
    const float mn = fmaxf(m, m2);
    s = (m == -INFINITY ? 0.f : s * fast_exp2((m - mn) * L2E)) + (m2 == -INFINITY ? 0.f : s2 * fast_exp2((m2 - mn) * L2E));
    bi = (m2 > m || (m2 == m && a2 < bi)) ? a2 : bi;
    m = mn;
  }
  const float lse = m + __log2f(s) * LN2;
  if (threadIdx.x == 0) {
    // arg-max = first element of the first vector that attains the maximum (tf.argmax); the row is still intact here
    float x[8];
    LogitVec<T>::load(row + (long long)bi * VEC, x);
    int am = bi * VEC;
#pragma unroll
    for (int e = VEC - 1; e >= 0; --e) if (x[e] == m && bi * VEC + e < V) am = bi * VEC + e;
    const float loss = nonpad ? (lse - lab_logit) : 0.f;
    if (row_loss) row_loss[r] = loss;
    if (nonpad) atomicAdd(&stats[0], (double)loss);
    if (t_valid_dev == nullptr || t < *t_valid_dev) {   // positions added by shape bucketing (t >= *t_valid_dev) are not part of the batch
      if ((long long)am == raw_label) atomicAdd(&stats[1], 1.0);
      atomicAdd(&stats[2], 1.0);
    }
    if (nonpad) atomicAdd(&stats[3], 1.0);
  }
  if (dlogits == nullptr) return;
  __syncthreads();                                   // thread 0 has read its arg-max vector: the row may now be overwritten in place
  // ---- pass 2: gradient (row re-read hits L1/L2); padded columns [V, ld_d) are zeroed
  const float gscale = nonpad ? inv_batch : 0.f;
  const float lb = lse * L2E;
  const int lab_vec = label / VEC, lab_e = label % VEC;
  __nv_bfloat16* drow = dlogits + r * ld_d;
  auto store = [&](const float (&g)[8], int c0) {
    if (VEC == 8) {
      *reinterpret_cast<uint4*>(drow + c0) = make_uint4(pack_bf16x2(g[0], g[1]), pack_bf16x2(g[2], g[3]), pack_bf16x2(g[4], g[5]),
                                                        pack_bf16x2(g[6], g[7]));
    } else {
      *reinterpret_cast<uint2*>(drow + c0) = make_uint2(pack_bf16x2(g[0], g[1]), pack_bf16x2(g[2], g[3]));
    }
  };
  for (int i = threadIdx.x; i < nfull; i += blockDim.x) {
    float x[8];
    LogitVec<T>::load(row + (long long)i * VEC, x);
    float g[8];
#pragma unroll
    for (int e = 0; e < VEC; ++e) g[e] = fast_exp2(fmaf(x[e], L2E, -lb)) * gscale;
    if (i == lab_vec) {
#pragma unroll
      for (int e = 0; e < VEC; ++e) if (e == lab_e) g[e] -= gscale;
    }
    store(g, i * VEC);
  }
  if (nfull < nvec && (nfull % (int)blockDim.x) == (int)threadIdx.x) {
    float x[8];
    LogitVec<T>::load(row + (long long)nfull * VEC, x);
    float g[8];
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      g[e] = fast_exp2(fmaf(x[e], L2E, -lb)) * gscale;
      if (nfull * VEC + e == label) g[e] -= gscale;
      if (nfull * VEC + e >= V) g[e] = 0.f;
    }
    store(g, nfull * VEC);
  }
  for (long long c = (long long)nvec * VEC + threadIdx.x; c < ld_d; c += blockDim.x) drow[c] = __float2bfloat16(0.f);
}

// ------------------------------------------------------------------------------------------------
// column sums: out[n] += sum_m X[m, n]   (bias gradients = reduce of dY over tokens)
//   block = 32 column groups (8 bf16 = 16 bytes each: one warp = 512 contiguous bytes) x 8 row lanes; grid.y splits the rows.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const __nv_bfloat16* __restrict__ X, long long ld, float* __restrict__ out, long long rows, int N,
                   int rows_per_block) {
  __shared__ float sm[8][264];
  const int cg = threadIdx.x & 31, rl = threadIdx.x >> 5;        // a warp reads 512 contiguous bytes of one row
  const int c0 = blockIdx.x * 256 + cg * 8;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  if (c0 < N) {
#pragma unroll 8
    for (long long r = r0 + rl; r < r1; r += 8) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(X + r * ld + c0));
      const float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y), c = unpack_bf16x2(v.z), d = unpack_bf16x2(v.w);
      acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y; acc[4] += c.x; acc[5] += c.y; acc[6] += d.x; acc[7] += d.y;
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) sm[rl][cg * 8 + e] = acc[e];
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < N) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += sm[k][threadIdx.x];
    atomicAdd(out + c, t);
  }
}

// ------------------------------------------------------------------------------------------------
// K8 fused Adam over the flat fp32 arena, TF form + bf16 shadow refresh
//   m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= alpha * m / (sqrt(v) + eps) ; shadow = bf16(p)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ g,
            __nv_bfloat16* __restrict__ shadow, long long n4, float alpha, float b1, float b2, float eps,
            const float* __restrict__ alpha_dev) {
  if (alpha_dev != nullptr) alpha = *alpha_dev;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pv = reinterpret_cast<float4*>(p)[i], mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    const float4 gv = reinterpret_cast<const float4*>(g)[i];
#define BX_ADAM1(c)                                       \
  mv.c = b1 * mv.c + (1.f - b1) * gv.c;                   \
  vv.c = b2 * vv.c + (1.f - b2) * gv.c * gv.c;            \
  pv.c -= alpha * mv.c / (sqrtf(vv.c) + eps);
    BX_ADAM1(x) BX_ADAM1(y) BX_ADAM1(z) BX_ADAM1(w)
#undef BX_ADAM1
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (shadow) reinterpret_cast<uint2*>(shadow)[i] = make_uint2(pack_bf16x2(pv.x, pv.y), pack_bf16x2(pv.z, pv.w));
  }
}

// per-step scalars derived ON THE DEVICE from the iteration counter, so a captured CUDA graph can be replayed:
//   state[0] = keras `optimizer.iterations` (0-based; incremented here, i.e. once per training step)
//   out: seed = base_seed + iterations ; lr = CustomSchedule(iterations) (train.py:21-34, fp32) or const_lr ;
//        alpha = lr * sqrt(1 - b2^t) / (1 - b1^t), t = iterations + 1      (TF-form Adam)
__global__ void step_params_kernel(long long* state, unsigned long long base_seed, float d_model, float warmup, float const_lr,
                                   float b1, float b2, unsigned long long* seed_out, float* alpha_out, float* lr_out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const long long it = state[0];
  float lr = const_lr;
  if (const_lr < 0.f) {
    const float step = (float)it;
    const float arg1 = 1.0f / sqrtf(step);                                  // rsqrt(0) = inf on the first step
    const float arg2 = step * (float)pow((double)warmup, -1.5);
    lr = (1.0f / sqrtf(d_model)) * fminf(arg1, arg2);
  }
  const double t = (double)(it + 1);
  const double alpha = (double)lr * sqrt(1.0 - pow((double)b2, t)) / (1.0 - pow((double)b1, t));
  *seed_out = base_seed + (unsigned long long)it;
  *alpha_out = (float)alpha;
  *lr_out = lr;
  state[0] = it + 1;
}

__global__ void __launch_bounds__(256)
cast_f32_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(in)[i];
    reinterpret_cast<uint2*>(out)[i] = make_uint2(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w));
  }
}
__global__ void __launch_bounds__(256)
cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __bfloat162float(in[i]);
}
// dQ accumulator fp32 [rows, d] -> column block [col0, col0+d) of a bf16 [rows, ld] buffer
__global__ void __launch_bounds__(256)
cast_f32_bf16_strided_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long rows, int d,
                             long long ld_out) {
  const int d4 = d >> 2;
  const long long total = rows * d4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / d4;
    const int c4 = (int)(i - r * d4);
    const float4 a = reinterpret_cast<const float4*>(in)[i];
    *reinterpret_cast<uint2*>(out + r * ld_out + c4 * 4) = make_uint2(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w));
  }
}

// key-padding bytes: 1 where token id == 0   (positionalencoding.py:25-30, compact form)
__global__ void pad_mask_kernel(const long long* __restrict__ ids, uint8_t* __restrict__ mask, long long n, int S,
                                int id_stride, int id_offset) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    mask[i] = ids[(i / S) * id_stride + id_offset + (i % S)] == 0 ? 1 : 0;
}

// attention backward pre-pass: delta[b,h,q] = sum_c dO[b,q,h*dk+c] * O[b,q,h*dk+c]
//   one thread per (b, h, q), q fastest: coalesced stores; each thread streams its own 2*dk-byte rows (16-byte loads)
__global__ void __launch_bounds__(256)
attn_delta_kernel(const __nv_bfloat16* __restrict__ O, const __nv_bfloat16* __restrict__ O_lo, long long ld_o,
                  const __nv_bfloat16* __restrict__ dO, long long ld_do, float* __restrict__ delta, int B, int S, int H, int dk) {
  // dk/8 adjacent lanes share one (row, head): every lane reads 16 contiguous bytes of O and dO, so a warp reads
  // 512 contiguous bytes of each; a sub-warp shuffle finishes the dot product
  const int lph = dk >> 3;                                              // lanes per head: 4 (dk 32), 8 (dk 64) or 16 (dk 128)
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long total = (long long)B * S * H * lph;
  const bool live = idx < total;
  const long long i = live ? idx : total - 1;
  const int l = (int)(i % lph);
  const int h = (int)((i / lph) % H);
  const long long row = i / ((long long)lph * H);                       // b * S + q
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(O + row * ld_o + h * dk + l * 8));
  const uint4 d = __ldg(reinterpret_cast<const uint4*>(dO + row * ld_do + h * dk + l * 8));
  float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), a2 = unpack_bf16x2(a.z), a3 = unpack_bf16x2(a.w);
  if (O_lo != nullptr) {          // + the rounding residual the forward kept: O to ~16 mantissa bits (see attn_fwd_kernel)
    const uint4 r = __ldg(reinterpret_cast<const uint4*>(O_lo + row * ld_o + h * dk + l * 8));
    const float2 r0 = unpack_bf16x2(r.x), r1 = unpack_bf16x2(r.y), r2 = unpack_bf16x2(r.z), r3 = unpack_bf16x2(r.w);
    a0.x += r0.x; a0.y += r0.y; a1.x += r1.x; a1.y += r1.y; a2.x += r2.x; a2.y += r2.y; a3.x += r3.x; a3.y += r3.y;
  }
  const float2 d0 = unpack_bf16x2(d.x), d1 = unpack_bf16x2(d.y), d2 = unpack_bf16x2(d.z), d3 = unpack_bf16x2(d.w);
  float s = a0.x * d0.x + a0.y * d0.y + a1.x * d1.x + a1.y * d1.y + a2.x * d2.x + a2.y * d2.y + a3.x * d3.x + a3.y * d3.y;
  for (int o = lph >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (live && l == 0) {
    const long long b = row / S;
    const int q = (int)(row % S);
    delta[(b * H + h) * S + q] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
static inline int grid_for(long long work, int threads, int max_blocks_per_sm = 8) {
  long long b = (work + threads - 1) / threads;
  const long long cap = (long long)sm_count() * max_blocks_per_sm;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

int embed_fwd(const long long* ids, int id_stride, int id_offset, const float* table, const float* pos, float* out_f32,
              void* out_bf16, long long rows, int S, int d, long long V, float rate, unsigned long long seed, unsigned site,
              cudaStream_t st, const unsigned long long* seed_dev) {
  BX_REQUIRE(d % 4 == 0, "embed: d_model %% 4 != 0");
  embed_fwd_kernel<<<grid_for(rows * (d / 4), 256), 256, 0, st>>>(ids, table, pos, out_f32, (__nv_bfloat16*)out_bf16, rows, S,
                                                                 id_stride, id_offset, d, V, sqrtf((float)d), rate, seed, site, seed_dev);
  BX_POST_LAUNCH();
  return 0;
}
int embed_bwd(const long long* ids, int id_stride, int id_offset, const void* g1, const void* g2, float* dtable, long long rows,
              int S, int d, long long V, float rate, unsigned long long seed, unsigned site, cudaStream_t st,
              const unsigned long long* seed_dev) {
  BX_REQUIRE(d % 4 == 0, "embed: d_model %% 4 != 0");
  embed_bwd_kernel<<<grid_for(rows * (d / 4), 256), 256, 0, st>>>(ids, (const __nv_bfloat16*)g1, (const __nv_bfloat16*)g2, dtable, rows,
                                                                 S, id_stride, id_offset, d, V, sqrtf((float)d), rate, seed, site,
                                                                 seed_dev);
  BX_POST_LAUNCH();
  return 0;
}

#define BX_LN_DISPATCH(NAME, ...)                                         \
  if (d <= 256) NAME<1 BX_LN_EXTRA> __VA_ARGS__;                          \
  else if (d <= 512) NAME<2 BX_LN_EXTRA> __VA_ARGS__;                     \
  else NAME<4 BX_LN_EXTRA> __VA_ARGS__;

int residual_ln_fwd(const void* y, const void* x, const float* gamma, const float* beta, void* z, void* out, float* mean,
                    float* rstd, void* keep_bits, long long rows, int d, float eps, float rate, unsigned long long seed, unsigned site,
                    cudaStream_t st, const unsigned long long* seed_dev) {
  BX_REQUIRE(d % 8 == 0 && d <= 1024, "layernorm: need d %% 8 == 0 and d <= 1024 (got %d)", d);
  const int grid = grid_for(rows * 32, 256);
#define BX_LN_EXTRA
  BX_LN_DISPATCH(residual_ln_fwd_kernel, <<<grid, 256, 0, st>>>((const __nv_bfloat16*)y, (const __nv_bfloat16*)x, gamma, beta,
                                                               (__nv_bfloat16*)z, (__nv_bfloat16*)out, mean, rstd, (uint8_t*)keep_bits,
                                                               rows, d, eps, rate, seed, site, seed_dev))
#undef BX_LN_EXTRA
  BX_POST_LAUNCH();
  return 0;
}
template <int NV, bool F32, bool REGEN, bool FULL>
static int ln_bwd_go(const void* g1, const void* g2, const void* z, const float* mean, const float* rstd, const float* gamma, void* dz, void* dy_sub,
                     float* dgamma, float* dbeta, float* dbias, const void* keep_bits, long long rows, int d, float rate, unsigned long long seed,
                     unsigned site, cudaStream_t st, const unsigned long long* seed_dev, int grid, int threads, size_t smem) {
  auto kern = residual_ln_bwd_kernel<NV, F32, REGEN, FULL>;
  static std::mutex mu;
  static bool attr_done[kMaxDevices] = {};
  {
    std::lock_guard<std::mutex> lock(mu);
    const int dev = current_device();
    if (!attr_done[dev]) {
      BX_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      attr_done[dev] = true;
    }
  }
  kern<<<grid, threads, smem, st>>>(g1, (const __nv_bfloat16*)g2, (const __nv_bfloat16*)z, mean, rstd, gamma, (__nv_bfloat16*)dz,
                                    (__nv_bfloat16*)dy_sub, dgamma, dbeta, dbias, (const uint8_t*)keep_bits, rows, d, rate, seed, site, seed_dev);
  return 0;
}

int residual_ln_bwd(const void* g1, int g1_f32, const void* g2, const void* z, const float* mean, const float* rstd,
                    const float* gamma, void* dz, void* dy_sub, float* dgamma, float* dbeta, float* dbias, const void* keep_bits,
                    long long rows, int d, float rate, unsigned long long seed, unsigned site, cudaStream_t st,
                    const unsigned long long* seed_dev) {
  BX_REQUIRE(d % 8 == 0 && d <= 1024, "layernorm: need d %% 8 == 0 and d <= 1024 (got %d)", d);
  const bool regen = rate > 0.f && keep_bits == nullptr;
  const int threads = (d <= 512 && !regen) ? 512 : 256;
  const int grid = grid_for(rows * 32, threads, 1);
  const size_t smem = (size_t)(threads / 32) * kLnBwdStages * ((size_t)d * (g1_f32 ? 4 : 2) + 4 * (size_t)d);      // >= [warps][3][d] floats
  BX_REQUIRE(smem <= 227 * 1024, "layernorm backward: input ring of %zu bytes does not fit", smem);
#define BX_LN_BWD_GO(NV, F32, REGEN) \
  rc = (d == 256 * NV) ? ln_bwd_go<NV, F32, REGEN, true>(g1, g2, z, mean, rstd, gamma, dz, dy_sub, dgamma, dbeta, dbias, keep_bits, rows, d, rate, seed, site, st, seed_dev, grid, threads, smem) \
                       : ln_bwd_go<NV, F32, REGEN, false>(g1, g2, z, mean, rstd, gamma, dz, dy_sub, dgamma, dbeta, dbias, keep_bits, rows, d, rate, seed, site, st, seed_dev, grid, threads, smem)
#define BX_LN_BWD_NV(F32, REGEN) \
  do { if (d <= 256) BX_LN_BWD_GO(1, F32, REGEN); else if (d <= 512) BX_LN_BWD_GO(2, F32, REGEN); else BX_LN_BWD_GO(4, F32, REGEN); } while (0)
  int rc = 0;
  if (g1_f32 && regen) BX_LN_BWD_NV(true, true);
  else if (g1_f32) BX_LN_BWD_NV(true, false);
  else if (regen) BX_LN_BWD_NV(false, true);
  else BX_LN_BWD_NV(false, false);
#undef BX_LN_BWD_NV
#undef BX_LN_BWD_GO
  if (rc) return rc;
  BX_POST_LAUNCH();
  return 0;
}

int xent(const void* logits, int logits_f32, long long ld, const long long* labels, int label_stride, int label_offset,
         int T_len, void* dlogits, long long ld_d, float* row_loss, double* stats, long long rows, int V, float inv_batch,
         cudaStream_t st, const int* t_valid) {
  BX_REQUIRE(rows > 0 && rows < (1ll << 31), "xent: bad row count");
  if (logits_f32)
    xent_kernel<float><<<(unsigned)rows, 256, 0, st>>>((const float*)logits, ld, labels, label_stride, label_offset, T_len,
                                                       (__nv_bfloat16*)dlogits, ld_d, row_loss, stats, V, inv_batch, t_valid);
  else
    xent_kernel<__nv_bfloat16><<<(unsigned)rows, 256, 0, st>>>((const __nv_bfloat16*)logits, ld, labels, label_stride,
                                                               label_offset, T_len, (__nv_bfloat16*)dlogits, ld_d, row_loss,
                                                               stats, V, inv_batch, t_valid);
  BX_POST_LAUNCH();
  return 0;
}

int colsum_bf16(const void* X, long long ld, float* out, long long rows, int N, cudaStream_t st) {
  BX_REQUIRE(ld % 8 == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0, "colsum: X must be 16-byte aligned with ld %% 8 == 0");
  const int bx_ = (N + 255) / 256;
  long long by = (long long)sm_count() * 6 / bx_;
  if (by < 1) by = 1;
  long long rpb = (rows + by - 1) / by;
  rpb = (rpb + 63) / 64 * 64;
  by = (rows + rpb - 1) / rpb;
  colsum_bf16_kernel<<<dim3(bx_, (unsigned)by), 256, 0, st>>>((const __nv_bfloat16*)X, ld, out, rows, N, (int)rpb);
  BX_POST_LAUNCH();
  return 0;
}

int step_params(long long* state, unsigned long long base_seed, float d_model, float warmup, float const_lr, float b1, float b2,
                unsigned long long* seed_out, float* alpha_out, float* lr_out, cudaStream_t st) {
  step_params_kernel<<<1, 32, 0, st>>>(state, base_seed, d_model, warmup, const_lr, b1, b2, seed_out, alpha_out, lr_out);
  BX_POST_LAUNCH();
  return 0;
}
int adam_step(float* p, float* m, float* v, const float* g, void* shadow, long long n, float alpha, float b1, float b2,
              float eps, cudaStream_t st, const float* alpha_dev) {
  BX_REQUIRE(n % 4 == 0, "adam: arena length must be a multiple of 4");
  adam_kernel<<<grid_for(n / 4, 256), 256, 0, st>>>(p, m, v, g, (__nv_bfloat16*)shadow, n / 4, alpha, b1, b2, eps, alpha_dev);
  BX_POST_LAUNCH();
  return 0;
}
int cast_f32_bf16(const float* in, void* out, long long n, cudaStream_t st) {
  BX_REQUIRE(n % 4 == 0, "cast: length must be a multiple of 4");
  cast_f32_bf16_kernel<<<grid_for(n / 4, 256), 256, 0, st>>>(in, (__nv_bfloat16*)out, n / 4);
  BX_POST_LAUNCH();
  return 0;
}
int cast_bf16_f32(const void* in, float* out, long long n, cudaStream_t st) {
  cast_bf16_f32_kernel<<<grid_for(n, 256), 256, 0, st>>>((const __nv_bfloat16*)in, out, n);
  BX_POST_LAUNCH();
  return 0;
}
int cast_f32_bf16_strided(const float* in, void* out, long long rows, int d, long long ld_out, cudaStream_t st) {
  BX_REQUIRE(d % 4 == 0 && ld_out % 4 == 0, "cast: d and ld must be multiples of 4");
  cast_f32_bf16_strided_kernel<<<grid_for(rows * (d / 4), 256), 256, 0, st>>>(in, (__nv_bfloat16*)out, rows, d, ld_out);
  BX_POST_LAUNCH();
  return 0;
}
int pad_mask(const long long* ids, int id_stride, int id_offset, uint8_t* mask, long long rows, int S, cudaStream_t st) {
  pad_mask_kernel<<<grid_for(rows * S, 256), 256, 0, st>>>(ids, mask, rows * S, S, id_stride, id_offset);
  BX_POST_LAUNCH();
  return 0;
}
int attn_delta(const void* O, long long ld_o, const void* dO, long long ld_do, float* delta, int B, int S, int H, int dk,
               cudaStream_t st, const void* O_lo) {
  BX_REQUIRE((dk == 32 || dk == 64 || dk == 128) && ld_o % 8 == 0 && ld_do % 8 == 0, "attn_delta: depth must be 32, 64 or 128, leading dimensions multiples of 8");
  const long long n = (long long)B * S * H * (dk / 8);
  attn_delta_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const __nv_bfloat16*)O, (const __nv_bfloat16*)O_lo, ld_o,
                                                                 (const __nv_bfloat16*)dO, ld_do, delta, B, S, H, dk);
  BX_POST_LAUNCH();
  return 0;
}

}  // namespace bx
