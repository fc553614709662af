"""Micro-benchmark of the tcgen05 GEMM on the FFN / projection shapes of transformer-base (run under gpurun)."""
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transformer_b200 import _lib as L


def bench(M, N, K, a_major, b_major, out_f32, flags=0, splits=1, bias=True, iters=20):
    dev = torch.device("cuda:0")
    A = torch.randn((M, K) if a_major == 0 else (K, M), device=dev).bfloat16()
    B = torch.randn((N, K) if b_major == 0 else (K, N), device=dev).bfloat16()
    D = torch.zeros(M, N, device=dev, dtype=torch.float32 if out_f32 else torch.bfloat16)
    bias_t = torch.randn(N, device=dev) if bias else None
    lda = A.shape[1]; ldb = B.shape[1]
    sp = L.stream_ptr()
    def run():
        L.call("b200xf_gemm_bf16", L.ptr(A), lda, a_major, L.ptr(B), ldb, b_major, L.ptr(D), N, int(out_f32), L.ptr(bias_t), None, 0,
               flags, splits, M, N, K, sp)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 2.0 * M * N * K / ms / 1e9
    # cuBLAS for context only
    At = A if a_major == 0 else A.t(); Bt = B.t() if b_major == 0 else B
    for _ in range(3):
        torch.matmul(At, Bt)
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters):
        torch.matmul(At, Bt)
    e1.record(); torch.cuda.synchronize()
    ms_ref = e0.elapsed_time(e1) / iters
    return dict(M=M, N=N, K=K, majors=f"{a_major}{b_major}", out="f32" if out_f32 else "bf16", splits=splits, ms=round(ms, 4),
                tflops=round(tf, 1), cublas_tflops=round(2.0 * M * N * K / ms_ref / 1e9, 1))


if __name__ == "__main__":
    if len(sys.argv) > 1:
        L.load().b200xf_gemm_force_cta_group(int(sys.argv[1]))
    M = 32768
    rows = [
        bench(M, 2048, 512, 0, 1, False, flags=1),            # FFN dense0 fwd (bias+relu)
        bench(M, 512, 2048, 0, 1, True),                      # FFN dense1 fwd
        bench(M, 1536, 512, 0, 1, False),                     # QKV projection
        bench(M, 512, 512, 0, 1, True),                       # out projection
        bench(M, 2048, 512, 0, 0, False, bias=False),         # dgrad dense1
        bench(M, 512, 2048, 0, 0, True, flags=2, bias=False), # dgrad dense0 (+residual accumulate)
        bench(512, 2048, M, 1, 1, True, flags=2, splits=8, bias=False),   # wgrad dense0
        bench(2048, 512, M, 1, 1, True, flags=2, splits=4, bias=False),   # wgrad dense1
        bench(M, 32768, 512, 0, 1, False),                    # vocab projection
        bench(8192, 8192, 8192, 0, 0, False, bias=False),     # square reference point
    ]
    for r in rows:
        print(json.dumps(r))
