"""Fused Dense + dropout + residual + LayerNorm kernel against the GEMM + stand-alone LayerNorm pair at transformer-base shapes."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transformer_b200 import _lib as L
dev = torch.device("cuda:0")


def bench(M, N, K, rate, iters=10):
    A = torch.randn(M, K, device=dev).bfloat16(); W = (torch.randn(K, N, device=dev) * K ** -0.5).bfloat16()
    bias = torch.randn(N, device=dev) * 0.1; x = torch.randn(M, N, device=dev).bfloat16()
    gamma = torch.ones(N, device=dev); beta = torch.zeros(N, device=dev)
    z = torch.empty(M, N, dtype=torch.bfloat16, device=dev); out = torch.empty_like(z); y = torch.empty_like(z)
    mean = torch.empty(M, device=dev); rstd = torch.empty(M, device=dev); keep = torch.zeros(M, N // 8, dtype=torch.uint8, device=dev)
    sp = L.stream_ptr()

    def fused():
        L.call("b200xf_gemm_bias_dropout_residual_ln", L.ptr(A), K, L.ptr(W), N, L.ptr(bias), L.ptr(x), L.ptr(gamma), L.ptr(beta), L.ptr(z), L.ptr(out),
               L.ptr(mean), L.ptr(rstd), L.ptr(keep), M, N, K, 1e-6, rate, 1, 2, sp)

    def pair():
        L.call("b200xf_gemm_bf16", L.ptr(A), K, 0, L.ptr(W), N, 1, L.ptr(y), N, 0, L.ptr(bias), None, 0, 0, 1, M, N, K, sp)
        L.call("b200xf_residual_ln_fwd", L.ptr(y), L.ptr(x), L.ptr(gamma), L.ptr(beta), L.ptr(z), L.ptr(out), L.ptr(mean), L.ptr(rstd), L.ptr(keep),
               M, N, 1e-6, rate, 1, 2, sp)
    res = {"M": M, "N": N, "K": K, "rate": rate}
    for name, fn in (("fused_us", fused), ("gemm_plus_ln_us", pair)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record(); torch.cuda.synchronize()
        res[name] = round(e0.elapsed_time(e1) / iters * 1e3, 1)
    return res


if __name__ == "__main__":
    for K in (512, 2048):
        for rate in (0.0, 0.1):
            print(json.dumps(bench(32768, 512, K, rate)))
