"""One call of the fused Dense + LayerNorm kernel (timing build prints its per-phase clocks)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transformer_b200 import _lib as L
dev = torch.device("cuda:0")
M, N = 32768, 512
K = int(sys.argv[1]) if len(sys.argv) > 1 else 512
rate = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
A = torch.randn(M, K, device=dev).bfloat16(); W = (torch.randn(K, N, device=dev) * K ** -0.5).bfloat16()
bias = torch.randn(N, device=dev) * 0.1; x = torch.randn(M, N, device=dev).bfloat16()
gamma = torch.ones(N, device=dev); beta = torch.zeros(N, device=dev)
z = torch.empty(M, N, dtype=torch.bfloat16, device=dev); out = torch.empty_like(z)
mean = torch.empty(M, device=dev); rstd = torch.empty(M, device=dev); keep = torch.zeros(M, N // 8, dtype=torch.uint8, device=dev)
for _ in range(2):
    L.call("b200xf_gemm_bias_dropout_residual_ln", L.ptr(A), K, L.ptr(W), N, L.ptr(bias), L.ptr(x), L.ptr(gamma), L.ptr(beta), L.ptr(z), L.ptr(out),
           L.ptr(mean), L.ptr(rstd), L.ptr(keep), M, N, K, 1e-6, rate, 1, 2, L.stream_ptr())
    torch.cuda.synchronize()
    print("----")
