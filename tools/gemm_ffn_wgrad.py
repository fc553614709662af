"""One FFN dense0 weight-gradient GEMM at transformer-base shapes (dW1[512, 2048] += x^T[512, 32768] . dh[32768, 2048], split-K, fp32
TMA reduce-add): the launch whose DRAM traffic bench.py's roofline.traffic quotes (captured with ncu --set full)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transformer_b200 import _lib as L
dev = torch.device("cuda:0")
M, d, f = 32768, 512, 2048
x = torch.randn(M, d, device=dev).bfloat16(); dh = torch.randn(M, f, device=dev).bfloat16()
dW = torch.zeros(d, f, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(4):
    flush.zero_()
    L.call("b200xf_gemm_bf16", L.ptr(x), d, 1, L.ptr(dh), f, 1, L.ptr(dW), f, 1, None, None, 0, 2, 0, d, f, M, L.stream_ptr())
torch.cuda.synchronize()
