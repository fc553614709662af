import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transformer_b200 import _lib as L
cg = int(sys.argv[1]); M, N, K = (int(x) for x in sys.argv[2:5])
L.load().b200xf_gemm_force_cta_group(cg)
dev = torch.device("cuda:0")
A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16()
D = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
for _ in range(3):
    L.call("b200xf_gemm_bf16", L.ptr(A), K, 0, L.ptr(B), K, 0, L.ptr(D), N, 0, None, None, 0, 0, 1, M, N, K, L.stream_ptr())
torch.cuda.synchronize()
