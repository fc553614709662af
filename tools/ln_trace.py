"""Per-warp clock trace of the fused Dense + LayerNorm kernel (library built with -DBX_LN2_TRACE): blocks 0 and 1 (one cluster)."""
import ctypes, os, sys
import numpy as np
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
for _ in range(3):
    L.call("b200xf_gemm_bias_dropout_residual_ln", L.ptr(A), K, L.ptr(W), N, L.ptr(bias), L.ptr(x), L.ptr(gamma), L.ptr(beta), L.ptr(z), L.ptr(out),
           L.ptr(mean), L.ptr(rstd), L.ptr(keep), M, N, K, 1e-6, rate, 1, 2, L.stream_ptr())
    torch.cuda.synchronize()
buf = np.zeros((2, 20, 64), dtype=np.int64)
lib = L.load()
lib.b200xf_debug_ln2_trace.argtypes = [ctypes.c_void_p]
rc = lib.b200xf_debug_ln2_trace(buf.ctypes.data)
assert rc == 0
print(f"K {K} rate {rate}")
for blk in range(2):
    t0 = buf[blk, 1, 0]
    m = buf[blk, 1]
    print(f"block {blk} MMA warp (per row block: start, accumulator free, first stage full, issued):")
    for i in range(0, 16, 4):
        print("   ", [int(v - t0) for v in m[i:i + 4]])
    for w in (4, 8, 12, 16):
        e = buf[blk, w]
        print(f"block {blk} epilogue warp {w}: events 0..14 per row block (0 start, 1 acc ready, 2 ld, 3 bias/drop, 4 x, 5 ld, 6 bias/drop, 7 -, 8 z written, 9 z store, 10 wg sums, 11 peer, 12 box free, 13 out written, 14 barrier)")
        for i in range(0, 60, 15):
            print("   ", [int(v - t0) for v in e[i:i + 15]])
