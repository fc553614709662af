"""residual + dropout + LayerNorm backward kernel alone at the transformer-base shape: time and achieved HBM rate."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transformer_b200 import _lib as L
dev = torch.device("cuda:0")
M, d, rate = 32768, 512, 0.1
g1 = torch.randn(M, d, device=dev).bfloat16(); g2 = torch.randn(M, d, device=dev).bfloat16(); z = torch.randn(M, d, device=dev).bfloat16()
mean = z.float().mean(-1); rstd = (z.float().var(-1, unbiased=False) + 1e-6).rsqrt()
gamma = torch.ones(d, device=dev); dz = torch.empty_like(z); dy = torch.empty_like(z)
dg = torch.zeros(d, device=dev); db = torch.zeros(d, device=dev); dbias = torch.zeros(d, device=dev)
keep = torch.randint(0, 256, (M, d // 8), dtype=torch.uint8, device=dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
sp = L.stream_ptr()
for g2p, name in ((L.ptr(g2), "two gradient streams"), (None, "one gradient stream")):
    def run():
        L.call("b200xf_residual_ln_bwd", L.ptr(g1), 0, g2p, L.ptr(z), L.ptr(mean), L.ptr(rstd), L.ptr(gamma), L.ptr(dz), L.ptr(dy), L.ptr(dg), L.ptr(db),
               L.ptr(dbias), L.ptr(keep), M, d, rate, 1, 2, sp)
    ts = []
    for _ in range(8):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts = sorted(ts)[1:-1]
    us = sum(ts) / len(ts)
    nbytes = M * d * 2 * ((3 if g2p else 2) + 2) + M * d // 8
    print(json.dumps({"case": name, "us": round(us, 1), "GB/s": round(nbytes / us / 1e3, 1), "bytes": nbytes}))
