"""Fused Adam (+ bf16 shadow) over a transformer-base sized arena (63 M parameters)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transformer_b200 import _lib as L
dev = torch.device("cuda:0")
n = 63_000_000 // 4 * 4
p = torch.randn(n, device=dev); m = torch.zeros(n, device=dev); v = torch.zeros(n, device=dev); g = torch.randn(n, device=dev) * 1e-3
sh = torch.empty(n, dtype=torch.bfloat16, device=dev)
ts = []
for _ in range(6):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    L.call("b200xf_adam_step", L.ptr(p), L.ptr(m), L.ptr(v), L.ptr(g), L.ptr(sh), n, 1e-3, 0.9, 0.98, 1e-9, L.stream_ptr())
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
us = sorted(ts)[len(ts) // 2]
print(json.dumps({"params": n, "us": round(us, 1), "GB/s": round(n * 30 / us / 1e3, 1)}))
