"""Cross-entropy + gradient at the BASELINE shape: 32768 rows x 32768 bf16 logits, in place; then the stand-alone column sums."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transformer_b200 import _lib as L
dev = torch.device("cuda:0")
B, T, V = 64, 512, 32768
lg = (torch.randn(B * T, V, device=dev) * 2).bfloat16()
tar = torch.randint(1, V, (B, T + 1), device=dev)
stats = torch.zeros(4, dtype=torch.float64, device=dev); cs = torch.zeros(V, device=dev)
src = lg.clone()
ts = []
for _ in range(4):
    lg.copy_(src)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e2 = torch.cuda.Event(enable_timing=True)
    e0.record()
    L.call("b200xf_xent", L.ptr(lg), 0, V, L.ptr(tar), T + 1, 1, T, L.ptr(lg), V, None, L.ptr(stats), B * T, V, 1.0 / B, L.stream_ptr())
    e1.record()
    L.call("b200xf_colsum_bf16", L.ptr(lg), V, L.ptr(cs), B * T, V, L.stream_ptr())
    e2.record(); torch.cuda.synchronize()
    ts.append((e0.elapsed_time(e1) * 1e3, e1.elapsed_time(e2) * 1e3))
ts.sort()
x, c = ts[len(ts) // 2]
print(json.dumps({"xent_us": round(x, 1), "xent_GB/s": round(2 * B * T * V * 2 / x / 1e3, 1), "colsum_us": round(c, 1), "colsum_GB/s": round(B * T * V * 2 / c / 1e3, 1)}))
