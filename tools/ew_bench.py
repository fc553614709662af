"""Micro-benchmark of the HBM-bound kernels at transformer-base shapes (32768 tokens, d 512, V 32768): each line gives
the kernel's time, its algorithmic bytes and the fraction of the measured HBM peak (MEASURED_PEAKS.json) reached."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from transformer_b200 import _lib as L

try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6551.0


def timeit(fn, iters=10, flush=None):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        if flush is not None:
            flush.add_(1)                      # evict L2 between launches (buffer > 126 MB)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters


def main():
    dev = torch.device("cuda:0")
    rows, d, V, B, S, H = 32768, 512, 32768, 64, 512, 8
    sp = L.stream_ptr()
    flush = torch.zeros(64 << 20, device=dev)   # 256 MB
    out = []
    f = lambda *s: torch.randn(*s, device=dev)
    h16 = lambda *s: torch.randn(*s, device=dev).bfloat16()
    y, x, g, bta = h16(rows, d), h16(rows, d), f(d), f(d)
    z, o16 = torch.empty_like(y), torch.empty_like(y)
    mean, rstd = torch.empty(rows, device=dev), torch.empty(rows, device=dev)
    keep = torch.zeros(rows, d // 8, dtype=torch.uint8, device=dev)

    def ln_fwd():
        L.call("b200xf_residual_ln_fwd", L.ptr(y), L.ptr(x), L.ptr(g), L.ptr(bta), L.ptr(z), L.ptr(o16), L.ptr(mean), L.ptr(rstd),
               L.ptr(keep), rows, d, 1e-6, 0.1, 1, 3, sp)
    out.append(("residual_ln_fwd", timeit(ln_fwd, flush=flush), rows * d * 2 * 4))
    g1, g2 = h16(rows, d), h16(rows, d)
    dz = torch.empty_like(y); dy = torch.empty_like(y)
    dg, db, dbias = torch.zeros(d, device=dev), torch.zeros(d, device=dev), torch.zeros(d, device=dev)

    def ln_bwd():
        L.call("b200xf_residual_ln_bwd", L.ptr(g1), 0, L.ptr(g2), L.ptr(z), L.ptr(mean), L.ptr(rstd), L.ptr(g), L.ptr(dz), L.ptr(dy),
               L.ptr(dg), L.ptr(db), L.ptr(dbias), L.ptr(keep), rows, d, 0.1, 1, 3, sp)
    out.append(("residual_ln_bwd", timeit(ln_bwd, flush=flush), rows * d * 2 * 5))

    logits = (torch.randn(rows, V, device=dev) * 2).bfloat16()
    tar = torch.randint(1, V, (B, S + 1), device=dev)
    stats = torch.zeros(4, dtype=torch.float64, device=dev)

    def xent():
        L.call("b200xf_xent", L.ptr(logits), 0, V, L.ptr(tar), S + 1, 1, S, L.ptr(logits), V, None, L.ptr(stats), rows, V, 1.0 / B, sp)
    out.append(("xent (in place, bf16)", timeit(xent, iters=5), rows * V * 4))
    V2 = V + 8                                   # > 32768 columns: the two-pass fallback kernel, for comparison
    logits2 = (torch.randn(rows, V2, device=dev) * 2).bfloat16()

    def xent2():
        L.call("b200xf_xent", L.ptr(logits2), 0, V2, L.ptr(tar), S + 1, 1, S, L.ptr(logits2), V2, None, L.ptr(stats), rows, V2, 1.0 / B, sp)
    out.append(("xent two-pass fallback", timeit(xent2, iters=5), rows * V2 * 4))
    del logits2
    cs = torch.zeros(V, device=dev)
    out.append(("colsum_bf16 vocab", timeit(lambda: L.call("b200xf_colsum_bf16", L.ptr(logits), V, L.ptr(cs), rows, V, sp), iters=5), rows * V * 2))
    del logits
    h = torch.randn(rows, 2048, device=dev).bfloat16()
    cs2 = torch.zeros(2048, device=dev)
    out.append(("colsum_bf16 dff", timeit(lambda: L.call("b200xf_colsum_bf16", L.ptr(h), 2048, L.ptr(cs2), rows, 2048, sp), flush=flush), rows * 2048 * 2))
    O = torch.randn(rows, d, device=dev).bfloat16(); dO = torch.randn(rows, d, device=dev).bfloat16()
    delta = torch.empty(B, H, S, device=dev)
    out.append(("attn_delta", timeit(lambda: L.call("b200xf_attn_delta", L.ptr(O), d, L.ptr(dO), d, L.ptr(delta), B, S, H, 64, sp), flush=flush),
                rows * d * 4 + rows * H * 4))
    for name, ms, byt in out:
        gbps = byt / ms / 1e6
        print(json.dumps(dict(kernel=name, us=round(ms * 1e3, 1), algorithmic_MB=round(byt / 1e6, 1), GBps=round(gbps), frac_of_hbm_peak=round(gbps / PEAK, 3))))


if __name__ == "__main__":
    main()
