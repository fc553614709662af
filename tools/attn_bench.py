"""Micro-benchmark of the attention kernels at transformer-base shapes (B*S = 32768 tokens, 8 heads, dk 64)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transformer_b200 import _lib as L


def bench(B, H, S, dk, causal, iters=10, bwd=True):
    dev = torch.device("cuda:0")
    d = H * dk
    qkv = torch.randn(B * S, 3 * d, device=dev).bfloat16()
    O = torch.empty(B * S, d, device=dev, dtype=torch.bfloat16)
    dO = torch.randn(B * S, d, device=dev).bfloat16()
    lse = torch.empty(B, H, S, device=dev)
    delta = torch.empty(B, H, S, device=dev)
    dq = torch.zeros(B * S, d, device=dev)
    dqkv = torch.empty(B * S, 3 * d, device=dev, dtype=torch.bfloat16)
    kpad = torch.zeros(B, S, dtype=torch.uint8, device=dev)
    sp = L.stream_ptr()
    Q, K, V = qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]

    def fwd():
        L.call("b200xf_attn_fwd", L.ptr(Q), 3 * d, L.ptr(K), 3 * d, L.ptr(V), 3 * d, L.ptr(O), d, L.ptr(lse), L.ptr(kpad), causal, None, 0, 0,
               B, H, S, S, dk, sp)

    def back():
        L.call("b200xf_attn_delta", L.ptr(O), d, L.ptr(dO), d, L.ptr(delta), B, S, H, dk, sp)
        L.call("b200xf_attn_bwd", L.ptr(Q), 3 * d, L.ptr(K), 3 * d, L.ptr(V), 3 * d, L.ptr(dO), d, L.ptr(lse), L.ptr(delta), L.ptr(dqkv), 3 * d, 1,
               L.ptr(dqkv[:, d:2 * d]), 3 * d, L.ptr(dqkv[:, 2 * d:]), 3 * d, L.ptr(kpad), causal, None, 0, 0, B, H, S, S, dk, sp)

    out = {}
    for name, fn in (("fwd", fwd),) + ((("bwd", back),) if bwd else ()):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        fl = 4.0 * B * H * S * S * dk * (1 if name == "fwd" else 2)
        out[name] = dict(ms=round(ms, 4), dense_tflops=round(fl / ms / 1e9, 1))
    return dict(B=B, H=H, S=S, dk=dk, causal=causal, **out)


if __name__ == "__main__":
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    if quick:
        print(json.dumps(bench(64, 8, 512, 64, 0, iters=2)))
        print(json.dumps(bench(64, 8, 512, 64, 1, iters=2)))
    else:
        for S in (128, 256, 512, 1024, 2048):
            for causal in (0, 1):
                print(json.dumps(bench(32768 // S, 8, S, 64, causal)))
