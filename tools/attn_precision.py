"""Where does the attention-backward error come from in the near-uniform-attention regime (freshly initialised model, S = 512)?
Variants: delta from the bf16 O the forward wrote (training path) vs delta from an fp32 O; dQ accumulated in bf16 (red.global)
vs fp32 (TMA reduce).  Inputs: q, k small (logit std ~0.1 -> P close to uniform), like a model at initialisation."""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transformer_b200 import _lib as L


def run(S, qk_scale, B=2, H=8, dk=64, causal=0, seed=0, v_common=0.0):
    dev = torch.device("cuda:0")
    d = H * dk
    g = torch.Generator().manual_seed(seed)
    qkv = torch.randn(B * S, 3 * d, generator=g).to(dev)
    qkv[:, :2 * d] *= qk_scale
    if v_common:                 # a component shared by every key's value vector (bias / LayerNorm beta of the producing layer)
        qkv[:, 2 * d:] = 0.3 * qkv[:, 2 * d:] + v_common * torch.randn(1, d, generator=g).to(dev)
    qkv = qkv.bfloat16()
    Q, K, V = qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]
    q4 = Q.double().reshape(B, S, H, dk).permute(0, 2, 1, 3).contiguous().requires_grad_(True)
    k4 = K.double().reshape(B, S, H, dk).permute(0, 2, 1, 3).contiguous().requires_grad_(True)
    v4 = V.double().reshape(B, S, H, dk).permute(0, 2, 1, 3).contiguous().requires_grad_(True)
    logits = (q4 @ k4.transpose(-1, -2)) / math.sqrt(dk)
    if causal:
        logits = logits + (1 - torch.tril(torch.ones(S, S, device=dev, dtype=torch.float64))) * -1e9
    o_ref = torch.softmax(logits, -1) @ v4
    O = torch.empty(B * S, d, dtype=torch.bfloat16, device=dev)
    lse = torch.empty(B, H, S, device=dev)
    sp = L.stream_ptr()
    L.call("b200xf_attn_fwd", L.ptr(Q), 3 * d, L.ptr(K), 3 * d, L.ptr(V), 3 * d, L.ptr(O), d, L.ptr(lse), None, causal, None, 0, 0, B, H, S, S, dk, sp)
    dO = (torch.randn(B * S, d, generator=g).to(dev) * 0.01).bfloat16()
    do4 = dO.double().reshape(B, S, H, dk).permute(0, 2, 1, 3)
    (o_ref * do4).sum().backward()
    delta16 = torch.empty(B, H, S, device=dev)
    L.call("b200xf_attn_delta", L.ptr(O), d, L.ptr(dO), d, L.ptr(delta16), B, S, H, dk, sp)
    delta_exact = (o_ref.detach() * do4).sum(-1).float().contiguous()
    out = {"S": S, "qk_scale": qk_scale, "causal": causal, "v_common": v_common}
    for dname, delta in (("delta_from_bf16_O", delta16), ("delta_from_fp32_O", delta_exact)):
        for qname, bf in (("dq_bf16", 1), ("dq_fp32", 0)):
            dq = torch.zeros(B * S, d, dtype=torch.bfloat16 if bf else torch.float32, device=dev)
            dKV = torch.empty(B * S, 2 * d, dtype=torch.bfloat16, device=dev)
            L.call("b200xf_attn_bwd", L.ptr(Q), 3 * d, L.ptr(K), 3 * d, L.ptr(V), 3 * d, L.ptr(dO), d, L.ptr(lse), L.ptr(delta), L.ptr(dq), d, bf,
                   L.ptr(dKV[:, :d]), 2 * d, L.ptr(dKV[:, d:]), 2 * d, None, causal, None, 0, 0, B, H, S, S, dk, sp)
            torch.cuda.synchronize()
            r = {}
            for nm, got, ref in (("dq", dq, q4.grad), ("dk", dKV[:, :d], k4.grad), ("dv", dKV[:, d:], v4.grad)):
                got4 = got.double().reshape(B, S, H, dk).permute(0, 2, 1, 3)
                r[nm] = round((got4 - ref).norm().item() / ref.norm().item(), 5)
            out[f"{dname}|{qname}"] = r
    return out


if __name__ == "__main__":
    for S in (128, 512, 2048):
        for sc in (1.0, 0.1):
            print(json.dumps(run(S, sc)))
    print(json.dumps(run(512, 0.1, causal=1)))
    for vc in (0.3, 1.0, 3.0):
        print(json.dumps(run(512, 0.1, v_common=vc)))
        print(json.dumps(run(512, 1.0, v_common=vc)))
