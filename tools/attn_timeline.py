import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transformer_b200 import _lib as L
dev = torch.device("cuda:0")
B, H, S, dk = 64, 8, 512, 64
d = H * dk
qkv = torch.randn(B * S, 3 * d, device=dev).bfloat16(); O = torch.empty(B * S, d, device=dev, dtype=torch.bfloat16)
dO = torch.randn(B * S, d, device=dev).bfloat16(); lse = torch.empty(B, H, S, device=dev); delta = torch.empty(B, H, S, device=dev)
dq = torch.zeros(B * S, d, device=dev); dqkv = torch.empty(B * S, 3 * d, device=dev, dtype=torch.bfloat16)
kpad = torch.zeros(B, S, dtype=torch.uint8, device=dev); sp = L.stream_ptr()
Q, K, V = qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]
tl = torch.zeros(256, dtype=torch.int64, device=dev)
SEL = int(sys.argv[1]) if len(sys.argv) > 1 else -1      # blockIdx.x of the traced CTA (-1: last)
for causal in (0, 1):
    for it in range(2):
        L.call("b200xf_debug_set_timeline", L.ptr(tl))
        tl.zero_(); tl[255] = SEL
        L.call("b200xf_attn_fwd", L.ptr(Q), 3 * d, L.ptr(K), 3 * d, L.ptr(V), 3 * d, L.ptr(O), d, L.ptr(lse), L.ptr(kpad), causal, None, 0, 0, B, H, S, S, dk, sp)
        torch.cuda.synchronize()
        f = tl.cpu().tolist()
        tl.zero_(); tl[255] = SEL
        L.call("b200xf_attn_delta", L.ptr(O), d, L.ptr(dO), d, L.ptr(delta), B, S, H, dk, sp)
        L.call("b200xf_attn_bwd", L.ptr(Q), 3 * d, L.ptr(K), 3 * d, L.ptr(V), 3 * d, L.ptr(dO), d, L.ptr(lse), L.ptr(delta), L.ptr(dqkv), 3 * d, 1,
               L.ptr(dqkv[:, d:2 * d]), 3 * d, L.ptr(dqkv[:, 2 * d:]), 3 * d, L.ptr(kpad), causal, None, 0, 0, B, H, S, S, dk, sp)
        torch.cuda.synchronize()
        bw = tl.cpu().tolist()
    t0 = f[0]
    print(f"FWD causal={causal}: alloc+init {f[1]-t0}; per tile [wait_S, pass1, pass2, (fence), wait_O, accO]; final {f[2]-t0} .. stored {f[3]-t0}")
    for j in range(4):
        e = f[8 + j * 8: 8 + j * 8 + 6]
        if e[0]: print(f"   tile {j}: start {e[0]-t0:6d} | waitS {e[1]-e[0]:5d} load+max {e[2]-e[1]:5d} exp+store {e[3]-e[2]:5d} fence {e[4]-e[3]:4d}")
    t0 = bw[60]
    print(f"BWD causal={causal}: epilogue start {bw[61]-t0} end {bw[62]-t0}")
    for n in range(4):
        e = bw[64 + n * 8: 64 + n * 8 + 8]
        if e[0]: print(f"   tile {n}: start {e[0]-t0:6d} | waitSdP {e[2]-e[0]:5d} compute+st {e[3]-e[2]:5d} | dq_full(n) seen at {e[4]-t0:6d} reduce issued {e[5]-t0:6d}")
L.call("b200xf_debug_set_timeline", None)
