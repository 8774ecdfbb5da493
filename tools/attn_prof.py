"""Small driver for ncu: one forward + backward attention call at the bench shape (B=8, S=2048, Hq=32, Hkv=8)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from llamax_b200 import ops
B, S, Hq, Hkv, D, P = 8, 2048, 32, 8, 128, 0
ld = (Hq + 2 * Hkv) * D
g = torch.randn(B * S, ld, device="cuda").bfloat16()
q, k, v = g[:, : Hq * D], g[:, Hq * D : (Hq + Hkv) * D], g[:, (Hq + Hkv) * D :]
dout = torch.randn(B * S, Hq * D, device="cuda").bfloat16()
dqkv = torch.empty_like(g)
dq, dk, dv = dqkv[:, : Hq * D], dqkv[:, Hq * D : (Hq + Hkv) * D], dqkv[:, (Hq + Hkv) * D :]
for _ in range(2):
    o, lse = ops.attn_fwd(q, k, v, B, S, Hq, Hkv, D, P)
    ops.attn_bwd(q, k, v, o, lse, dout, dq, dk, dv, B, S, Hq, Hkv, D, P)
torch.cuda.synchronize()
print("ok")
