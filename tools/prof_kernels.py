"""One launch of each hot kernel at the bench shapes (Llama-3.1-8B, M = 8 x 2048 tokens) — the command profiled by
`ncu --set full` for profiles/ (keeps the capture small; bench.py itself supplies the launch list)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from llamax_b200 import ops

torch.manual_seed(0)
M, D, F, Hq, Hkv, hd, B, S = 16384, 4096, 14336, 32, 8, 128, 8, 2048
dev = "cuda"
reps = int(os.environ.get("REPS", "2"))
for _ in range(reps):
    # 1. bf16 grad_input GEMM of the w1|w3 group: [M, 2F + 16] x [D, 2F + 16]^T
    dab = torch.randn(M, 2 * F + 16, device=dev).bfloat16()
    wt = torch.randn(D, 2 * F + 16, device=dev).bfloat16()
    ops.bf16_gemm(dab, wt)
    del dab, wt
    # 2. int8 forward GEMM of w1 with dequant + LoRA epilogue
    xq = torch.randint(-127, 128, (M, D), device=dev, dtype=torch.int8)
    w8 = torch.randint(-127, 128, (F, D), device=dev, dtype=torch.int8)
    xs, ws = torch.rand(M, device=dev).bfloat16(), torch.rand(F, device=dev).bfloat16()
    h, lb = torch.randn(M, 16, device=dev).bfloat16(), torch.randn(F, 8, device=dev).bfloat16()
    ops.int8_gemm_dequant(xq, w8, xs, ws, lora_h=h[:, :8], lora_b=lb, lora_scale=1.0)
    # 3. int8 forward GEMM of w2 with residual
    gq = torch.randint(-127, 128, (M, F), device=dev, dtype=torch.int8)
    w2 = torch.randint(-127, 128, (D, F), device=dev, dtype=torch.int8)
    res = torch.randn(M, D, device=dev).bfloat16()
    ops.int8_gemm_dequant(gq, w2, xs, ws[:D].contiguous(), lora_h=h[:, :8], lora_b=lb[:D].contiguous(), resid=res)
    del gq, w2
    # 4/5. attention forward / backward (causal)
    qkv = torch.randn(M, (Hq + 2 * Hkv) * hd, device=dev).bfloat16()
    q, k, v = qkv[:, : Hq * hd], qkv[:, Hq * hd : (Hq + Hkv) * hd], qkv[:, (Hq + Hkv) * hd :]
    o, lse = ops.attn_fwd(q, k, v, B, S, Hq, Hkv, hd, 0)
    dqkv = torch.empty_like(qkv)
    ops.attn_bwd(q, k, v, o, lse, torch.randn_like(o), dqkv[:, : Hq * hd], dqkv[:, Hq * hd : (Hq + Hkv) * hd],
                 dqkv[:, (Hq + Hkv) * hd :], B, S, Hq, Hkv, hd, 0)
    # 6. LoRA weight gradient, 7/8. fused norm+quant, SwiGLU+quant
    ab = torch.randn(M, 2 * F, device=dev).bfloat16()
    ops.lora_wgrad(ab[:, :F], h[:, :8], 1.0)
    x = torch.randn(M, D, device=dev).bfloat16()
    w1 = torch.ones(D, device=dev).bfloat16()
    xn, rstd, _, _ = ops.rmsnorm_fwd(x, w1, 1e-5, quant=True)
    ops.swiglu_fwd(ab[:, :F], ab[:, F:], quant=True, want_g=True)
    ops.rowquant_int8(x)                      # attention output -> wo operand
    # 9. fused LoRA backward pair (dh + dB in one pass over dY), 10. RMSNorm backward
    bt = torch.randn(8, F, device=dev).bfloat16()
    ops.lora_bwd_pair(ab[:, :F], bt, h[:, :8], torch.empty(M, 8, device=dev).bfloat16(), 1.0)
    ops.rmsnorm_bwd(xn, x, w1, rstd, res, want_dw=True)
    # 11. w2 grad_input GEMM [M, D + LoRA] -> [M, F] with the SwiGLU backward as its epilogue (dg never written)
    wt2 = torch.randn(F, D, device=dev).bfloat16()
    dab2 = torch.empty(M, 2 * F + 16, device=dev, dtype=torch.bfloat16)
    ops.bf16_gemm_swiglu_bwd(x, wt2, ab[:, :F], ab[:, F:], out_ab=dab2, want_g=True, lora_h=h[:, :8], lora_b=lb)
    del wt2, dab2
    # 12/13. mixed-input GEMMs (opt-in mode): grad_input of wo from the int8 weight as stored, weight-only forward of wq
    w8o = torch.randint(-127, 128, (D, D), device=dev, dtype=torch.int8)
    so = torch.rand(D, device=dev).bfloat16()
    ops.bf16_int8_gemm_bwd(x, w8o, so)
    ops.bf16_int8_gemm(x, w8o, so)
    del w8o
    del ab, qkv, dqkv
torch.cuda.synchronize()
print("ok")
