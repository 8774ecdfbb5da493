"""A/B of LLAMAX_GEMM_SYNC_EVERY on the 256x256 GEMM shapes of the step (set the env outside)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from llamax_b200 import ops
from tools.gemm_vs_cublas import sustained
M, D, F = 16384, 4096, 14336
x = torch.randn(M, D, device="cuda").bfloat16()
wt2 = torch.randn(F, D, device="cuda").bfloat16()
ab = torch.randn(M, 2 * F, device="cuda").bfloat16()
dab = torch.empty(M, 2 * F + 64, device="cuda", dtype=torch.bfloat16)
h = torch.randn(M, 8, device="cuda").bfloat16(); lb = torch.randn(F, 8, device="cuda").bfloat16()
ts, clk = sustained(lambda: ops.bf16_gemm_swiglu_bwd(x, wt2, ab[:, :F], ab[:, F:], out_ab=dab, want_g=True, lora_h=h, lora_b=lb), 1.5)
print("SYNC_EVERY", os.environ.get("LLAMAX_GEMM_SYNC_EVERY"), f"w2-bwd+swiglu {2.0*M*F*D/ts/1e9:.0f} TF/s [{clk}]")
xq = torch.randint(-127, 128, (M, D), device="cuda", dtype=torch.int8)
w8 = torch.randint(-127, 128, (F, D), device="cuda", dtype=torch.int8)
xs, ws = torch.rand(M, device="cuda").bfloat16(), torch.rand(F, device="cuda").bfloat16()
ts, clk = sustained(lambda: ops.int8_gemm_dequant(xq, w8, xs, ws, lora_h=h, lora_b=lb, lora_scale=1.0), 1.5)
print("SYNC_EVERY", os.environ.get("LLAMAX_GEMM_SYNC_EVERY"), f"int8 w1 fwd {2.0*M*F*D/ts/1e9:.0f} TOP/s [{clk}]")
dy = torch.randn(M, 6208, device="cuda").bfloat16()[:, :6168]; wq = torch.randn(D, 6208, device="cuda").bfloat16()[:, :6168]
ts, clk = sustained(lambda: ops.bf16_gemm(dy, wq), 1.5)
print("SYNC_EVERY", os.environ.get("LLAMAX_GEMM_SYNC_EVERY"), f"wqkv-bwd K=6168 {2.0*M*D*6168/ts/1e9:.0f} TF/s [{clk}]")
