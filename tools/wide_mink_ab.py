"""A/B of LLAMAX_GEMM_WIDE_MINK on the plain short-K bf16 GEMM shapes of the step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from llamax_b200 import ops
from tools.gemm_vs_cublas import sustained
for (M, N, K) in ((16384, 4096, 6168), (8192, 128256, 4096), (16384, 4096, 4096)):
    Kp = (K + 63) // 64 * 64
    a = torch.randn(M, Kp, device="cuda").bfloat16()[:, :K]
    b = torch.randn(N, Kp, device="cuda").bfloat16()[:, :K]
    ts, clk = sustained(lambda: ops.bf16_gemm(a, b), 1.5)
    print("WIDE_MINK", os.environ.get("LLAMAX_GEMM_WIDE_MINK"), f"[{M},{N},{K}] {2.0*M*N*K/ts/1e9:.0f} TF/s [{clk}]", flush=True)
    del a, b
