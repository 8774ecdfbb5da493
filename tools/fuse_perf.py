"""Isolated timing of the w2 grad_input GEMM, the SwiGLU backward and the fused GEMM + SwiGLU-backward epilogue at the
8B shape (M = 16384, F = 14336, K = 4096, LoRA r = 8).  usage: python tools/fuse_perf.py [iters]"""
import sys

import torch

sys.path.insert(0, ".")
from llamax_b200 import ops  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
M, F, K, R = 16384, 14336, 4096, 8
dev = "cuda"
dy = torch.randn(M, K, device=dev).bfloat16()
wt = (torch.randn(F, K, device=dev) * 0.02).bfloat16()
ab = torch.randn(M, 2 * F, device=dev).bfloat16()
a, b = ab[:, :F], ab[:, F:]
h = torch.randn(M, R, device=dev).bfloat16()
lb = (torch.randn(F, R, device=dev) * 0.1).bfloat16()
dab = torch.empty(M, 2 * F + 16, device=dev, dtype=torch.bfloat16)
dg = torch.empty(M, F, device=dev, dtype=torch.bfloat16)


def timed(fn, name):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{name:34s} {ms * 1e3:8.1f} us   {2.0 * M * F * K / ms / 1e9:7.1f} TFLOP/s-equivalent")
    return ms


for rep in range(2):
    t_g = timed(lambda: ops.bf16_gemm(dy, wt, out=dg, lora_h=h, lora_b=lb), "bf16_gemm (rank-8 epilogue)")
    t_s = timed(lambda: ops.swiglu_bwd(dg, a, b, want_g=True, out_ab=dab), "swiglu_bwd (+g)")
    t_f = timed(lambda: ops.bf16_gemm_swiglu_bwd(dy, wt, a, b, out_ab=dab, want_g=True, lora_h=h, lora_b=lb),
                "fused gemm + swiglu_bwd epilogue")
    print(f"  two launches {1e3 * (t_g + t_s):.1f} us vs fused {1e3 * t_f:.1f} us")
