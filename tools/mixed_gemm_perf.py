"""Mixed-input GEMM (bf16 x int8 expanded in shared memory) against the same GEMM on a de-quantised bf16 operand, at the
step's shapes, burst and sustained (power-capped), same box.   usage: python tools/mixed_gemm_perf.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from llamax_b200 import ops
from tools.gemm_vs_cublas import burst, sustained

dev = "cuda"
M = 16384
QUICK = len(sys.argv) > 1 and sys.argv[1] == "quick"   # two shapes only, no equality check
# grad_input shapes (N = in_features, K1 = sum of out_features, tail = LoRA rows): wqkv, wo, w1|w3, w2
for name, N, K1, Rt in (("wqkv-bwd", 4096, 6144, 24), ("wo-bwd", 4096, 4096, 0), ("w13-bwd", 4096, 28672, 16), ("w2-bwd", 14336, 4096, 0)):
    if QUICK and name != "wo-bwd":
        continue
    K = K1 + Rt
    Kp = (K + 63) // 64 * 64
    dy = torch.randn(M, Kp, device=dev).bfloat16()[:, :K]
    W8 = torch.randint(-127, 128, (K1, N), device=dev, dtype=torch.int8)
    s = (torch.rand(K1, device=dev) * 0.01 + 1e-3).bfloat16()
    tail = (torch.randn(Rt, N, device=dev) * 0.05).bfloat16() if Rt else None
    wt = torch.zeros(N, Kp, device=dev, dtype=torch.bfloat16)[:, :K]
    ops.dequant_weight(W8, s, transpose=True, apply_scale=True, out=wt[:, :K1])
    if Rt:
        wt[:, K1:].copy_(tail.t())
    fl = 2.0 * M * N * K
    for nm, fn in (("bf16 operand", lambda: ops.bf16_gemm(dy, wt)), ("mixed int8", lambda: ops.bf16_int8_gemm_bwd(dy, W8, s, tail=tail))):
        tb = burst(fn)
        ts, clk = sustained(fn, 1.5)
        print(f"{name:9s} [{M},{N},{K}] {nm:13s} burst {fl / tb / 1e9:7.0f}  sustained {fl / ts / 1e9:7.0f} TF/s  [{clk}]", flush=True)
    assert QUICK or torch.equal(ops.bf16_gemm(dy, wt), ops.bf16_int8_gemm_bwd(dy, W8, s, tail=tail))
    del dy, W8, wt
# weight-only forward shapes
for name, N, K in (("w1-fwd", 14336, 4096), ("wq-fwd", 4096, 4096), ("w2-fwd", 4096, 14336)):
    if QUICK and name != "wq-fwd":
        continue
    x = torch.randn(M, K, device=dev).bfloat16()
    W8 = torch.randint(-127, 128, (N, K), device=dev, dtype=torch.int8)
    s = (torch.rand(N, device=dev) * 0.01 + 1e-3).bfloat16()
    wd = ops.dequant_weight(W8, None, transpose=False, apply_scale=False)
    fl = 2.0 * M * N * K

    def two():
        ops.dequant_weight(W8, None, transpose=False, apply_scale=False, out=wd)
        return ops.bf16_gemm(x, wd, col_scale=s, round_before_scale=True)

    for nm, fn in (("dequant+gemm", two), ("mixed int8", lambda: ops.bf16_int8_gemm(x, W8, s))):
        tb = burst(fn)
        ts, clk = sustained(fn, 1.5)
        print(f"{name:9s} [{M},{N},{K}] {nm:13s} burst {fl / tb / 1e9:7.0f}  sustained {fl / ts / 1e9:7.0f} TF/s  [{clk}]", flush=True)
    del x, W8, wd
