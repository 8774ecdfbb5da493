"""Isolated timing of the HBM-bound passes at the bench shapes (M = 16384 tokens, 8B dims): achieved GB/s against the
algorithmic bytes of each pass. CUDA events around a graph replay of 8 calls, best of 5."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from llamax_b200 import ops

torch.manual_seed(0)
M, D, F, Hq, Hkv, hd, B, S = 16384, 4096, 14336, 32, 8, 128, 8, 2048
dev = "cuda"


def timeit(fn, n=5, reps=8):
    """ms per call: `reps` calls captured in one CUDA graph and replayed (best of n). A Python call of these ops costs
    30-60 us of host time (allocations, ctypes): timed one by one, every pass shorter than that reads as ~65 us."""
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / reps)
    return min(ts)


def report(name, ms, nbytes):
    print(f"  {name:34s} {ms * 1e3:8.1f} us  {nbytes / ms / 1e6:7.0f} GB/s", flush=True)


x = torch.randn(M, D, device=dev).bfloat16()
dy = torch.randn(M, D, device=dev).bfloat16()
res = torch.randn(M, D, device=dev).bfloat16()
w = torch.ones(D, device=dev).bfloat16()
xn, rstd, _, _ = ops.rmsnorm_fwd(x, w, 1e-5, quant=False)
report("rmsnorm_fwd", timeit(lambda: ops.rmsnorm_fwd(x, w, 1e-5, quant=False)), 4.0 * M * D)
report("rmsnorm_fwd + quant", timeit(lambda: ops.rmsnorm_fwd(x, w, 1e-5, quant=True)), 5.0 * M * D)
report("rmsnorm_bwd (+resid, +dw)", timeit(lambda: ops.rmsnorm_bwd(dy, x, w, rstd, res, want_dw=True)), 8.0 * M * D)
report("rowquant_int8", timeit(lambda: ops.rowquant_int8(x)), 3.0 * M * D)
ab = torch.randn(M, 2 * F, device=dev).bfloat16()
report("swiglu_fwd + quant + g", timeit(lambda: ops.swiglu_fwd(ab[:, :F], ab[:, F:], quant=True, want_g=True)), 7.0 * M * F)
report("swiglu_fwd + quant", timeit(lambda: ops.swiglu_fwd(ab[:, :F], ab[:, F:], quant=True, want_g=False)), 5.0 * M * F)
dg = torch.randn(M, F, device=dev).bfloat16()
dab = torch.empty(M, 2 * F + 16, device=dev).bfloat16()
report("swiglu_bwd (+g)", timeit(lambda: ops.swiglu_bwd(dg, ab[:, :F], ab[:, F:], want_g=True, out_ab=dab)), 12.0 * M * F)
del dg, dab
qkv = torch.randn(M, (Hq + 2 * Hkv) * hd, device=dev).bfloat16()
from llamax_b200.modelling.llama import LlamaConfig, build_rope  # noqa: E402  (the product's own table builder)
rope = build_rope(LlamaConfig(D, 1, hd, Hq, Hkv, F, max_seq_len=4096, vocab_size=1024, rope_base=500000,
                              is_llama3_1=True))[:S].contiguous().to(dev)
report("rope (q|k in place)", timeit(lambda: ops.rope_(qkv, rope, B, S, Hq + Hkv, hd)), 4.0 * M * (Hq + Hkv) * hd)
h = torch.randn(M, 8, device=dev).bfloat16()
report("lora_wgrad [M,F]^T [M,8]", timeit(lambda: ops.lora_wgrad(ab[:, :F], h, 1.0)), 2.0 * M * (F + 8))
report("lora_wgrad [M,D]^T [M,8]", timeit(lambda: ops.lora_wgrad(x, h, 1.0)), 2.0 * M * (D + 8))
a8 = torch.randn(8, D, device=dev).bfloat16()
report("lora_down x[M,D] @ A^T[8,D]", timeit(lambda: ops.bf16_gemm(x, a8)), 2.0 * M * (D + 8))
a8f = torch.randn(8, F, device=dev).bfloat16()
g = ab[:, :F]
report("lora_down g[M,F] @ A^T[8,F]", timeit(lambda: ops.bf16_gemm(g, a8f)), 2.0 * M * (F + 8))
bt = torch.randn(8, F, device=dev).bfloat16()
dh = torch.empty(M, 8, device=dev).bfloat16()
report("lora_bwd_pair dY[M,F] (dh + dB)", timeit(lambda: ops.lora_bwd_pair(g, bt, h, dh, 1.0)), 2.0 * M * (F + 16))
bt16 = torch.randn(16, 2 * F, device=dev).bfloat16()
h16 = torch.randn(M, 16, device=dev).bfloat16()
dh16 = torch.empty(M, 16, device=dev).bfloat16()
report("lora_bwd_pair dY[M,2F] r16", timeit(lambda: ops.lora_bwd_pair(ab, bt16, h16, dh16, 1.0)), 2.0 * M * (2 * F + 32))
btd = torch.randn(8, D, device=dev).bfloat16()
report("lora_bwd_pair dY[M,D]", timeit(lambda: ops.lora_bwd_pair(x, btd, h, dh, 1.0)), 2.0 * M * (D + 16))
o = torch.randn(M, Hq * hd, device=dev).bfloat16()
dbt = torch.randn(8, F, device=dev).bfloat16()
dh = torch.empty(M, 8, device=dev).bfloat16()
report("lora_bwd_pair dY[M,F] (dh + dB)", timeit(lambda: ops.lora_bwd_pair(g, bt, h, dh, 1.0)), 2.0 * M * (F + 16))
bt16 = torch.randn(16, 2 * F, device=dev).bfloat16()
h16 = torch.randn(M, 16, device=dev).bfloat16()
dh16 = torch.empty(M, 16, device=dev).bfloat16()
report("lora_bwd_pair dY[M,2F] r16", timeit(lambda: ops.lora_bwd_pair(ab, bt16, h16, dh16, 1.0)), 2.0 * M * (2 * F + 32))
btd = torch.randn(8, D, device=dev).bfloat16()
report("lora_bwd_pair dY[M,D]", timeit(lambda: ops.lora_bwd_pair(x, btd, h, dh, 1.0)), 2.0 * M * (D + 16))
o = torch.randn(M, Hq * hd, device=dev).bfloat16()
print("ok")
