"""INT8 GEMM at the forward shapes with and without the LoRA / residual epilogue terms (graph replays): how much of
the kernel time is epilogue.  usage: python tools/int8_epi_perf.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from llamax_b200 import ops

M = 16384
dev = "cuda"


def timeit(fn, n=5, reps=6):
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


for N, K in ((14336, 4096), (4096, 4096), (1024, 4096), (4096, 14336)):
    a = torch.randint(-127, 128, (M, K), device=dev, dtype=torch.int8)
    w = torch.randint(-127, 128, (N, K), device=dev, dtype=torch.int8)
    sa, sw = torch.rand(M, device=dev).bfloat16(), torch.rand(N, device=dev).bfloat16()
    h = torch.randn(M, 8, device=dev).bfloat16()
    lb = torch.randn(N, 8, device=dev).bfloat16()
    res = torch.randn(M, N, device=dev).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    ops_ = 2.0 * M * N * K
    t0 = timeit(lambda: ops.int8_gemm_dequant(a, w, sa, sw, out=out))
    t1 = timeit(lambda: ops.int8_gemm_dequant(a, w, sa, sw, out=out, lora_h=h, lora_b=lb, lora_scale=1.0))
    t2 = timeit(lambda: ops.int8_gemm_dequant(a, w, sa, sw, out=out, lora_h=h, lora_b=lb, lora_scale=1.0, resid=res))
    t3 = timeit(lambda: torch._int_mm(a, w.t()))
    print(f"N={N:5d} K={K:5d}: dequant only {t0*1e3:7.1f} us {ops_/t0/1e9:6.0f} TOP/s | +LoRA r8 {t1*1e3:7.1f} us "
          f"{ops_/t1/1e9:6.0f} | +LoRA+resid {t2*1e3:7.1f} us {ops_/t2/1e9:6.0f} | cuBLASLt _int_mm (int32 out) "
          f"{t3*1e3:7.1f} us {ops_/t3/1e9:6.0f}", flush=True)
