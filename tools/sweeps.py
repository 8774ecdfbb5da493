"""BASELINE.json configs[3] and configs[4] as measured sweeps (evidence under profiles/, not bench lines):
  * INT8 row-scaled GEMM over the Llama-8B projection shapes, M = 1k..32k: forward (int8 tcgen05 + dequant epilogue)
    next to cuBLASLt torch._int_mm (accumulators only, no dequant), and grad_input (bf16 tcgen05 over (s*W)^T)
  * prefix-LM attention, B*S = 16384 tokens, S = 1k..16k, prefix fraction 0..75 %, GQA 32/8, head_dim 128, fwd + bwd
CUDA events, best of 5 after 2 warm-ups; operands of every case exceed L2 or are re-created per case."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from llamax_b200 import ops

dev = "cuda"


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


print("== INT8 row-scaled GEMM sweep (TOP/s = 2*M*N*K / time); nominal int8 peak 4500, measured bf16 sustained 1401.6 TF/s")
print(f"{'N x K':>14s} {'M':>6s} | {'fwd int8+dequant':>17s} {'% nominal':>9s} {'cuBLASLt _int_mm':>17s} | {'grad_input bf16':>16s} {'% measured':>10s}")
for (N, K) in [(4096, 4096), (14336, 4096), (4096, 14336), (6144, 4096)]:
    w8 = torch.randint(-127, 128, (N, K), device=dev, dtype=torch.int8)
    ws = torch.rand(N, device=dev).bfloat16()
    wt = (torch.randn(K, N, device=dev) * 0.02).bfloat16()           # (s*W)^T operand of grad_input
    for M in (1024, 2048, 4096, 8192, 16384, 32768):
        xq = torch.randint(-127, 128, (M, K), device=dev, dtype=torch.int8)
        xs = torch.rand(M, device=dev).bfloat16()
        t_f = timeit(lambda: ops.int8_gemm_dequant(xq, w8, xs, ws))
        t_c = timeit(lambda: torch._int_mm(xq, w8.t()))
        dy = torch.randn(M, N, device=dev).bfloat16()
        t_g = timeit(lambda: ops.bf16_gemm(dy, wt))
        ops_ = 2.0 * M * N * K
        print(f"{N:>6d} x {K:<6d} {M:>6d} | {ops_ / t_f / 1e9:14.0f} T/s {100 * ops_ / t_f / 1e9 / 4500:8.1f}% {ops_ / t_c / 1e9:14.0f} T/s | "
              f"{ops_ / t_g / 1e9:13.0f} T/s {100 * ops_ / t_g / 1e9 / 1401.6:9.1f}%", flush=True)
        del xq, dy
    del w8, wt

print()
print("== prefix-LM attention sweep (B*S = 16384, Hq 32 / Hkv 8, head_dim 128; TFLOP/s over UNMASKED pairs only)")
print(f"{'S':>6s} {'B':>3s} {'prefix':>7s} | {'fwd ms':>8s} {'TF/s':>6s} | {'bwd ms':>8s} {'TF/s':>6s}")
Hq, Hkv, D = 32, 8, 128
for S in (1024, 2048, 4096, 8192, 16384):
    B = 16384 // S
    ld = (Hq + 2 * Hkv) * D
    g = torch.randn(B * S, ld, device=dev).bfloat16()
    q, k, v = g[:, : Hq * D], g[:, Hq * D : (Hq + Hkv) * D], g[:, (Hq + Hkv) * D :]
    dout = torch.randn(B * S, Hq * D, device=dev).bfloat16()
    dqkv = torch.empty_like(g)
    dq, dk, dv = dqkv[:, : Hq * D], dqkv[:, Hq * D : (Hq + Hkv) * D], dqkv[:, (Hq + Hkv) * D :]
    for frac in (0.0, 0.25, 0.5, 0.75):
        P = int(S * frac)
        o, lse = ops.attn_fwd(q, k, v, B, S, Hq, Hkv, D, P)
        pairs = S * P + (S - P) * (S - P + 1) / 2
        fl = 4.0 * B * Hq * D * pairs
        tf = timeit(lambda: ops.attn_fwd(q, k, v, B, S, Hq, Hkv, D, P))
        tb = timeit(lambda: ops.attn_bwd(q, k, v, o, lse, dout, dq, dk, dv, B, S, Hq, Hkv, D, P))
        print(f"{S:>6d} {B:>3d} {P:>7d} | {tf:8.3f} {fl / tf / 1e9:6.0f} | {tb:8.3f} {2.5 * fl / tb / 1e9:6.0f}", flush=True)
    del g, dout, dqkv
print("ok")
