"""HBM-bound passes timed the way the training step runs them: at the power cap, each one right after a long tensor-bound
GEMM (so SM / L2 clocks are the step's, ~1.45 GHz, not the 1.9 GHz an idle part boosts to), CUDA events around the single
launch. Next to them: a plain device copy of the same size under the same conditions — what the memory system itself gives
at those clocks. Usage: python tools/ew_sustained.py [rounds]"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from llamax_b200 import ops

torch.manual_seed(0)
M, D, F, Hq, Hkv, hd, B, S = 16384, 4096, 14336, 32, 8, 128, 8, 2048
dev = "cuda"
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 12

ga = torch.randn(M, 28688, device=dev).bfloat16()
gb = torch.randn(4096, 28688, device=dev).bfloat16()   # [N, K], K-major like every operand of bf16_gemm
gout = torch.empty(M, 4096, device=dev, dtype=torch.bfloat16)


def burn():
    ops.bf16_gemm(ga, gb, out=gout)


x = torch.randn(M, D, device=dev).bfloat16()
dy = torch.randn(M, D, device=dev).bfloat16()
res = torch.randn(M, D, device=dev).bfloat16()
w = torch.ones(D, device=dev).bfloat16()
_, rstd, _, _ = ops.rmsnorm_fwd(x, w, 1e-5, quant=False)
ab = torch.randn(M, 2 * F, device=dev).bfloat16()
qkv = torch.randn(M, (Hq + 2 * Hkv) * hd, device=dev).bfloat16()
from llamax_b200.modelling.llama import LlamaConfig, build_rope  # noqa: E402

rope = build_rope(LlamaConfig(D, 1, hd, Hq, Hkv, F, max_seq_len=4096, vocab_size=1024, rope_base=500000,
                              is_llama3_1=True))[:S].contiguous().to(dev)
h = torch.randn(M, 8, device=dev).bfloat16()
bt = torch.randn(8, F, device=dev).bfloat16()
btd = torch.randn(8, D, device=dev).bfloat16()
dh = torch.empty(M, 8, device=dev).bfloat16()
y = torch.empty_like(x)
big = torch.randn(M, F, device=dev).bfloat16()
big2 = torch.empty_like(big)
i8 = torch.empty(M, D, device=dev, dtype=torch.int8)

cases = [
    ("copy 134 MB (torch copy_)", lambda: y.copy_(x), 4.0 * M * D),
    ("copy 470 MB (torch copy_)", lambda: big2.copy_(big), 4.0 * M * F),
    ("read-only 134 MB (torch sum)", lambda: x.view(torch.int32).sum(), 2.0 * M * D),
    ("rmsnorm_fwd", lambda: ops.rmsnorm_fwd(x, w, 1e-5, quant=False), 4.0 * M * D),
    ("rmsnorm_fwd + quant", lambda: ops.rmsnorm_fwd(x, w, 1e-5, quant=True), 5.0 * M * D),
    ("rmsnorm_bwd (+resid, +dw)", lambda: ops.rmsnorm_bwd(dy, x, w, rstd, res, want_dw=True), 8.0 * M * D),
    ("rowquant_int8", lambda: ops.rowquant_int8(x), 3.0 * M * D),
    ("swiglu_fwd + quant + g", lambda: ops.swiglu_fwd(ab[:, :F], ab[:, F:], quant=True, want_g=True), 7.0 * M * F),
    ("rope (q|k in place)", lambda: ops.rope_(qkv, rope, B, S, Hq + Hkv, hd), 4.0 * M * (Hq + Hkv) * hd),
    ("lora_wgrad [M,F]^T [M,8]", lambda: ops.lora_wgrad(ab[:, :F], h, 1.0), 2.0 * M * (F + 8)),
    ("lora_wgrad [M,D]^T [M,8]", lambda: ops.lora_wgrad(x, h, 1.0), 2.0 * M * (D + 8)),
    ("lora_bwd_pair dY[M,F]", lambda: ops.lora_bwd_pair(ab[:, :F], bt, h, dh, 1.0), 2.0 * M * (F + 16)),
    ("lora_bwd_pair dY[M,D]", lambda: ops.lora_bwd_pair(x, btd, h, dh, 1.0), 2.0 * M * (D + 16)),
    ("lora_bwd_pair dY[M,1024]", lambda: ops.lora_bwd_pair(x[:, :1024], btd[:, :1024], h, dh, 1.0), 2.0 * M * (1024 + 16)),
]
only = os.environ.get("EW_ONLY")
if only:
    cases = [c for c in cases if any(k in c[0] for k in only.split(","))]


def smi():
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip().splitlines()[0]
        return out
    except Exception:
        return "?"


# warm: every case once, then hold the cap for ~2 s
for _, fn, _ in cases:
    fn()
torch.cuda.synchronize()
for _ in range(600):
    burn()
torch.cuda.synchronize()
print("after warm-up burn: SM MHz, W =", smi(), flush=True)

times = {name: [] for name, _, _ in cases}
for r in range(rounds):
    evs = []
    for name, fn, _ in cases:
        for _ in range(6):
            burn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        evs.append((name, e0, e1))
    if r == rounds // 2:
        print("under load (queue full): SM MHz, W =", smi(), flush=True)
    torch.cuda.synchronize()
    for name, e0, e1 in evs:
        times[name].append(e0.elapsed_time(e1))
print(f"{'pass':40s} {'median us':>10s} {'min us':>8s} {'GB/s (median)':>14s}")
for name, _, nbytes in cases:
    ts = sorted(times[name])
    med = ts[len(ts) // 2]
    print(f"  {name:38s} {med * 1e3:10.1f} {ts[0] * 1e3:8.1f} {nbytes / med / 1e6:14.0f}", flush=True)
print("ok")
