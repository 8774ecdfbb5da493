"""Per-k-block SM-clock timeline of the mixed-input GEMM's MMA issuer and one converter warp (cluster 0, leader CTA).
Needs the -DLX_MIX_TRACE build:  LLAMAX_B200_LIB=llamax_b200/csrc/libllamax_b200_mixtrace.so python tools/mixed_gemm_trace.py"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from llamax_b200 import _lib, ops

M, N, K = 16384, 4096, 4096
mode = sys.argv[1] if len(sys.argv) > 1 else "bwd"
if mode == "bwd":
    dy = torch.randn(M, K, device="cuda").bfloat16()
    W8 = torch.randint(-127, 128, (K, N), device="cuda", dtype=torch.int8)
    s = torch.rand(K, device="cuda").bfloat16()
    fn = lambda: ops.bf16_int8_gemm_bwd(dy, W8, s)
else:
    x = torch.randn(M, K, device="cuda").bfloat16()
    W8 = torch.randint(-127, 128, (N, K), device="cuda", dtype=torch.int8)
    s = torch.rand(N, device="cuda").bfloat16()
    fn = lambda: ops.bf16_int8_gemm(x, W8, s)
for _ in range(3):
    fn()
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (2 * 64 * 8))()
lib = _lib.load()
lib.llamax_debug_mix_trace.argtypes = [ctypes.c_void_p]
assert lib.llamax_debug_mix_trace(buf) == 0
t = torch.tensor(list(buf)).view(2, 64, 8)
t0 = int(t[0, 0, 0])
print(f"== {mode}: MMA issuer (cycles since its first mark): k-block | top | A full | B converted | issued+committed | period")
prev = None
for i in range(40):
    r = [int(v) - t0 for v in t[0, i, :4]]
    print(f"{i:3d} " + " ".join(f"{v:8d}" for v in r) + (f" {r[0] - prev:7d}" if prev is not None else ""))
    prev = r[0]
print("== converter warp 12 (every 4th k-block): k-block | top | raw full | B slot free | stores issued | fence done | arrived | period")
prev = None
for i in range(20):
    r = [int(v) - t0 for v in t[1, i, :6]]
    print(f"{4 * i:3d} " + " ".join(f"{v:8d}" for v in r) + (f" {r[0] - prev:7d}" if prev is not None else ""))
    prev = r[0]
