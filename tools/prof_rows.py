"""One launch of each HBM-bound row pass at the bench shape (for `ncu --set full -k regex:...`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from llamax_b200 import ops

torch.manual_seed(0)
M, D, F = 16384, 4096, 14336
dev = "cuda"
x = torch.randn(M, D, device=dev).bfloat16()
dy = torch.randn(M, D, device=dev).bfloat16()
res = torch.randn(M, D, device=dev).bfloat16()
w = torch.ones(D, device=dev).bfloat16()
ab = torch.randn(M, 2 * F, device=dev).bfloat16()
_, rstd, _, _ = ops.rmsnorm_fwd(x, w, 1e-5, quant=False)
for _ in range(int(os.environ.get("REPS", "2"))):
    ops.rowquant_int8(x)
    ops.rmsnorm_fwd(x, w, 1e-5, quant=True)
    ops.rmsnorm_bwd(dy, x, w, rstd, res, want_dw=True)
    ops.swiglu_fwd(ab[:, :F], ab[:, F:], quant=True, want_g=True)
torch.cuda.synchronize()
print("ok")
