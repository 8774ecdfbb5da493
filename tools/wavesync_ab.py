import os, sys
sys.path.insert(0, "/root/repo")
import torch
from llamax_b200 import ops
from tools.gemm_vs_cublas import burst, sustained
M,N,K=16384,4096,28688
Kp=(K+63)//64*64
a=torch.randn(M,Kp,device="cuda").bfloat16()[:,:K]; b=torch.randn(N,Kp,device="cuda").bfloat16()[:,:K]
ref=torch.matmul(a.float()[:256], b.float().t())
out=ops.bf16_gemm(a,b)
print("check", (out[:256].float()-ref).abs().max().item()/ref.abs().max().item())
fl=2.0*M*N*K
ts,clk=sustained(lambda: ops.bf16_gemm(a,b), 2.0)
print("wavesync", os.environ.get("LLAMAX_GEMM_WAVESYNC"), f"sustained {fl/ts/1e9:.0f} TF/s [{clk}]")
