for s in 0 1 2 4; do echo "== LLAMAX_LORA_PAIR_SPLITS=$s"; EW_ONLY=lora,copy LLAMAX_LORA_PAIR_SPLITS=$s timeout 300 python tools/ew_sustained.py 10 | grep -E "lora|copy 134|load"; done
cat > /tmp/prof_lora.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from llamax_b200 import ops
M, D, F = 16384, 4096, 14336
x = torch.randn(M, D, device="cuda").bfloat16(); h = torch.randn(M, 8, device="cuda").bfloat16()
btd = torch.randn(8, D, device="cuda").bfloat16(); dh = torch.empty(M, 8, device="cuda").bfloat16()
for _ in range(2):
    ops.lora_bwd_pair(x, btd, h, dh, 1.0); ops.lora_wgrad(x, h, 1.0)
torch.cuda.synchronize(); print("ok")
PY
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base function -k regex:'lora_bwd_pair_kernel|lora_wgrad_tc_kernel' -s 2 -c 2 -o gpurun_out/prof_lora python /tmp/prof_lora.py > gpurun_out/ncu_lora.log 2>&1; echo "rc=$?"
