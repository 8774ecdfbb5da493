#!/bin/bash
# Same-box A/B of an environment switch:  [BENCH_ARGS=...] tools/ab.sh VAR A_VALUE B_VALUE  (alternates A B A B)
VAR=$1; A=$2; B=$3
for i in 1 2; do
  for v in "$A" "$B"; do
    env $VAR=$v python bench.py --no-cpu-baseline --steps 4 $BENCH_ARGS 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['kernels']
print('$VAR=$v', 'ms/step', d['ms_per_step'], 'e2e_ms', d['e2e']['ms_per_step'], 'bf16', k['bf16_gemm']['ms_per_step'], 'int8', k['int8_gemm']['ms_per_step'], 'lora', k['lora_wgrad']['ms_per_step'], 'attn_fwd', k['attn_fwd']['ms_per_step'], 'attn_bwd', k['attn_bwd']['ms_per_step'], 'sm', d['clocks']['sm_mhz'])
"
  done
done
