# both arms after the shared config builder: the two `config` dicts must be equal
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_ref_cfg.json 2> gpurun_out/r2_ref_cfg.err; echo "ref rc=$?"
timeout 600 python bench.py > gpurun_out/r2_bench_1gpu_final3.json 2> gpurun_out/r2_bench_1gpu_final3.err; echo "bench rc=$?"
python - <<'PY'
import json
a=json.loads([l for l in open('gpurun_out/r2_bench_1gpu_final3.json') if l.startswith('{')][-1])
b=json.loads([l for l in open('gpurun_out/r2_ref_cfg.json') if l.startswith('{')][-1])
print("same config:", a['config']==b['config'], "| same metric/unit:", a['metric']==b['metric'], a['unit']==b['unit'])
print(a['value'], a['ms_per_step'], a['e2e']['value'], a['clocks'], '| ref', b['value'], b['ms_per_step'], b['cpu_baseline']['cores'])
print(a['lm_head_rows'][:60])
PY
