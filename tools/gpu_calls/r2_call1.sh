# round 2, GPU call 1: regression + microbenchmarks + incumbents + first r2 bench line
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?"
timeout 120 tools/xu_bench > gpurun_out/r2_xu_bench.log 2>&1; echo "xu rc=$?"
timeout 900 python tools/incumbents.py attn gemm peak > gpurun_out/r2_incumbents.log 2>&1; echo "inc rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; echo "bench rc=$?"
timeout 600 python tools/incumbents.py flex > gpurun_out/r2_incumbents_flex.log 2>&1; echo "flex rc=$?"
tail -3 gpurun_out/r2_pytest1.log; cat gpurun_out/r2_xu_bench.log; cat gpurun_out/r2_incumbents.log; cat gpurun_out/r2_incumbents_flex.log | tail -20
