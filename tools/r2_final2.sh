# final records of round 2, second half (row kernels + grouped LoRA backward): suite, bench line, tables, ncu
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r2_bench_1gpu_final2.json 2> gpurun_out/r2_bench_1gpu_final2.err; echo "bench rc=$?"
timeout 300 python tools/ew_perf.py > gpurun_out/r2_ew_perf_final2.txt 2>&1; echo "ew_perf rc=$?"
timeout 300 python tools/ew_sustained.py 12 > gpurun_out/r2_ew_sustained_final2.txt 2>&1; echo "ew_sustained rc=$?"
bash tools/prof_round2c.sh
