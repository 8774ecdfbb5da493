# re-entry check of HEAD (container re-created): suite, smoke, default bench line, reference arm
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/r2_bench_1gpu_reentry.json 2> gpurun_out/r2_bench_1gpu_reentry.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2_bench_1gpu_reentry.json
