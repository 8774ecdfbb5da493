timeout 300 python tools/sanitize_cases.py > gpurun_out/sanitize_plain.log 2>&1; rc=$?; echo "plain rc=$rc"; tail -3 gpurun_out/sanitize_plain.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_cases.py > gpurun_out/r2_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -8 gpurun_out/r2_sanitizer_memcheck.log
