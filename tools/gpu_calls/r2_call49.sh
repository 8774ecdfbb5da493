# where the step is not inside a kernel: host enqueue time and CUPTI idle intervals
timeout 600 python tools/gap_profile.py 32 > gpurun_out/r2_gap_profile.txt 2>&1; echo "rc=$?"; tail -40 gpurun_out/r2_gap_profile.txt
