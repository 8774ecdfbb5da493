# kernel trace of one step of the text workload with device time per kernel name
timeout 600 python tools/gap_profile.py 32 text > gpurun_out/r2_gap_profile.txt 2>&1; echo "rc=$?"; grep -v Warning gpurun_out/r2_gap_profile.txt | head -50
