# kernel trace of one step of the speech prefix-LM workload (per-kernel device time incl. the torch / cuFFT front-end)
timeout 600 python tools/gap_profile.py 32 speech > gpurun_out/r2_gap_profile_speech.txt 2>&1; echo "rc=$?"; grep -v Warning gpurun_out/r2_gap_profile_speech.txt | head -52
