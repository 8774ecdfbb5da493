# ncu launch list + full capture of the hot kernels at HEAD (same recipe as tools/prof_round2c.sh)
bash tools/prof_round2c.sh
