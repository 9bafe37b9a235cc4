#!/bin/bash
# r2r: GPU tests of the staged-history pitch phase, A/B against the previous build (h1 = same sources without it),
# and a per-instruction profile with the caches left warm (--cache-control none: the r2p profile flushed them before
# every replay, which overstated the table loads)
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > $O/r2r_tests.log
cat $O/r2r_tests.log
L=rnnoise_b200/librnnoise_b200
REPS=2 AB_STEPS=600 bash tools/ab_libs.sh h1:${L}_h1.so new:${L}.so > $O/r2r_ab_4096.txt 2>&1
cat $O/r2r_ab_4096.txt
RNNOISE_B200_OVERLAP=0 ncu --set full --cache-control none --clock-control none --import-source on --launch-skip 60 --launch-count 14 -f -o $O/r2r_full \
    python bench.py --steps 2 --warmup 8 --no-cpu-baseline > $O/r2r_ncu_full.log 2>&1
ls -la $O/r2r_full.ncu-rep
ncu -i $O/r2r_full.ncu-rep --page raw --csv > $O/r2r_full_raw.csv 2>/dev/null
