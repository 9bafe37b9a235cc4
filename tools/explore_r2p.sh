#!/bin/bash
# r2p: source-level profile (stall sampling per SASS line) of one frame's kernels at 4096 streams, serialised pass
O=gpurun_out; mkdir -p $O
RNNOISE_B200_OVERLAP=0 ncu --set full --clock-control none --import-source on --launch-skip 60 --launch-count 14 -f -o $O/r2p_full \
    python bench.py --steps 2 --warmup 8 --no-cpu-baseline > $O/r2p_ncu_full.log 2>&1
ls -la $O/r2p_full.ncu-rep
ncu -i $O/r2p_full.ncu-rep --page raw --csv > $O/r2p_full_raw.csv 2>/dev/null
tail -3 $O/r2p_ncu_full.log
L=rnnoise_b200/librnnoise_b200
REPS=2 AB_STEPS=600 bash tools/ab_libs.sh ps2:${L}_ps2.so def:${L}.so > $O/r2p_ab_ps2.txt 2>&1
cat $O/r2p_ab_ps2.txt
