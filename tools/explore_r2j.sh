#!/bin/bash
# r2j: same-box A/B of the occupancy choices of k_spectrum / k_synthesis on top of the chain4 build without rotation
O=gpurun_out; mkdir -p $O
L=rnnoise_b200/librnnoise_b200
REPS=2 AB_STEPS=400 bash tools/ab_libs.sh c4rot:${L}_c4rot.so b0:${L}_b0.so b1:${L}_b1.so b2:${L}_b2.so b3:${L}_b3.so > $O/r2j_ab_4096.txt 2>&1
for S in 1024 16384; do
  for v in b0 b1 b3; do
    RNNOISE_B200_LIB_PATH=$PWD/${L}_$v.so timeout 300 python bench.py --streams $S --steps 300 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['roofline']['kernel_ms_per_step']
print('$S $v ms/step', round(d['ms_per_step'],4), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), {a: round(b*1e3,1) for a,b in k.items()})" >> $O/r2j_ab_sizes.txt
  done
done
cat $O/r2j_ab_4096.txt $O/r2j_ab_sizes.txt
