#!/bin/bash
# r2q: GPU tests of the phase-latency changes (synthesis output phase, spectrum follower / band sums / table staging,
# packed GRU epilogue parameters, two-stage weight ring), A/B against the previous build (b2) and the h-prefetch variant
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > $O/r2q_tests.log
cat $O/r2q_tests.log
L=rnnoise_b200/librnnoise_b200
REPS=2 AB_STEPS=600 bash tools/ab_libs.sh b2:${L}_b2.so h1:${L}_h1.so new:${L}.so > $O/r2q_ab_4096.txt 2>&1
cat $O/r2q_ab_4096.txt
for S in 256 1024 16384; do
  for v in b2 new; do
    lib=${L}_$v.so; [ $v = new ] && lib=${L}.so
    RNNOISE_B200_LIB_PATH=$PWD/$lib timeout 300 python bench.py --streams $S --steps 300 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['roofline']['kernel_ms_per_step']
print('$S $v ms/step', round(d['ms_per_step'],4), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), {a: round(b*1e3,1) for a,b in k.items()})" >> $O/r2q_ab_sizes.txt
  done
done
cat $O/r2q_ab_sizes.txt
