#!/bin/bash
# r2k: GPU tests, A/B of the batched-load build against the previous default (b1), heads tile shapes, racecheck
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > $O/r2k_tests.log
L=rnnoise_b200/librnnoise_b200
REPS=2 AB_STEPS=400 bash tools/ab_libs.sh b1:${L}_b1.so new:${L}.so > $O/r2k_ab_4096.txt 2>&1
for S in 4096 16384 2048; do
  python tools/ab_env.py --streams $S "RNNOISE_B200_HEADS_TILE=32" "RNNOISE_B200_HEADS_TILE=32w" "RNNOISE_B200_HEADS_TILE=16" > $O/r2k_heads_$S.json 2>$O/r2k_heads_$S.err
done
for S in 1024 16384; do
  for v in b1 new; do
    lib=${L}_$v.so; [ $v = new ] && lib=${L}.so
    RNNOISE_B200_LIB_PATH=$PWD/$lib timeout 300 python bench.py --streams $S --steps 300 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['roofline']['kernel_ms_per_step']
print('$S $v ms/step', round(d['ms_per_step'],4), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), {a: round(b*1e3,1) for a,b in k.items()})" >> $O/r2k_ab_sizes.txt
  done
done
compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitizer_run.py 300 4 > $O/r2k_sanitizer_racecheck.log 2>&1
tail -3 $O/r2k_sanitizer_racecheck.log
cat $O/r2k_tests.log $O/r2k_ab_4096.txt $O/r2k_ab_sizes.txt
