#!/bin/bash
# r2i: GPU tests of the current tree, then same-box A/B of library builds (tools/ab_libs.sh) at 4096 streams and two
# other batch sizes, then racecheck of the new shared-memory plans.
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > $O/r2i_tests.log
L=rnnoise_b200/librnnoise_b200
REPS=2 AB_STEPS=400 bash tools/ab_libs.sh r2h:${L}_r2h.so c4rot:${L}_c4rot.so new:${L}.so norot:${L}_norot.so noc4:${L}_noc4.so > $O/r2i_ab_4096.txt 2>&1
for S in 1024 16384; do
  for v in r2h new; do
    lib=${L}_$v.so; [ $v = new ] && lib=${L}.so
    RNNOISE_B200_LIB_PATH=$PWD/$lib timeout 300 python bench.py --streams $S --steps 300 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['roofline']['kernel_ms_per_step']
print('$S $v ms/step', round(d['ms_per_step'],4), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), {a: round(b*1e3,1) for a,b in k.items()})" >> $O/r2i_ab_sizes.txt
  done
done
compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitizer_run.py 300 4 > $O/r2i_sanitizer_racecheck.log 2>&1
tail -3 $O/r2i_sanitizer_racecheck.log
cat $O/r2i_tests.log $O/r2i_ab_4096.txt $O/r2i_ab_sizes.txt
