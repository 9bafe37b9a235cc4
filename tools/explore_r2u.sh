#!/bin/bash
# r2u: the round's last GPU minutes -- full GPU test suite on the final default build (the r2t "pf" configuration)
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/r2u_tests.log
cat $O/r2u_tests.log
python bench.py --steps 600 --warmup 30 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'])" | tee $O/r2u_bench_short.txt
