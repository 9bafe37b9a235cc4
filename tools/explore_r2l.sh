#!/bin/bash
# r2l: state check of HEAD after the container was re-created: GPU tests + default bench line
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > $O/r2l_tests.log
python bench.py > $O/r2l_bench.json 2> $O/r2l_bench.err
cat $O/r2l_tests.log; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2l_bench.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'])
print({a: round(b*1e3,1) for a,b in d['roofline']['kernel_ms_per_step'].items()})
PY
