#!/bin/bash
# A/B of two builds of the library on the same box: tools/ab_libs.sh name1:path1.so name2:path2.so ...
# (each given lib is benchmarked in turn, the whole list REPS times; paths relative to the repo root)
mkdir -p gpurun_out
for rep in $(seq 1 ${REPS:-2}); do
  for v in "$@"; do
    name="${v%%:*}"; lib="${v#*:}"
    RNNOISE_B200_LIB_PATH="$PWD/$lib" timeout 300 python bench.py --steps ${AB_STEPS:-200} --warmup 20 --no-cpu-baseline > gpurun_out/abl_${name}.json 2>gpurun_out/abl_${name}.err
    python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/abl_{n}.json").read().strip().splitlines()[-1])
    k = {a: b * 1e3 for a, b in d["roofline"]["kernel_ms_per_step"].items()}
    print(n, "ms/step", round(d["ms_per_step"], 4), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), {a: round(b, 1) for a, b in k.items()})
except Exception as e:
    print(n, "FAILED", e)
PY
  done
done
