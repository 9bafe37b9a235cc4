#!/bin/bash
# A/B of compile-time knobs on the GPU box: tools/ab.sh name1:"-DX=1" name2:"-DX=2 -DY=3" ...
# Rebuilds the library for each variant, runs the device-pointer bench, writes gpurun_out/ab_<name>.json.
# The LAST variant stays built, so put the default last.
mkdir -p gpurun_out
for v in "$@"; do
  name="${v%%:*}"; flags="${v#*:}"
  RNNOISE_B200_NVCC_FLAGS="$flags" python rnnoise_b200/build.py --force >/dev/null 2>gpurun_out/ab_${name}.build.log || { echo "$name: build failed"; continue; }
  timeout 300 python bench.py --steps ${AB_STEPS:-200} --warmup 20 --no-cpu-baseline > gpurun_out/ab_${name}.json 2>gpurun_out/ab_${name}.err
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/ab_{n}.json").read().strip().splitlines()[-1])
    k = {a: b * 1e3 for a, b in d["roofline"]["kernel_ms_per_step"].items()}
    print(n, "ms/step", round(d["ms_per_step"], 4), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "e2e16", round(d["e2e"]["int16_pcm"]["value"]), {a: round(b, 1) for a, b in k.items()})
except Exception as e:
    print(n, "FAILED", e)
PY
done
