#!/bin/bash
# tools/gpu_round_final2.sh <tag> -- last call of the round: GPU tests of the two newest builds (new = unified band-term
# layout; pf = new + L2 prefetch of the GRU state rows), A/B of both against prev (the r2s state), then the default bench,
# the launch list and the full ncu capture with the fastest valid one.  Choice -> gpurun_out/<tag>_choice.txt.
tag=${1:-rF}
O=gpurun_out; mkdir -p $O
L=$PWD/rnnoise_b200/librnnoise_b200
python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > $O/${tag}_tests_new.log
RNNOISE_B200_LIB_PATH=${L}_pf.so python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "default_model or other_models or goldens or full_size or tensor_core" 2>&1 | tail -8 > $O/${tag}_tests_pf.log
cat $O/${tag}_tests_new.log $O/${tag}_tests_pf.log
REPS=2 AB_STEPS=600 bash tools/ab_libs.sh prev:rnnoise_b200/librnnoise_b200_prev.so new:rnnoise_b200/librnnoise_b200.so pf:rnnoise_b200/librnnoise_b200_pf.so > $O/${tag}_ab_4096.txt 2>&1
cat $O/${tag}_ab_4096.txt
choice=$(python - "$O/${tag}_tests_new.log" "$O/${tag}_tests_pf.log" "$O/${tag}_ab_4096.txt" <<'PY'
import re, sys
def ok(f):
    t = open(f).read()
    return " passed" in t and " failed" not in t and "error" not in t.lower()
ms = {"prev": [], "new": [], "pf": []}
for line in open(sys.argv[3]):
    m = re.match(r"(prev|new|pf) ms/step ([0-9.]+)", line)
    if m: ms[m.group(1)].append(float(m.group(2)))
mean = lambda v: sum(v) / len(v) if v else 1e9
cand = {"prev": mean(ms["prev"]) * 0.998}          # a newer build has to win by more than the noise
if ok(sys.argv[1]): cand["new"] = mean(ms["new"])
if ok(sys.argv[2]): cand["pf"] = mean(ms["pf"]) * 1.001   # ... and the prefetch variant has to beat `new`
print(min(cand, key=cand.get))
PY
)
echo "choice: $choice" | tee $O/${tag}_choice.txt
if [ "$choice" != new ]; then export RNNOISE_B200_LIB_PATH=${L}_$choice.so; fi
python bench.py > $O/${tag}_bench.json 2> $O/${tag}_bench.err
python bench.py --streams 16384 --steps 300 --warmup 20 --no-cpu-baseline > $O/${tag}_bench_16384.json 2>> $O/${tag}_bench.err
python bench.py --streams 256 --steps 300 --warmup 20 --no-cpu-baseline > $O/${tag}_bench_256.json 2>> $O/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${tag}_ncu_bench.log 2>&1
RNNOISE_B200_OVERLAP=0 ncu --set full --clock-control none --import-source on --launch-skip 60 --launch-count 14 -f -o $O/${tag}_full \
    python bench.py --steps 2 --warmup 8 --no-cpu-baseline > $O/${tag}_ncu_full.log 2>&1
ncu -i $O/${tag}_full.ncu-rep --page raw --csv > $O/${tag}_full_raw.csv 2>/dev/null
timeout 200 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitizer_run.py 300 4 > $O/${tag}_sanitizer_racecheck.log 2>&1
tail -n 2 $O/${tag}_sanitizer_racecheck.log
python - $O/${tag}_bench.json $O/${tag}_bench_16384.json $O/${tag}_bench_256.json <<'PY'
import json, sys
for f in sys.argv[1:]:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, 'ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'cpu', (d.get('cpu_baseline') or {}).get('value'))
PY
ls -la $O | grep ${tag}_ | awk '{print $5, $9}'
