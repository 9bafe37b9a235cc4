#!/bin/bash
# tools/gpu_round_final.sh <tag> -- one call: validate the newest build, A/B it against the previous one (prev = the last
# committed state whose GPU tests are on record), then run the standard single-GPU evidence set (tools/gpu_round.sh)
# with whichever of the two is valid and faster.  The choice is written to gpurun_out/<tag>_choice.txt.
tag=${1:-rF}
O=gpurun_out; mkdir -p $O
L=$PWD/rnnoise_b200/librnnoise_b200
python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > $O/${tag}_tests.log
cat $O/${tag}_tests.log
REPS=2 AB_STEPS=600 bash tools/ab_libs.sh prev:rnnoise_b200/librnnoise_b200_prev.so new:rnnoise_b200/librnnoise_b200.so > $O/${tag}_ab_4096.txt 2>&1
cat $O/${tag}_ab_4096.txt
choice=$(python - "$O/${tag}_tests.log" "$O/${tag}_ab_4096.txt" <<'PY'
import re, sys
ok = " passed" in open(sys.argv[1]).read() and " failed" not in open(sys.argv[1]).read() and "error" not in open(sys.argv[1]).read().lower()
ms = {"prev": [], "new": []}
for line in open(sys.argv[2]):
    m = re.match(r"(prev|new) ms/step ([0-9.]+)", line)
    if m: ms[m.group(1)].append(float(m.group(2)))
mean = lambda v: sum(v) / len(v) if v else 1e9
print("new" if ok and mean(ms["new"]) <= mean(ms["prev"]) * 1.002 else "prev")
PY
)
echo "choice: $choice" | tee $O/${tag}_choice.txt
if [ "$choice" = prev ]; then export RNNOISE_B200_LIB_PATH=${L}_prev.so; fi
python bench.py > $O/${tag}_bench.json 2> $O/${tag}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/${tag}_bench_reference.json 2>> $O/${tag}_bench.err
python bench.py --model little --streams 16384 --steps 300 --warmup 20 --no-cpu-baseline > $O/${tag}_bench_little_16384.json 2>> $O/${tag}_bench.err
python bench.py --model little_b --streams 16384 --steps 300 --warmup 20 --no-cpu-baseline > $O/${tag}_bench_little_b_16384.json 2>> $O/${tag}_bench.err
python tools/sweep_streams.py --no-cpu 64 256 1024 4096 16384 65536 262144 > $O/${tag}_sweep_streams.md 2>> $O/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${tag}_ncu_bench.log 2>&1
RNNOISE_B200_OVERLAP=0 ncu --set full --clock-control none --import-source on --launch-skip 60 --launch-count 14 -f -o $O/${tag}_full \
    python bench.py --steps 2 --warmup 8 --no-cpu-baseline > $O/${tag}_ncu_full.log 2>&1
ncu -i $O/${tag}_full.ncu-rep --page raw --csv > $O/${tag}_full_raw.csv 2>/dev/null
python tools/pcie_probe.py > $O/${tag}_pcie.json 2> $O/${tag}_pcie.err
for tool in memcheck racecheck; do
  timeout 240 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitizer_run.py 300 6 > $O/${tag}_sanitizer_$tool.log 2>&1
done
tail -2 $O/${tag}_sanitizer_memcheck.log $O/${tag}_sanitizer_racecheck.log
python - "$O/${tag}_bench.json" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'cpu', (d.get('cpu_baseline') or {}).get('value'))
PY
cat $O/${tag}_sweep_streams.md
ls -la $O | grep ${tag}_ | awk '{print $5, $9}'
