#!/bin/bash
# tools/gpu_round.sh <tag> [parts...] -- the standard single-GPU evidence run of a round, on the GPU box (under gpurun).
# Everything is written under gpurun_out/<tag>_*; summaries worth keeping are copied to profiles/ by hand afterwards.
# parts (default: all): tests bench configs sweep ncu pcie sanitizer
tag=${1:-rX}; shift
parts=${*:-tests bench configs sweep ncu pcie sanitizer}
O=gpurun_out
mkdir -p $O
has() { [[ " $parts " == *" $1 "* ]]; }
if has tests; then python -m pytest tests -m gpu -q 2>&1 | tail -30 > $O/${tag}_tests.log; fi
if has bench; then
  python bench.py > $O/${tag}_bench.json 2> $O/${tag}_bench.err
  python bench.py --impl reference --steps 3 --warmup 1 > $O/${tag}_bench_reference.json 2>> $O/${tag}_bench.err
fi
if has configs; then   # BASELINE configs[3]: 'little' model, 16384 streams (both readings of "half-size")
  python bench.py --model little --streams 16384 --steps 300 --warmup 20 > $O/${tag}_bench_little_16384.json 2>> $O/${tag}_bench.err
  python bench.py --model little_b --streams 16384 --steps 300 --warmup 20 > $O/${tag}_bench_little_b_16384.json 2>> $O/${tag}_bench.err
fi
if has sweep; then python tools/sweep_streams.py 64 256 1024 4096 16384 65536 262144 > $O/${tag}_sweep_streams.md 2>> $O/${tag}_bench.err; fi
if has pcie; then python tools/pcie_probe.py > $O/${tag}_pcie.json 2> $O/${tag}_pcie.err; fi
if has ncu; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${tag}_launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${tag}_ncu_bench.log 2>&1
  # one frame's kernels of both lanes, full set (skip the first frames: caches, clocks)
  RNNOISE_B200_OVERLAP=0 ncu --set full --clock-control none --import-source on --launch-skip 60 --launch-count 14 -f -o $O/${tag}_full \
      python bench.py --steps 2 --warmup 8 --no-cpu-baseline > $O/${tag}_ncu_full.log 2>&1
  ncu -i $O/${tag}_full.ncu-rep --page raw --csv > $O/${tag}_full_raw.csv 2>/dev/null
fi
if has sanitizer; then
  for tool in memcheck racecheck; do
    compute-sanitizer --tool $tool --print-limit 20 python tools/sanitizer_run.py 300 6 > $O/${tag}_sanitizer_$tool.log 2>&1
  done
fi
ls -la $O | grep ${tag}_ | awk '{print $5, $9}'
