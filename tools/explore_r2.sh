#!/bin/bash
# exploration run of round 2 (on the GPU box): kernel / scheduling variants side by side
O=gpurun_out; T=${1:-r2d}
python -m pytest tests -m gpu -q --tb=short > $O/${T}_tests_full.log 2>&1; tail -15 $O/${T}_tests_full.log > $O/${T}_tests.log
A="RNNOISE_B200_PITCH_KERNEL=v1;RNNOISE_B200_NET_KERNEL=layers;RNNOISE_B200_TAIL_OVERLAP=0"   # the round-1 pipeline (with this round's kernel bodies)
for S in 4096 64 256 1024 16384; do python tools/ab_env.py --streams $S "$A" "" "RNNOISE_B200_NET_KERNEL=fused" "RNNOISE_B200_NET_KERNEL=layers" "RNNOISE_B200_PITCH_KERNEL=v2" > $O/${T}_ab_$S.json 2>> $O/${T}_ab.err; done
# source-level profile of the two latency-bound kernels (one lane, no overlap)
RNNOISE_B200_OVERLAP=0 RNNOISE_B200_LANES=1 ncu --set full --clock-control none --import-source on --kernel-name regex:"k_heads2|k_tc2" --launch-skip 40 --launch-count 5 -f -o $O/${T}_net \
    python bench.py --streams 2048 --steps 2 --warmup 8 --no-cpu-baseline > $O/${T}_ncu_net.log 2>&1
ncu -i $O/${T}_net.ncu-rep --page raw --csv > $O/${T}_net_raw.csv 2>/dev/null
ncu -i $O/${T}_net.ncu-rep --page source --csv --print-source sass > $O/${T}_net_source.csv 2>/dev/null
tail -5 $O/${T}_tests.log; cat $O/${T}_ab_4096.json
