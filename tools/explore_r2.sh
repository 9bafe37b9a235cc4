#!/bin/bash
O=gpurun_out; T=${1:-r2f}
python -m pytest tests -m gpu -q --tb=short > $O/${T}_tests_full.log 2>&1; tail -15 $O/${T}_tests_full.log > $O/${T}_tests.log
for S in 4096 64 256 1024 2048 8192 16384; do python tools/ab_env.py --streams $S "" "RNNOISE_B200_LANES=1" "RNNOISE_B200_LANES=2" "RNNOISE_B200_LANES=3" > $O/${T}_ab_$S.json 2>> $O/${T}_ab.err; done
python tools/ab_env.py --streams 4096 "RNNOISE_B200_NET_KERNEL=fused" "RNNOISE_B200_HEADS_TILE=8" "RNNOISE_B200_HEADS_TILE=32" "RNNOISE_B200_PITCH_KERNEL=v2" > $O/${T}_ab_4096_more.json 2>> $O/${T}_ab.err
python tools/ab_env.py --streams 1024 "RNNOISE_B200_NET_KERNEL=fused" "RNNOISE_B200_HEADS_TILE=8" > $O/${T}_ab_1024_more.json 2>> $O/${T}_ab.err
tail -5 $O/${T}_tests.log; cat $O/${T}_ab_4096.json
