#!/bin/bash
# exploration run of round 2 (on the GPU box): kernel / scheduling variants side by side
O=gpurun_out; T=${1:-r2c}
python -m pytest tests -m gpu -q --tb=short > $O/${T}_tests_full.log 2>&1; tail -15 $O/${T}_tests_full.log > $O/${T}_tests.log
A="RNNOISE_B200_PITCH_KERNEL=v1;RNNOISE_B200_NET_KERNEL=layers;RNNOISE_B200_TAIL_OVERLAP=0"   # the round-1 pipeline
B="RNNOISE_B200_PITCH_KERNEL=v1;RNNOISE_B200_NET_KERNEL=layers"                              # + tail stage
C="RNNOISE_B200_PITCH_KERNEL=v1"                                                              # + fused network kernel
E="RNNOISE_B200_NET_KERNEL=layers"                                                            # group pitch kernel, per-layer network
for S in 4096 64 1024 16384; do python tools/ab_env.py --streams $S "$A" "$B" "$C" "$E" "" > $O/${T}_ab_$S.json 2>> $O/${T}_ab.err; done
python tools/ab_env.py --streams 4096 "RNNOISE_B200_NET_CLUSTER=8" "RNNOISE_B200_TAIL_OVERLAP=0" "RNNOISE_B200_LANES=1" "RNNOISE_B200_LANES=3" > $O/${T}_ab_4096_more.json 2>> $O/${T}_ab.err
python tools/ab_env.py --streams 2048 "$A" "" "RNNOISE_B200_NET_CLUSTER=4" "RNNOISE_B200_LANES=1" > $O/${T}_ab_2048.json 2>> $O/${T}_ab.err
RNNOISE_B200_LIB_PATH=rnnoise_b200/librnnoise_b200_pg8.so python tools/ab_env.py --streams 4096 "" > $O/${T}_ab_4096_pg8.json 2>> $O/${T}_ab.err
RNNOISE_B200_LIB_PATH=rnnoise_b200/librnnoise_b200_pgt.so python tools/pitch_timing.py 2048 > $O/${T}_pitch_timing_pgt.json 2>> $O/${T}_ab.err
tail -5 $O/${T}_tests.log; cat $O/${T}_ab_4096.json; cat $O/${T}_pitch_timing_*.json
