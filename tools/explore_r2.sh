#!/bin/bash
# exploration run of round 2 (on the GPU box): kernel variants side by side
O=gpurun_out; T=${1:-r2b}
python -m pytest tests -m gpu -q --tb=short > $O/${T}_tests_full.log 2>&1; tail -15 $O/${T}_tests_full.log > $O/${T}_tests.log
V1L="RNNOISE_B200_PITCH_KERNEL=v1;RNNOISE_B200_NET_KERNEL=layers"; V1F="RNNOISE_B200_PITCH_KERNEL=v1"; V2L="RNNOISE_B200_NET_KERNEL=layers"
python tools/ab_env.py --streams 4096 "$V1L" "$V1F" "$V2L" "" "RNNOISE_B200_PITCH_KERNEL=v1;RNNOISE_B200_NET_CLUSTER=8" "RNNOISE_B200_PITCH_KERNEL=v1;RNNOISE_B200_NET_CONV1=0" > $O/${T}_ab_4096.json 2> $O/${T}_ab.err
for S in 64 1024 16384; do python tools/ab_env.py --streams $S "$V1L" "$V1F" "" "RNNOISE_B200_PITCH_KERNEL=v1;RNNOISE_B200_NET_CLUSTER=8" > $O/${T}_ab_$S.json 2>> $O/${T}_ab.err; done
for v in pg8 pg4; do RNNOISE_B200_LIB_PATH=rnnoise_b200/librnnoise_b200_$v.so python tools/ab_env.py --streams 4096 "$V2L" > $O/${T}_ab_4096_$v.json 2>> $O/${T}_ab.err; done
for v in pgt pg8t; do RNNOISE_B200_LIB_PATH=rnnoise_b200/librnnoise_b200_$v.so python tools/pitch_timing.py 2048 > $O/${T}_pitch_timing_$v.json 2>> $O/${T}_ab.err; done
python tools/pcie_probe.py > $O/${T}_pcie.json 2> $O/${T}_pcie.err
python tools/tolerance_experiment.py --streams 512 --frames 2000 > $O/${T}_tolerance.json 2> $O/${T}_tolerance.err
tail -5 $O/${T}_tests.log; cat $O/${T}_ab_4096.json; cat $O/${T}_pitch_timing_*.json
