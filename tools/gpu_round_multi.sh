#!/bin/bash
# tools/gpu_round_multi.sh <tag> [ngpus] -- multi-GPU evidence run (under gpurun --gpus N): BASELINE configs[2] (65536 streams over
# 8 GPUs) in both launch modes, weak scaling 1/2/4/8 at 4096 streams per GPU, total-stream points of configs[4] at 2/4/8 GPUs, the
# 2-GPU parity test.  Kept short: the box is charged N x its wall time.
tag=${1:-rX}; N=${2:-8}
O=gpurun_out; mkdir -p $O
python -m pytest tests/test_gpu_soak.py -m gpu -q --tb=short -k "second_device" > $O/${tag}_tests_multi.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node $N --master-port 29511 bench.py --gpus $N --streams 8192 --steps 300 --warmup 20 > $O/${tag}_bench_c3_torchrun_${N}gpu.json 2> $O/${tag}_multi.err
python bench.py --gpus $N --single-process --streams 8192 --steps 300 --warmup 20 > $O/${tag}_bench_c3_single_process_${N}gpu.json 2>> $O/${tag}_multi.err
for G in 2 4 $N; do
  [ $G -le $N ] || continue
  $TR --nproc-per-node $G --master-port $((29520 + G)) bench.py --gpus $G --steps 600 --warmup 30 > $O/${tag}_bench_4096_torchrun_${G}gpu.json 2>> $O/${tag}_multi.err
done
python bench.py --gpus $N --single-process --steps 600 --warmup 30 > $O/${tag}_bench_4096_single_process_${N}gpu.json 2>> $O/${tag}_multi.err
python bench.py --steps 600 --warmup 30 --no-cpu-baseline > $O/${tag}_bench_4096_1gpu.json 2>> $O/${tag}_multi.err
echo "| total streams | GPUs | streams / GPU | device-resident frames/s | e2e frames/s | ms/step |" > $O/${tag}_sweep_multi.md
echo "|---:|---:|---:|---:|---:|---:|" >> $O/${tag}_sweep_multi.md
for GS in "$N 16384" "$N 262144" "4 262144" "2 262144" "2 1024"; do
  set -- $GS; G=$1; S=$2
  [ $G -le $N ] || continue
  python bench.py --gpus $G --single-process --streams $((S / G)) --steps 60 --warmup 10 2>> $O/${tag}_multi.err | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('| $S | $G | %d | %.3e | %.3e | %.4f |' % ($S // $G, d['value'], d['e2e']['value'], d['ms_per_step']))
" >> $O/${tag}_sweep_multi.md
done
cat $O/${tag}_sweep_multi.md; tail -3 $O/${tag}_tests_multi.log
