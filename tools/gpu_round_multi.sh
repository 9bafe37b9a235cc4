#!/bin/bash
# tools/gpu_round_multi.sh <tag> [ngpus] -- multi-GPU evidence run (under gpurun --gpus N): BASELINE configs[2] (65536 streams over
# 8 GPUs) in both launch modes, weak scaling 1/2/4/8, the total-stream sweep of configs[4] at 2/4/8 GPUs, the 2-GPU parity test.
tag=${1:-rX}; N=${2:-8}
O=gpurun_out; mkdir -p $O
python -m pytest tests/test_gpu_soak.py -m gpu -q --tb=short -k "second_device" > $O/${tag}_tests_multi.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
# configs[2]: 8192 streams per GPU
$TR --nproc-per-node $N --master-port 29511 bench.py --gpus $N --streams 8192 --steps 500 --warmup 30 > $O/${tag}_bench_c3_torchrun_${N}gpu.json 2> $O/${tag}_multi.err
python bench.py --gpus $N --single-process --streams 8192 --steps 500 --warmup 30 > $O/${tag}_bench_c3_single_process_${N}gpu.json 2>> $O/${tag}_multi.err
# weak scaling at the headline size (4096 per GPU), both modes
for G in 2 4 $N; do
  [ $G -le $N ] || continue
  $TR --nproc-per-node $G --master-port $((29520 + G)) bench.py --gpus $G --steps 1000 --warmup 30 > $O/${tag}_bench_4096_torchrun_${G}gpu.json 2>> $O/${tag}_multi.err
  python bench.py --gpus $G --single-process --steps 1000 --warmup 30 > $O/${tag}_bench_4096_single_process_${G}gpu.json 2>> $O/${tag}_multi.err
done
python bench.py --steps 1000 --warmup 30 --no-cpu-baseline > $O/${tag}_bench_4096_1gpu.json 2>> $O/${tag}_multi.err
# configs[4]: TOTAL streams sharded over G GPUs (single process: one host thread, no NCCL)
echo "| total streams | GPUs | streams / GPU | device-resident frames/s | e2e frames/s | ms/step |" > $O/${tag}_sweep_multi.md
echo "|---:|---:|---:|---:|---:|---:|" >> $O/${tag}_sweep_multi.md
for G in 2 4 $N; do
  [ $G -le $N ] || continue
  for S in 1024 16384 65536 262144; do
    python bench.py --gpus $G --single-process --streams $((S / G)) --steps 100 --warmup 10 2>> $O/${tag}_multi.err | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('| $S | $G | %d | %.3e | %.3e | %.4f |' % ($S // $G, d['value'], d['e2e']['value'], d['ms_per_step']))
" >> $O/${tag}_sweep_multi.md
  done
done
cat $O/${tag}_sweep_multi.md; tail -3 $O/${tag}_tests_multi.log
