#!/bin/bash
# short re-measurement of the single-process multi-GPU mode (host worker thread per device) on an N-GPU box
tag=${1:-rX}; N=${2:-8}; O=gpurun_out; mkdir -p $O
python -m pytest tests/test_gpu_soak.py -m gpu -q --tb=short -k "second_device" > $O/${tag}_tests_multi.log 2>&1
python bench.py --gpus $N --single-process --streams 8192 --steps 300 --warmup 20 > $O/${tag}_bench_c3_single_process_${N}gpu.json 2> $O/${tag}_multi.err
python bench.py --gpus $N --single-process --steps 600 --warmup 30 > $O/${tag}_bench_4096_single_process_${N}gpu.json 2>> $O/${tag}_multi.err
python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N --master-port 29533 bench.py --gpus $N --steps 600 --warmup 30 > $O/${tag}_bench_4096_torchrun_${N}gpu.json 2>> $O/${tag}_multi.err
tail -2 $O/${tag}_tests_multi.log; cut -c1-200 $O/${tag}_bench_4096_single_process_${N}gpu.json
