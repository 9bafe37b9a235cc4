#!/usr/bin/env python3
"""Does splitting a batch into independent sub-batches ("lanes") on separate CUDA streams raise whole-GPU
throughput?  Same total streams, device-resident input, prefilter hints, CUDA-event timing per lane.
usage: python tools/lanes_experiment.py [total_streams] [lanes ...]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import rnnoise_b200 as rb

S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
lanes_list = [int(x) for x in sys.argv[2:]] or [1, 2, 4]
model = rb.Model(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "models", "default.bin"))
POOL, K, WARM = 16, 300, 30
for L in lanes_list:
    s = S // L
    batches = [rb.Batch(model, s) for _ in range(L)]
    streams = [torch.cuda.Stream(priority=-1) for _ in range(L)]
    pools = [(torch.randn(POOL, s, 480, device="cuda") * 1000) for _ in range(L)]
    outs = [torch.empty(s, 480, device="cuda") for _ in range(L)]
    vads = [torch.empty(s, device="cuda") for _ in range(L)]
    for b, st in zip(batches, streams):
        b.set_stream(st.cuda_stream)
    torch.cuda.synchronize()

    def run(n, first):
        for i in range(n):
            for l, b in enumerate(batches):
                if first and i == 0:
                    b.prefilter_device(pools[l][0].data_ptr())
                b.prefilter_device(pools[l][(i + 1) % POOL].data_ptr())
                b.process_device(outs[l].data_ptr(), pools[l][i % POOL].data_ptr(), vads[l].data_ptr())
    run(WARM, True)
    for b in batches: b.sync()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    # continue the hint chain: frame WARM's prefilter was already issued by the last warm-up iteration
    for i in range(K):
        for l, b in enumerate(batches):
            b.prefilter_device(pools[l][(WARM + i + 1) % POOL].data_ptr())
            b.process_device(outs[l].data_ptr(), pools[l][(WARM + i) % POOL].data_ptr(), vads[l].data_ptr())
    for b in batches: b.sync()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"lanes {L} x {s} streams: {dt / K * 1e3:.4f} ms per step of {S} streams -> {S * K / dt / 1e6:.3f} M frames/s")
    for b in batches: b.destroy()
