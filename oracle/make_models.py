#!/usr/bin/env python3
"""Synthesise RNNoise model blobs with the REFERENCE's own tooling (test infrastructure only).

The released weights (src/rnnoise_data.c) are a network download (reference download_model.sh:4-31)
and are absent offline, so "default model" means: default architecture (cond 128 / GRU 384,
train_rnnoise.py:48-49), default block sparsity (rnnoise.py:43-50), int8 quantised
(dump_rnnoise_weights.py --quantize), from a seeded random init.

Pipeline per model (all reference code is executed where it lies under /root/reference):
  seeded rnnoise.RNNoise  -> sparsification.common.sparsify_matrix on the six GRU matrices
  -> torch.save(checkpoint) -> torch/rnnoise/dump_rnnoise_weights.py --quantize  (30 MB C text)
  -> gcc -DDUMP_BINARY_WEIGHTS -DDISABLE_DEBUG_FLOAT src/write_weights.c  -> weights_blob.bin
  -> tests/golden/models/<name>.bin     (committed; the GPU box has no /root/reference)

Models:
  default : seed 1234, cond128/gru384, densities .2/.3/.5             ("gentle": gains ~0.4-0.6)
  hot     : same, GRU W x3, conv W x2, head W x8                        (saturating gates)
  little  : seed 4321, same dims, densities x0.5 (README:119-125 "little" = more sparsity)
  g256    : seed 99, cond128/gru256 (dimension-generality check; compared against the port only)
  tiny    : seed 7, cond96/gru128 (conv2 K = 288 is not a whole swizzle atom -> rows padded to 384);
            its state dict is kept as tiny_ckpt.npz for the exporter test
  little_b: seed 4322, cond128/gru192 (BASELINE configs[3] read literally: half-width GRUs; K = 192 padded to 256)

Usage: python oracle/make_models.py [names...]
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get("RNNOISE_REFERENCE", "/root/reference")
DST = os.path.join(REPO, "tests", "golden", "models")

SPECS = {
    "default": dict(seed=1234, cond=128, gru=384, density_scale=1.0, hot=False),
    "hot": dict(seed=1234, cond=128, gru=384, density_scale=1.0, hot=True),
    "little": dict(seed=4321, cond=128, gru=384, density_scale=0.5, hot=False),
    # other dimensions (the engine infers them from the blob): GRU 256 -> 2 swizzle atoms, 4 unit slices per CTA
    "g256": dict(seed=99, cond=128, gru=256, density_scale=1.0, hot=False),
    # small dims; its checkpoint is committed too (tiny_ckpt.npz) so the blob exporter
    # (rnnoise_b200/weights.py) can be pinned against the reference pipeline where /root/reference is absent
    "tiny": dict(seed=7, cond=96, gru=128, density_scale=1.0, hot=False),
    # the literal half-width reading of BASELINE configs[3] ("'little' half-size model"): GRU 192.  K = 192 is 1.5 swizzle
    # atoms (rows padded to 256 with zero weights), 48 units per CTA = 3 slices
    "little_b": dict(seed=4322, cond=128, gru=192, density_scale=1.0, hot=False),
}
KEEP_CKPT = {"tiny"}


def make_ckpt(path, seed, cond, gru, density_scale, hot):
    import torch
    sys.path.insert(0, os.path.join(REF, "torch", "rnnoise"))
    sys.path.insert(0, os.path.join(REF, "torch"))
    import rnnoise  # reference model definition
    from sparsification.common import sparsify_matrix
    torch.manual_seed(seed)
    m = rnnoise.RNNoise(cond_size=cond, gru_size=gru)
    with torch.no_grad():
        for g in (m.gru1, m.gru2, m.gru3):
            H = g.hidden_size
            for i, k in enumerate(["W_ir", "W_iz", "W_in"]):  # torch row order r,z,n
                d, bs, kd = rnnoise.sparse_params1[k]
                g.weight_ih_l0[i * H:(i + 1) * H] = sparsify_matrix(g.weight_ih_l0[i * H:(i + 1) * H], d * density_scale, bs, kd)
            for i, k in enumerate(["W_hr", "W_hz", "W_hn"]):
                d, bs, kd = rnnoise.sparse_params1[k]
                g.weight_hh_l0[i * H:(i + 1) * H] = sparsify_matrix(g.weight_hh_l0[i * H:(i + 1) * H], d * density_scale, bs, kd)
        if hot:
            for g in (m.gru1, m.gru2, m.gru3):
                g.weight_ih_l0 *= 3
                g.weight_hh_l0 *= 3
            m.conv1.weight *= 2
            m.conv2.weight *= 2
            m.dense_out.weight *= 8
            m.vad_dense.weight *= 8
    torch.save({"model_args": (), "model_kwargs": {"cond_size": cond, "gru_size": gru},
                "state_dict": m.state_dict()}, path)


def make(name):
    spec = SPECS[name]
    os.makedirs(DST, exist_ok=True)
    with tempfile.TemporaryDirectory() as td:
        ck = os.path.join(td, "ckpt.pth")
        make_ckpt(ck, **spec)
        gen = os.path.join(td, "gen")
        subprocess.run([sys.executable, os.path.join(REF, "torch", "rnnoise", "dump_rnnoise_weights.py"),
                        "--quantize", ck, gen], check=True, stdout=subprocess.DEVNULL)
        exe = os.path.join(td, "dump_weights_blob")
        subprocess.run(["gcc", "-O0", "-DDUMP_BINARY_WEIGHTS", "-DDISABLE_DEBUG_FLOAT",
                        "-I", os.path.join(REF, "include"), "-I", os.path.join(REF, "src"), "-I", gen,
                        os.path.join(REF, "src", "write_weights.c"), "-o", exe], check=True)
        subprocess.run([exe], cwd=td, check=True)
        out = os.path.join(DST, name + ".bin")
        shutil.copyfile(os.path.join(td, "weights_blob.bin"), out)
        if name in KEEP_CKPT:
            import numpy as np
            import torch
            sd = torch.load(ck, map_location="cpu")["state_dict"]
            np.savez(os.path.join(DST, name + "_ckpt.npz"), **{k: v.numpy() for k, v in sd.items()})
        print(name, os.path.getsize(out), "bytes ->", out)


if __name__ == "__main__":
    names = sys.argv[1:] or list(SPECS)
    for n in names:
        make(n)
