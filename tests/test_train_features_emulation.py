"""Training-feature extraction: the device source of k_train_features executed on
the host, thread id by thread id, must reproduce the UNMODIFIED reference built with -DTRAINING=1
(oracle/_ref/librnnoise_ref_training.so: denoise.c et al. + the dump_features frame loop) bit for bit."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import trainbind
from rnnoise_b200.synth_pcm import train_pair, train_params

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_SRC = os.path.join(ROOT, "tests", "emu", "emu_dsp.cpp")
EMU_SO = os.path.join(ROOT, "tests", "emu", "libemu_dsp.so")

pytestmark = pytest.mark.skipif(not trainbind.available(), reason="oracle/_ref training build missing (python oracle/build_ref.py)")


@pytest.fixture(scope="module")
def emu():
    deps = [EMU_SRC] + [os.path.join(ROOT, "rnnoise_b200", "csrc", f) for f in ("dsp_core.cuh", "dsp_stream.cuh", "dsp_tables.hpp")]
    if not os.path.exists(EMU_SO) or any(os.path.getmtime(d) > os.path.getmtime(EMU_SO) for d in deps):
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-DPITCH_NS=4",
                        "-I", os.path.join(ROOT, "rnnoise_b200", "csrc"), EMU_SRC, "-o", EMU_SO], check=True)
    E = C.CDLL(EMU_SO)
    E.emu_create.restype = C.c_void_p
    E.emu_destroy.argtypes = [C.c_void_p]
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
    E.emu_train.argtypes = [C.c_void_p, fp, fp, C.c_int, ip, ip, fp, ip, fp, ip]
    return E


@pytest.mark.parametrize("streams,frames", [((0, 1, 2, 4), 40), ((15, 9, 3), 60)])
def test_train_records_bit_identical_to_training_reference(emu, streams, frames):
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
    n = len(streams)
    pairs = [train_pair(s, frames) for s in streams]
    par = [train_params(s) for s in streams]
    lowpass = np.array([p[0] for p in par], np.int32); band_lp = np.array([p[1] for p in par], np.int32)
    noise_free = np.array([p[2] for p in par], np.int32)
    refs = [trainbind.RefTrain() for _ in streams]
    e = emu.emu_create()
    undefined = 0
    for f in range(frames):
        clean = np.ascontiguousarray(np.stack([p[0][f] for p in pairs])); noisy = np.ascontiguousarray(np.stack([p[1][f] for p in pairs]))
        vad = np.array([float((f // 7 + s) % 2) for s in streams], np.float32)
        rec = np.zeros((n, 98), np.float32); quiet = np.zeros(n, np.int32)
        emu.emu_train(e, clean.ctypes.data_as(fp), noisy.ctypes.data_as(fp), n, lowpass.ctypes.data_as(ip), band_lp.ctypes.data_as(ip),
                      vad.ctypes.data_as(fp), noise_free.ctypes.data_as(ip), rec.ctypes.data_as(fp), quiet.ctypes.data_as(ip))
        for q, s in enumerate(streams):
            want, wq, _ = refs[q].frame(clean[q], noisy[q], vad[q], noise_free[q], lowpass[q], band_lp[q])
            assert rec[q].tobytes() == want.tobytes(), (f, s, np.nonzero(rec[q] != want)[0][:8])
            assert quiet[q] == wq, (f, s)
            undefined += int(np.sum(want[65:97] == -1))
    assert undefined > 0   # the masks were exercised
    for r in refs:
        r.close()
    emu.emu_destroy(e)
