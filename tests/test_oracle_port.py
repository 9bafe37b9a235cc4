"""The oracle's C restatement (oracle/rnnoise_port.c) against the reference's golden vectors
(tests/golden/ref_*.npz, produced by the unmodified reference build) and, when the prebuilt
reference library is present (oracle/_ref), against the live reference, stage by stage."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import refbind
from oracle.portbind import Port, fptr
from rnnoise_b200.synth_pcm import stream_pcm

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODELS = ("default", "hot", "little")


def run_port(port, pcm):
    st = port.create()
    rows = [port.process_frame(st, pcm[f]) for f in range(pcm.shape[0])]
    port.destroy(st)
    return rows


@pytest.mark.parametrize("name", MODELS)
def test_port_matches_reference_golden(name, models_dir):
    g = np.load(os.path.join(GOLD, f"ref_{name}.npz"))
    port = Port(os.path.join(models_dir, name + ".bin"))
    frames = int(g["frames"])
    for s in g["streams"]:
        rows = run_port(port, stream_pcm(int(s), frames))
        feat = np.stack([r["features"] for r in rows])
        # DSP path: bit-identical to the reference build (integer-like rigour)
        assert np.array_equal(feat, g[f"s{s}_features"]), f"features differ (stream {s})"
        assert np.array_equal(np.stack([r["Ex"] for r in rows]), g[f"s{s}_Ex"])
        assert [r["pitch"] for r in rows] == list(g[f"s{s}_pitch"])
        assert [r["silence"] for r in rows] == list(g[f"s{s}_silence"])
        # network path: the port uses an exact reciprocal where AVX2 uses rcp_ps; it must sit inside
        # the reference's own cross-build envelope E_ref = |AVX2 - generic C| (SURVEY App. D rule: <= 2 E_ref)
        for key, floor in (("g_raw", 2e-4), ("vad", 2e-4), ("lastg", 2e-4), ("out", 0.25)):
            ref = g[f"s{s}_{key}"]
            mine = np.stack([np.atleast_1d(r[key]) for r in rows]).reshape(ref.shape)
            e_ref = np.abs(g[f"s{s}_generic_{key}"] - ref).max()
            err = np.abs(mine - ref).max()
            assert err <= 2 * e_ref + floor, (name, int(s), key, float(err), float(e_ref))


def test_tables_equal_reference_tables():
    if not refbind.available():
        pytest.skip("oracle/_ref not built")
    lib = C.CDLL(refbind.lib_path())
    t = Port().tables()
    hw = np.ctypeslib.as_array((C.c_float * 480).in_dll(lib, "rnn_half_window"))
    dct = np.ctypeslib.as_array((C.c_float * 1024).in_dll(lib, "rnn_dct_table"))
    assert np.array_equal(hw, t["half_window"]) and np.array_equal(dct, t["dct"])

    class KF(C.Structure):
        _fields_ = [("nfft", C.c_int), ("scale", C.c_float), ("shift", C.c_int), ("factors", C.c_int16 * 16),
                    ("bitrev", C.POINTER(C.c_int32)), ("twiddles", C.POINTER(C.c_float)), ("arch", C.c_void_p)]
    kf = KF.in_dll(lib, "rnn_kfft")
    assert np.float32(kf.scale) == np.float32(0.0010416667)
    assert np.array_equal(np.ctypeslib.as_array(kf.bitrev, (960,)), t["bitrev"])
    assert np.array_equal(np.ctypeslib.as_array(kf.twiddles, (1920,)), t["twiddles"])


LIVE = [("default", 128, 384), ("little_b", 128, 192)]   # reference builds of oracle/build_ref.py (model dims are compile-time there)


@pytest.mark.parametrize("name,cond,gru", LIVE)
def test_port_linear_layers_bit_exact_vs_reference_avx2(models_dir, name, cond, gru):
    """Every compute_linear of the model: port == reference AVX2 kernel, bit for bit."""
    if not refbind.available("rtcd", cond, gru):
        pytest.skip("oracle/_ref not built")
    mp = os.path.join(models_dir, name + ".bin")
    ref, port = refbind.RefLib(mp, "rtcd", cond, gru), Port(mp)
    if not hasattr(ref.lib, "rnn_compute_linear_avx2"):
        pytest.skip("no AVX2 object in this reference build")

    class RL(C.Structure):
        _fields_ = [("nb_in", C.c_int), ("nb_out", C.c_int), ("is_int8", C.c_int), ("w8", C.c_void_p), ("wf", C.c_void_p),
                    ("bias", C.c_void_p), ("subias", C.c_void_p), ("scale", C.c_void_p), ("diag", C.c_void_p)]

    class RM(C.Structure):
        _fields_ = [("cond", C.c_int), ("gru", C.c_int), ("conv1", RL), ("conv2", RL), ("gru_in", RL * 3),
                    ("gru_rec", RL * 3), ("dense_out", RL), ("vad_dense", RL), ("blob", C.c_void_p)]
    pm = RM.from_address(port.model)
    pl = dict(conv1=pm.conv1, conv2=pm.conv2, dense_out=pm.dense_out, vad_dense=pm.vad_dense)
    for k in range(3):
        pl[f"gru{k + 1}_input"], pl[f"gru{k + 1}_recurrent"] = pm.gru_in[k], pm.gru_rec[k]
    st = ref.create()
    rng = np.random.default_rng(0)
    for name in refbind.LAYERS:
        lay = getattr(st.contents.model, name)
        for t in range(4):
            x = rng.uniform(-1, 1, lay.nb_inputs).astype(np.float32)
            if t == 0:
                x[:6] = [1, -1, 0.5, 1.0039, 0.00394, -0.00394]  # quantiser edge cases
            a = np.zeros(lay.nb_outputs, np.float32); b = np.zeros_like(a)
            ref.lib.rnn_compute_linear_avx2(C.byref(lay), fptr(a), fptr(x))
            port.lib.rp_linear(C.byref(pl[name]), fptr(b), fptr(x), None)
            assert np.array_equal(a, b), name
    ref.destroy(st)


@pytest.mark.parametrize("name,cond,gru", LIVE)
def test_port_vs_live_reference_dsp_bit_exact(models_dir, name, cond, gru):
    if not refbind.available("rtcd", cond, gru):
        pytest.skip("oracle/_ref not built")
    mp = os.path.join(models_dir, name + ".bin")
    ref, port = refbind.RefLib(mp, "rtcd", cond, gru), Port(mp)
    gen = refbind.RefLib(mp, "generic", cond, gru)
    pcm = stream_pcm(3, 60)
    st, sp, sg = ref.create(), port.create(), gen.create()
    err = e_ref = 0.0
    for f in range(pcm.shape[0]):
        a, b = ref.process_frame_traced(st, pcm[f]), port.process_frame(sp, pcm[f])
        g_out, _ = gen.process_frame(sg, pcm[f])
        for k in ("xb", "X", "P", "Ex", "Ep", "Exp", "features"):
            assert np.array_equal(a[k], b[k]), (k, f)
        assert a["pitch"] == b["pitch"] and a["silence"] == b["silence"]
        err = max(err, float(np.abs(a["out"] - b["out"]).max())); e_ref = max(e_ref, float(np.abs(a["out"] - g_out).max()))
    assert err <= 2 * e_ref + 0.25, (err, e_ref)   # SURVEY App. D rule: inside the reference's own cross-build envelope


def test_port_training_frame_matches_reference_training_build():
    """rp_train_frame (TRAINING semantics of the feature path + ideal gains) against the unmodified reference
    built with -DTRAINING=1 (oracle/_ref/librnnoise_ref_training.so) and against the committed goldens."""
    from oracle import trainbind
    from rnnoise_b200.synth_pcm import train_pair, train_params
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_train.npz"))
    streams, frames = [int(s) for s in g["streams"]], int(g["frames"])
    port = Port(None)
    for q, s in enumerate(streams):
        clean, noisy = train_pair(s, frames)
        lp, blp, nf = train_params(s)
        cs, ns = port.create(), port.create()
        live = trainbind.RefTrain() if trainbind.available() else None
        for f in range(frames):
            vt = float((f // 7 + s) % 2)
            rec, quiet = port.train_frame(cs, ns, clean[f], noisy[f], vt, nf, lp, blp)
            assert rec.tobytes() == g["rec"][f, q].tobytes() and quiet == int(g["quiet"][f, q]), (s, f)
            if live:
                want, wq, _ = live.frame(clean[f], noisy[f], vt, nf, lp, blp)
                assert rec.tobytes() == want.tobytes() and quiet == wq, (s, f)
        port.destroy(cs); port.destroy(ns)
        if live:
            live.close()


def test_port_rejects_malformed_blobs(models_dir):
    port = Port()
    blob = open(os.path.join(models_dir, "default.bin"), "rb").read()
    L = port.lib
    assert L.rp_model_from_buffer(blob, len(blob))
    assert not L.rp_model_from_buffer(blob, len(blob) - 100)          # truncated record
    assert not L.rp_model_from_buffer(blob[64:], len(blob) - 64)        # header lost
    bad = bytearray(blob); bad[12:16] = (10 ** 9).to_bytes(4, "little")  # size > block_size
    assert not L.rp_model_from_buffer(bytes(bad), len(bad))
