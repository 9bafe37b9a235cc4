"""Host API layer (rnnoise_b200/csrc/rnnoise_api.c) over a mock engine (tests/mock/mock_engine.c): the split of a
batch into lanes, the pointer offsets every entry point hands to each lane, and the routing of per-stream calls --
for many batch sizes and lane counts, without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "rnnoise_b200", "csrc")
SO = os.path.join(ROOT, "tests", "mock", "libmock_api.so")
SRCS = [os.path.join(CSRC, "rnnoise_api.c"), os.path.join(CSRC, "model_blob.c"), os.path.join(ROOT, "tests", "mock", "mock_engine.c")]


@pytest.fixture(scope="module")
def M():
    deps = SRCS + [os.path.join(CSRC, "engine.h"), os.path.join(ROOT, "include", "rnnoise.h")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.run(["gcc", "-O1", "-g", "-fPIC", "-pthread", "-shared", "-Wall", "-DRNNOISE_BUILD", "-I", CSRC, *SRCS, "-o", SO, "-lpthread"], check=True)
    L = C.CDLL(SO)
    vp, ip = C.c_void_p, C.c_int
    L.rnnoise_model_from_filename.restype = vp; L.rnnoise_model_from_filename.argtypes = [C.c_char_p]
    L.rnnoise_batch_create.restype = vp; L.rnnoise_batch_create.argtypes = [vp, ip, ip]
    L.rnnoise_batch_destroy.argtypes = [vp]
    L.rnnoise_batch_get_lanes.argtypes = [vp]; L.rnnoise_batch_get_streams.argtypes = [vp]
    for nm in ("rnnoise_process_frame_batch", "rnnoise_process_frame_batch_async", "rnnoise_process_frame_batch_device",
               "rnnoise_process_frame_batch_s16", "rnnoise_process_frame_batch_s16_async", "rnnoise_process_frame_batch_device_s16"):
        getattr(L, nm).argtypes = [vp] * 4
    for nm in ("rnnoise_process_frames_batch", "rnnoise_process_frames_batch_s16", "rnnoise_process_frames_batch_device",
               "rnnoise_process_frames_batch_device_s16"):
        getattr(L, nm).argtypes = [vp, vp, vp, vp, ip]
    L.rnnoise_batch_train_features.argtypes = [vp] * 8
    L.rnnoise_batch_prefilter_device.argtypes = [vp, vp]
    L.rnnoise_batch_set_stream.argtypes = [vp, vp]
    L.rnnoise_batch_reset_stream.argtypes = [vp, ip]
    L.rnnoise_batch_debug_read.argtypes = [vp, ip, ip, C.POINTER(C.c_float), ip]
    L.rnnoise_batch_profile_read.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_char_p), ip, C.POINTER(ip)]
    L.rnnoise_batch_launches_per_frame.argtypes = [vp]
    L.mock_hint.restype = C.c_float; L.mock_hint.argtypes = [ip]
    L.rnnoise_batch_create_multi.restype = vp; L.rnnoise_batch_create_multi.argtypes = [vp, ip, C.POINTER(ip), ip]
    L.rnnoise_batch_get_devices.argtypes = [vp]
    L.rnnoise_batch_get_shard.argtypes = [vp, ip] + [C.POINTER(ip)] * 3
    L.rnnoise_process_frame_batch_device_multi.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.rnnoise_batch_prefilter_device_multi.argtypes = [vp, C.POINTER(vp)]
    L.rnnoise_batch_set_stream_multi.argtypes = [vp, C.POINTER(vp)]
    L.rnnoise_batch_debug_set_frame_counter.argtypes = [vp, C.c_longlong]
    L.rnnoise_batch_sync.argtypes = [vp]
    L.mock_fail_next.argtypes = [ip]
    L.model = L.rnnoise_model_from_filename(os.path.join(ROOT, "tests", "golden", "models", "tiny.bin").encode())
    assert L.model
    return L


def owner(L, b, s):
    buf = (C.c_float * 8)()
    assert L.rnnoise_batch_debug_read(b, 0, s, buf, 8) == 5
    return int(buf[0]), int(buf[1]), buf


CASES = [(1, None), (100, None), (1023, None), (1024, None), (3000, None), (4096, None), (12287, None), (12288, None)]


@pytest.mark.parametrize("S,env", CASES)
def test_single_device_batch_is_one_engine_over_the_whole_buffers(M, monkeypatch, S, env):
    """One device = one engine that receives the caller's buffers unsliced (the split of the DSP stages into lanes is
    the engine's business: rnnoise_batch_get_lanes reports it); every entry point hands over the right pointers."""
    L = M
    L.mock_reset_ids()
    b = L.rnnoise_batch_create(L.model, S, 0)
    assert b and L.rnnoise_batch_get_streams(b) == S
    lanes = L.rnnoise_batch_get_lanes(b)
    assert lanes == (2 if 1024 <= S < 32768 else 1)
    probe = sorted({0, 1, S // 2, S - 1} & set(range(S)))
    own = np.array([owner(L, b, s)[:2] for s in probe])
    assert np.all(own[:, 0] == 0) and own[:, 1].tolist() == probe
    own = np.stack([np.zeros(S, int), np.arange(S)], axis=1)
    first = np.array([0, S]); lanes = 1
    assert L.rnnoise_batch_launches_per_frame(b) == 10
    lane_of = own[:, 0].astype(np.float32); local = own[:, 1].astype(np.float32)

    # single-frame float / int16 entry points
    x = np.arange(S * 480, dtype=np.float32).reshape(S, 480) % 977
    for fn in (L.rnnoise_process_frame_batch, L.rnnoise_process_frame_batch_async, L.rnnoise_process_frame_batch_device):
        out = np.zeros_like(x); vad = np.zeros(S, np.float32)
        assert fn(b, out.ctypes.data, x.ctypes.data, vad.ctypes.data) == 0
        assert np.array_equal(out, 2 * x + (1000 * lane_of + local)[:, None]) and np.array_equal(vad, 100 * lane_of + local)
    x16 = (np.arange(S * 480).reshape(S, 480) % 501).astype(np.int16)
    for fn in (L.rnnoise_process_frame_batch_s16, L.rnnoise_process_frame_batch_s16_async, L.rnnoise_process_frame_batch_device_s16):
        out = np.zeros_like(x16); vad = np.zeros(S, np.float32)
        assert fn(b, out.ctypes.data, x16.ctypes.data, None if fn is L.rnnoise_process_frame_batch_s16_async else vad.ctypes.data) == 0
        assert np.array_equal(out.astype(np.int64), x16.astype(np.int64) + (1000 * own[:, 0] + own[:, 1])[:, None])
    # multi-frame: [S][T * 480] buffers, vad [S][T]
    T = 3
    xm = (np.arange(S * T * 480, dtype=np.float32) % 601).reshape(S, T * 480)
    for fn in (L.rnnoise_process_frames_batch, L.rnnoise_process_frames_batch_device):
        out = np.zeros_like(xm); vad = np.zeros((S, T), np.float32)
        assert fn(b, out.ctypes.data, xm.ctypes.data, vad.ctypes.data, T) == 0
        want = 2 * xm.reshape(S, T, 480) + (1000 * lane_of + local)[:, None, None] + 0.25 * np.arange(T, dtype=np.float32)[None, :, None]
        assert np.array_equal(out.reshape(S, T, 480), want.astype(np.float32))
        assert np.array_equal(vad, (100 * lane_of + local)[:, None] + (np.arange(T, dtype=np.float32) / np.float32(1000))[None, :])
    xm16 = (np.arange(S * T * 480) % 301).astype(np.int16).reshape(S, T * 480)
    for fn in (L.rnnoise_process_frames_batch_s16, L.rnnoise_process_frames_batch_device_s16):
        out = np.zeros_like(xm16)
        assert fn(b, out.ctypes.data, xm16.ctypes.data, None, T) == 0
        want = xm16.reshape(S, T, 480).astype(np.int64) + (1000 * own[:, 0] + own[:, 1])[:, None, None] + np.arange(T)[None, :, None]
        assert np.array_equal(out.reshape(S, T, 480).astype(np.int64), want)
    # training records with per-stream arrays (and with all of them NULL)
    clean = (np.arange(S * 480, dtype=np.float32) % 211).reshape(S, 480); noisy = clean[::-1].copy()
    vt = (np.arange(S) % 2).astype(np.float32); nf = (np.arange(S) % 3 == 0).astype(np.int32)
    lp = (100 + np.arange(S) % 300).astype(np.int32); bl = (np.arange(S) % 33).astype(np.int32)
    rec = np.zeros((S, 98), np.float32)
    assert L.rnnoise_batch_train_features(b, rec.ctypes.data, clean.ctypes.data, noisy.ctypes.data, vt.ctypes.data, nf.ctypes.data,
                                          lp.ctypes.data, bl.ctypes.data) == 0
    assert np.array_equal(rec[:, :96], clean[:, :96] + noisy[:, :96]) and np.array_equal(rec[:, 96], 1000 * lane_of + local)
    assert np.array_equal(rec[:, 97], (vt + 2 * nf + 4 * lp + 4096 * bl).astype(np.float32))
    assert L.rnnoise_batch_train_features(b, rec.ctypes.data, clean.ctypes.data, noisy.ctypes.data, None, None, None, None) == 0
    assert np.all(rec[:, 97] == 4 * 481 + 4096 * 32)
    # prefilter hint: every lane receives the start of its own slice
    assert L.rnnoise_batch_prefilter_device(b, x.ctypes.data) == 0
    for l in range(min(lanes, 16)):
        assert L.mock_hint(l) == x[first[l], 0]
    # per-stream routing: reset and debug; stream handling: one lane runs on the caller's stream, several are bracketed
    for s in {0, S - 1, S // 2}:
        assert L.rnnoise_batch_reset_stream(b, s) == 0
        lane, loc, buf = owner(L, b, s)
        assert int(buf[2]) == loc
    assert L.rnnoise_batch_reset_stream(b, S) == -1 and L.rnnoise_batch_reset_stream(b, -1) == -1
    assert L.rnnoise_batch_set_stream(b, C.c_void_p(0x1234)) == 0
    for l in range(lanes):
        _, _, buf = owner(L, b, int(first[l]))
        assert (buf[3], buf[4]) == (1.0, 0.0)   # every lane is bracketed with the caller's stream
    # profile: per-kernel times summed over the lanes
    ms = (C.c_float * 32)(); names = (C.c_char_p * 32)(); fr = C.c_int(0)
    assert L.rnnoise_batch_profile_read(b, ms, names, 32, C.byref(fr)) == 2 and fr.value == 7
    assert ms[0] == sum(1.0 + l for l in range(lanes)) and ms[1] == 10.0 * lanes and names[0] == b"k_a"
    L.rnnoise_batch_destroy(b)


@pytest.mark.parametrize("S,devs", [(8192, [0, 1]), (10001, [3, 1, 2]), (65536, list(range(8))), (5, [0, 1, 2, 3, 4, 5, 6, 7]), (300, [2, 0])])
def test_multi_device_shards_lanes_and_pointer_offsets(M, monkeypatch, S, devs):
    """rnnoise_batch_create_multi: contiguous shards per device (stream i -> device floor(i * G / S) up to the
    remainder rule), lanes inside each shard, per-device pointers of the *_multi device call, host-buffer calls
    over the whole batch."""
    L = M
    monkeypatch.delenv("RNNOISE_B200_LANES", raising=False)
    L.mock_reset_ids()
    arr = (C.c_int * len(devs))(*devs)
    b = L.rnnoise_batch_create_multi(L.model, S, arr, len(devs))
    assert b and L.rnnoise_batch_get_streams(b) == S
    G = L.rnnoise_batch_get_devices(b)
    assert G == min(len(devs), S)
    shards = []
    for k in range(G):
        d, f, n = C.c_int(), C.c_int(), C.c_int()
        assert L.rnnoise_batch_get_shard(b, k, C.byref(d), C.byref(f), C.byref(n)) == 0
        shards.append((d.value, f.value, n.value))
    assert L.rnnoise_batch_get_shard(b, G, None, None, None) == -1
    base, rem = divmod(S, G)
    assert [n for _, _, n in shards] == [base + (1 if k < rem else 0) for k in range(G)]
    assert [d for d, _, _ in shards] == devs[:G] and shards[0][1] == 0
    assert all(shards[k][1] + shards[k][2] == shards[k + 1][1] for k in range(G - 1)) and shards[-1][1] + shards[-1][2] == S
    # every stream is handled by an engine on its shard's device
    buf = (C.c_float * 8)()
    probe = sorted({0, S - 1, S // 2} | {f for _, f, _ in shards} | {f + n - 1 for _, f, n in shards})
    eng = {}
    for s in probe:
        assert L.rnnoise_batch_debug_read(b, 0, s, buf, 8) == 5
        k = max(i for i in range(G) if shards[i][1] <= s)
        assert int(buf[5]) == shards[k][0], (s, k)
        eng[s] = (int(buf[0]), int(buf[1]))
    # host-buffer call over the whole batch
    x = (np.arange(S * 480, dtype=np.float32) % 977).reshape(S, 480)
    out = np.zeros_like(x); vad = np.zeros(S, np.float32)
    assert L.rnnoise_process_frame_batch(b, out.ctypes.data, x.ctypes.data, vad.ctypes.data) == 0
    for s in probe:
        assert np.array_equal(out[s], 2 * x[s] + 1000 * eng[s][0] + eng[s][1]) and vad[s] == 100 * eng[s][0] + eng[s][1]
    # per-device pointers: each device's buffers hold only its shard
    xs = [np.ascontiguousarray(x[f:f + n]) for _, f, n in shards]
    outs = [np.zeros_like(a) for a in xs]; vads = [np.zeros(len(a), np.float32) for a in xs]
    P = lambda lst: (C.c_void_p * G)(*[a.ctypes.data for a in lst])
    assert L.rnnoise_process_frame_batch_device_multi(b, P(outs), P(xs), P(vads)) == 0
    assert np.array_equal(np.concatenate(outs), out) and np.array_equal(np.concatenate(vads), vad)
    assert L.rnnoise_process_frame_batch_device_multi(b, P(outs), P(xs), None) == 0
    assert L.rnnoise_batch_prefilter_device_multi(b, P(xs)) == 0
    # single-pointer device calls cannot address several devices
    if G > 1:
        assert L.rnnoise_process_frame_batch_device(b, out.ctypes.data, x.ctypes.data, None) == -1
        assert L.rnnoise_batch_set_stream(b, C.c_void_p(0x10)) == -1
    st = (C.c_void_p * G)(*[0x100 + k for k in range(G)])
    assert L.rnnoise_batch_set_stream_multi(b, st) == 0
    L.rnnoise_batch_destroy(b)


def test_call_order_errors_leave_the_batch_intact_and_enqueue_errors_poison_it(M, monkeypatch):
    L = M
    L.mock_reset_ids()
    S = 600
    devs = (C.c_int * 3)(0, 1, 2)
    b = L.rnnoise_batch_create_multi(L.model, S, devs, 3)       # three engines (devices) of 200 streams
    assert L.rnnoise_batch_get_devices(b) == 3
    x = np.ones((S, 480), np.float32); out = np.zeros_like(x)
    buf = (C.c_float * 8)()
    xs = [np.ascontiguousarray(x[k * 200:(k + 1) * 200]) for k in range(3)]; outs = [np.zeros_like(a) for a in xs]
    PX = (C.c_void_p * 3)(*[a.ctypes.data for a in xs]); PO = (C.c_void_p * 3)(*[a.ctypes.data for a in outs])

    def frames_of_lanes():
        res = []
        for s in (0, 256, 599):
            L.rnnoise_batch_debug_read(b, 0, s, buf, 8)
            res.append(int(buf[6]))
        return res
    # a pending prefilter hint: host-buffer and multi-frame calls are refused before ANY engine is touched
    assert L.rnnoise_batch_prefilter_device_multi(b, PX) == 0
    assert L.rnnoise_process_frame_batch(b, out.ctypes.data, x.ctypes.data, None) == -1
    assert L.rnnoise_process_frames_batch(b, out.ctypes.data, x.ctypes.data, None, 1) == -1
    assert L.rnnoise_batch_debug_set_frame_counter(b, 5) == -1
    assert frames_of_lanes() == [0, 0, 0]
    assert L.rnnoise_process_frame_batch_device_multi(b, PO, PX, None) == 0   # consumes the hint
    assert L.rnnoise_process_frame_batch(b, out.ctypes.data, x.ctypes.data, None) == 0
    assert frames_of_lanes() == [2, 2, 2]
    # third hint in a row is refused without touching any engine
    assert L.rnnoise_batch_prefilter_device_multi(b, PX) == 0 and L.rnnoise_batch_prefilter_device_multi(b, PX) == 0
    assert L.rnnoise_batch_prefilter_device_multi(b, PX) == -1
    assert L.rnnoise_process_frame_batch_device_multi(b, PO, PX, None) == 0
    # NULL arguments: refused, nothing enqueued
    assert L.rnnoise_process_frame_batch_device_multi(b, None, PX, None) == -1
    assert frames_of_lanes() == [3, 3, 3]
    # an enqueue error in the middle engine (the devices enqueue in parallel on their worker threads, so the others
    # have gone ahead): the engines are out of step -> poisoned, every later call fails
    L.mock_fail_next(1)
    assert L.rnnoise_process_frame_batch_device_multi(b, PO, PX, None) == -1
    assert frames_of_lanes() == [4, 3, 4]
    for _ in range(2):
        assert L.rnnoise_process_frame_batch_device_multi(b, PO, PX, None) == -1
        assert L.rnnoise_process_frame_batch(b, out.ctypes.data, x.ctypes.data, None) == -1
        assert L.rnnoise_process_frames_batch(b, out.ctypes.data, x.ctypes.data, None, 1) == -1
        assert L.rnnoise_batch_reset_stream(b, 0) == -1
    assert frames_of_lanes() == [4, 3, 4]
    L.rnnoise_batch_destroy(b)
    # frame-counter hook works on a fresh batch only
    b = L.rnnoise_batch_create(L.model, 10, 0)
    assert L.rnnoise_batch_debug_set_frame_counter(b, (1 << 30) - 3) == 0
    assert L.rnnoise_process_frame_batch(b, out.ctypes.data, x.ctypes.data, None) == 0
    assert L.rnnoise_batch_debug_set_frame_counter(b, 7) == -1
    L.rnnoise_batch_destroy(b)
