"""ctypes binding to the UNMODIFIED reference library in oracle/_ref/ (TEST INFRASTRUCTURE ONLY).

Mirrors the reference's private structs so tests can read intermediate state and call the
non-static internals without restating any reference logic:
  struct DenoiseState  src/denoise.c:68-88      RNNState  src/rnn.h:40-46
  LinearLayer          src/nnet.h:66-76         RNNoise   (generated header; stub in build_ref.py)
Entry points used (all exported by the plain -O2 build, no -fvisibility=hidden):
  rnnoise_*                       include/rnnoise.h:57-125
  rnn_biquad                      src/denoise.c:409
  rnn_compute_frame_features      src/denoise.c:347
  rnn_pitch_filter                src/denoise.c:421
  compute_rnn                     src/rnn.c:44
  rnn_pitch_downsample/search, rnn_remove_doubling   src/pitch.c:146,281,423
  rnn_fft_c                       src/kiss_fft.c:566 ; rnn_kfft, rnn_half_window, rnn_dct_table (rnnoise_tables.c)
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
FRAME, WINDOW, FREQ, NB_BANDS, NB_FEATURES = 480, 960, 481, 32, 65
PITCH_BUF = 1728


class LinearLayer(C.Structure):
    _fields_ = [("bias", C.c_void_p), ("subias", C.c_void_p), ("weights", C.c_void_p),
                ("float_weights", C.c_void_p), ("weights_idx", C.c_void_p), ("diag", C.c_void_p),
                ("scale", C.c_void_p), ("nb_inputs", C.c_int), ("nb_outputs", C.c_int)]


LAYERS = ["conv1", "conv2", "gru1_input", "gru1_recurrent", "gru2_input", "gru2_recurrent",
          "gru3_input", "gru3_recurrent", "dense_out", "vad_dense"]


class RNNoiseModelStruct(C.Structure):
    _fields_ = [(n, LinearLayer) for n in LAYERS]


def make_state_types(cond, gru):
    class RNNState(C.Structure):
        _fields_ = [("conv1_state", C.c_float * (2 * NB_FEATURES)), ("conv2_state", C.c_float * (2 * cond)),
                    ("gru1_state", C.c_float * gru), ("gru2_state", C.c_float * gru), ("gru3_state", C.c_float * gru)]

    class DenoiseState(C.Structure):
        _fields_ = [("model", RNNoiseModelStruct), ("arch", C.c_int),
                    ("analysis_mem", C.c_float * FRAME), ("memid", C.c_int),
                    ("synthesis_mem", C.c_float * FRAME), ("pitch_buf", C.c_float * PITCH_BUF),
                    ("pitch_enh_buf", C.c_float * PITCH_BUF), ("last_gain", C.c_float),
                    ("last_period", C.c_int), ("mem_hp_x", C.c_float * 2), ("lastg", C.c_float * NB_BANDS),
                    ("rnn", RNNState), ("delayed_X", C.c_float * (2 * FREQ)), ("delayed_P", C.c_float * (2 * FREQ)),
                    ("delayed_Ex", C.c_float * NB_BANDS), ("delayed_Ep", C.c_float * NB_BANDS),
                    ("delayed_Exp", C.c_float * NB_BANDS)]
    return RNNState, DenoiseState


def lib_path(kind="rtcd", cond=128, gru=384):
    name = {"rtcd": "librnnoise_ref_c%d_g%d.so", "generic": "librnnoise_ref_generic_c%d_g%d.so"}[kind]
    return os.path.join(HERE, "_ref", name % (cond, gru))


def bench_path(cond=128, gru=384, vnni=False):
    return os.path.join(HERE, "_ref", "ref_bench_%sc%d_g%d" % ("vnni_" if vnni else "", cond, gru))


def available(kind="rtcd", cond=128, gru=384):
    return os.path.exists(lib_path(kind, cond, gru))


def fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class RefLib:
    """One loaded reference build + one model blob."""

    def __init__(self, model_path, kind="rtcd", cond=128, gru=384):
        self.lib = C.CDLL(lib_path(kind, cond, gru))
        self.kind, self.cond, self.gru = kind, cond, gru
        self.RNNState, self.DenoiseState = make_state_types(cond, gru)
        L = self.lib
        L.rnnoise_get_size.restype = C.c_int
        L.rnnoise_model_from_filename.restype = C.c_void_p
        L.rnnoise_model_from_filename.argtypes = [C.c_char_p]
        L.rnnoise_create.restype = C.POINTER(self.DenoiseState)
        L.rnnoise_create.argtypes = [C.c_void_p]
        L.rnnoise_destroy.argtypes = [C.POINTER(self.DenoiseState)]
        L.rnnoise_process_frame.restype = C.c_float
        L.rnnoise_process_frame.argtypes = [C.POINTER(self.DenoiseState), C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.rnn_compute_frame_features.restype = C.c_int
        L.rnn_remove_doubling.restype = C.c_float
        assert L.rnnoise_get_size() == C.sizeof(self.DenoiseState), \
            (L.rnnoise_get_size(), C.sizeof(self.DenoiseState))
        self.model = L.rnnoise_model_from_filename(model_path.encode())
        assert self.model

    def create(self, arch=None):
        st = self.lib.rnnoise_create(self.model)
        assert st, "rnnoise_create failed"
        if arch is not None and self.kind == "rtcd":
            st.contents.arch = arch
        return st

    def destroy(self, st):
        self.lib.rnnoise_destroy(st)

    def process_frame(self, st, x):
        """x float32[480] -> (out float32[480], vad)."""
        x = np.ascontiguousarray(x, np.float32)
        out = np.empty(FRAME, np.float32)
        vad = self.lib.rnnoise_process_frame(st, fptr(out), fptr(x))
        return out, float(vad)

    def process_frame_traced(self, st, x):
        """Runs the frame twice: once on a scratch COPY of the state through the non-static stage
        functions to capture intermediates, once for real through rnnoise_process_frame().
        Returns a dict of numpy arrays."""
        L = self.lib
        x = np.ascontiguousarray(x, np.float32)
        tmp = self.DenoiseState()
        C.memmove(C.byref(tmp), st, C.sizeof(tmp))
        a_hp = np.array([-1.99599, 0.99600], np.float32)  # denoise.c:469
        b_hp = np.array([-2, 1], np.float32)              # denoise.c:470
        xb = np.empty(FRAME, np.float32)
        L.rnn_biquad(fptr(xb), tmp.mem_hp_x, fptr(x), fptr(b_hp), fptr(a_hp), FRAME)
        X = np.zeros(2 * FREQ, np.float32); P = np.zeros(2 * FREQ, np.float32)
        Ex = np.zeros(NB_BANDS, np.float32); Ep = np.zeros(NB_BANDS, np.float32); Exp = np.zeros(NB_BANDS, np.float32)
        feat = np.zeros(NB_FEATURES, np.float32)
        silence = L.rnn_compute_frame_features(C.byref(tmp), fptr(X), fptr(P), fptr(Ex), fptr(Ep), fptr(Exp), fptr(feat), fptr(xb))
        g = np.zeros(NB_BANDS, np.float32)
        vad = C.c_float(0)
        if not silence:
            L.compute_rnn(C.byref(tmp.model), C.byref(tmp.rnn), fptr(g), C.byref(vad), fptr(feat), tmp.arch)
        out, vad2 = self.process_frame(st, x)
        s = st.contents
        return dict(xb=xb, X=X, P=P, Ex=Ex, Ep=Ep, Exp=Exp, features=feat, silence=int(silence),
                    pitch=int(tmp.last_period), pitch_gain=float(tmp.last_gain), g_raw=g, vad=float(vad2),
                    vad_traced=float(vad.value), out=out, lastg=np.array(s.lastg, np.float32),
                    conv1_state=np.array(s.rnn.conv1_state, np.float32), conv2_state=np.array(s.rnn.conv2_state, np.float32),
                    gru1=np.array(s.rnn.gru1_state, np.float32), gru2=np.array(s.rnn.gru2_state, np.float32),
                    gru3=np.array(s.rnn.gru3_state, np.float32))
