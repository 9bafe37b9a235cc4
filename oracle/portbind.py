"""ctypes binding to oracle/librnnoise_port.so (the CPU restatement) -- TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "librnnoise_port.so")
FRAME, FREQ, NB_BANDS, NB_FEATURES = 480, 481, 32, 65
MAX_GRU, MAX_COND = 1024, 512


class Trace(C.Structure):
    _fields_ = [("xb", C.c_float * FRAME), ("X", C.c_float * (2 * FREQ)), ("P", C.c_float * (2 * FREQ)),
                ("Ex", C.c_float * NB_BANDS), ("Ep", C.c_float * NB_BANDS), ("Exp", C.c_float * NB_BANDS),
                ("features", C.c_float * NB_FEATURES), ("silence", C.c_int), ("pitch", C.c_int),
                ("pitch_gain", C.c_float), ("g_raw", C.c_float * NB_BANDS), ("vad", C.c_float)]


class State(C.Structure):
    _fields_ = [("analysis_mem", C.c_float * FRAME), ("synthesis_mem", C.c_float * FRAME),
                ("pitch_buf", C.c_float * 1728), ("last_gain", C.c_float), ("last_period", C.c_int),
                ("mem_hp_x", C.c_float * 2), ("lastg", C.c_float * NB_BANDS),
                ("conv1_state", C.c_float * (2 * NB_FEATURES)), ("conv2_state", C.c_float * (2 * MAX_COND)),
                ("gru_state", (C.c_float * MAX_GRU) * 3),
                ("delayed_X", C.c_float * (2 * FREQ)), ("delayed_P", C.c_float * (2 * FREQ)),
                ("delayed_Ex", C.c_float * NB_BANDS), ("delayed_Ep", C.c_float * NB_BANDS),
                ("delayed_Exp", C.c_float * NB_BANDS)]


def build(force=False):
    src = [os.path.join(HERE, "rnnoise_port.c"), os.path.join(HERE, "rnnoise_port.h")]
    if force or not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in src):
        subprocess.run(["make", "-C", HERE, "-B", "librnnoise_port.so"], check=True, capture_output=True)
    return SO


def fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class Port:
    def __init__(self, model_path=None):
        self.lib = L = C.CDLL(build())
        L.rp_model_from_file.restype = C.c_void_p
        L.rp_model_from_file.argtypes = [C.c_char_p]
        L.rp_model_from_buffer.restype = C.c_void_p
        L.rp_model_from_buffer.argtypes = [C.c_void_p, C.c_int]
        L.rp_model_free.argtypes = [C.c_void_p]
        L.rp_state_create.restype = C.POINTER(State)
        L.rp_state_destroy.argtypes = [C.POINTER(State)]
        L.rp_process_frame.restype = C.c_float
        L.rp_process_frame.argtypes = [C.c_void_p, C.POINTER(State), C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(Trace)]
        L.rp_state_size.restype = C.c_int
        L.rp_half_window.restype = C.POINTER(C.c_float)
        L.rp_dct_table.restype = C.POINTER(C.c_float)
        L.rp_twiddles.restype = C.POINTER(C.c_float)
        L.rp_bitrev.restype = C.POINTER(C.c_int)
        L.rp_tanh.restype = C.c_float; L.rp_tanh.argtypes = [C.c_float]
        L.rp_sigmoid.restype = C.c_float; L.rp_sigmoid.argtypes = [C.c_float]
        L.rp_quant_u8.restype = C.c_ubyte; L.rp_quant_u8.argtypes = [C.c_float]
        L.rp_pitch_search.restype = C.c_int
        L.rp_remove_doubling.restype = C.c_float
        L.rp_remove_doubling.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_int, C.c_float]
        assert L.rp_state_size() == C.sizeof(State)
        L.rp_train_frame.restype = C.c_int
        L.rp_train_frame.argtypes = [C.POINTER(State), C.POINTER(State), C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float,
                                     C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]
        self.model = None
        if model_path:
            self.model = L.rp_model_from_file(model_path.encode())
            assert self.model, "port failed to parse " + model_path

    def create(self):
        return self.lib.rp_state_create()

    def destroy(self, st):
        self.lib.rp_state_destroy(st)

    def train_frame(self, clean_st, noisy_st, clean, noisy, vad_target=0.0, noise_free=0, lowpass=481, band_lp=32):
        """Training-feature record of one frame (dump_features.c:466-491 semantics) -> (rec[98], quiet flag)."""
        c = np.ascontiguousarray(clean, np.float32); n = np.ascontiguousarray(noisy, np.float32)
        rec = np.zeros(98, np.float32)
        q = self.lib.rp_train_frame(clean_st, noisy_st, fptr(c), fptr(n), float(vad_target), int(noise_free), int(lowpass),
                                    int(band_lp), fptr(rec))
        return rec, q

    def process_frame(self, st, x, trace=True):
        x = np.ascontiguousarray(x, np.float32)
        out = np.empty(FRAME, np.float32)
        tr = Trace()
        vad = self.lib.rp_process_frame(self.model, st, fptr(out), fptr(x), C.byref(tr) if trace else None)
        d = dict(out=out, vad=float(vad))
        if trace:
            s = st.contents
            d.update(xb=np.array(tr.xb, np.float32), X=np.array(tr.X, np.float32), P=np.array(tr.P, np.float32),
                     Ex=np.array(tr.Ex, np.float32), Ep=np.array(tr.Ep, np.float32), Exp=np.array(tr.Exp, np.float32),
                     features=np.array(tr.features, np.float32), silence=int(tr.silence), pitch=int(tr.pitch),
                     pitch_gain=float(tr.pitch_gain), g_raw=np.array(tr.g_raw, np.float32),
                     lastg=np.array(s.lastg, np.float32))
        return d

    def tables(self):
        L = self.lib
        return dict(half_window=np.ctypeslib.as_array(L.rp_half_window(), (480,)).copy(),
                    dct=np.ctypeslib.as_array(L.rp_dct_table(), (1024,)).copy(),
                    twiddles=np.ctypeslib.as_array(L.rp_twiddles(), (1920,)).copy(),
                    bitrev=np.ctypeslib.as_array(L.rp_bitrev(), (960,)).copy())
