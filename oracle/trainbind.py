"""ctypes binding of oracle/_ref/librnnoise_ref_training.so (the unmodified reference, -DTRAINING=1, plus
the frame-loop driver oracle/ref_train.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_ref", "librnnoise_ref_training.so")


def available():
    return os.path.exists(SO)


class RefTrain:
    def __init__(self):
        self.lib = C.CDLL(SO)
        self.lib.ref_train_create.restype = C.c_void_p
        self.lib.ref_train_destroy.argtypes = [C.c_void_p]
        fp = C.POINTER(C.c_float)
        self.lib.ref_train_frame.restype = C.c_int
        self.lib.ref_train_frame.argtypes = [C.c_void_p, fp, fp, C.c_float, C.c_int, C.c_int, C.c_int, fp, fp]
        self.h = self.lib.ref_train_create()

    def frame(self, clean, noisy, vad_target=0.0, noise_free=0, lowpass=481, band_lp=32):
        """-> (rec[98], quiet flag, dbg dict)"""
        fp = C.POINTER(C.c_float)
        c = np.ascontiguousarray(clean, np.float32); n = np.ascontiguousarray(noisy, np.float32)
        rec = np.zeros(98, np.float32); dbg = np.zeros(2052, np.float32)
        q = self.lib.ref_train_frame(self.h, c.ctypes.data_as(fp), n.ctypes.data_as(fp), float(vad_target), int(noise_free),
                                     int(lowpass), int(band_lp), rec.ctypes.data_as(fp), dbg.ctypes.data_as(fp))
        return rec, q, dict(X=dbg[:962], P=dbg[962:1924], Ex=dbg[1924:1956], Ep=dbg[1956:1988], Exp=dbg[1988:2020], Ey=dbg[2020:2052])

    def close(self):
        if self.h:
            self.lib.ref_train_destroy(self.h)
            self.h = None
