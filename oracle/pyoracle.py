"""ctypes wrapper around the CPU oracle (TEST INFRASTRUCTURE ONLY -- see orc_fold.h).

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
legs.  Never imported by the product package.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liborc_fold.so")


def build(force=False):
    src = os.path.join(_HERE, "orc_fold.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        L = C.CDLL(_LIB)
        L.orc_params_load.restype = C.c_void_p
        L.orc_params_load.argtypes = [C.c_char_p]
        L.orc_params_free.argtypes = [C.c_void_p]
        L.orc_params_get.restype = C.c_int
        L.orc_params_get.argtypes = [C.c_void_p, C.c_char_p] + [C.c_int] * 6
        L.orc_eval.restype = C.c_int
        L.orc_eval.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_char_p]
        L.orc_mfe.restype = C.c_int
        L.orc_mfe.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.POINTER(C.c_longlong)]
        L.orc_pf.restype = C.c_double
        L.orc_pf.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orc_ensemble_defect.restype = C.c_double
        L.orc_ensemble_defect.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_char_p]
        L.orc_enumerate.restype = C.c_double
        L.orc_enumerate.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_enumerate_band.restype = C.c_int
        L.orc_enumerate_band.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_fold_batch.restype = C.c_int
        L.orc_fold_batch.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_fold_batch_fast.restype = C.c_int
        L.orc_fold_batch_fast.argtypes = L.orc_fold_batch.argtypes
        L.orc_mfe_fast.restype = C.c_int
        L.orc_mfe_fast.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_char_p]
        L.orc_pf_fast.restype = C.c_double
        L.orc_pf_fast.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        _lib = L
    return _lib


def split(seq):
    """'AAA&CCC' -> ('AAACCC', cut) with cut = 1-based first index of strand B (0: single)."""
    if "&" in seq:
        a, b = seq.split("&")[:2]
        return a + b, len(a) + 1
    return seq, 0


class Oracle:
    def __init__(self, par_path):
        self.P = lib().orc_params_load(par_path.encode())
        if not self.P:
            raise RuntimeError("oracle: cannot load parameter file " + par_path)

    def get(self, name, *idx):
        idx = list(idx) + [0] * (6 - len(idx))
        return lib().orc_params_get(self.P, name.encode(), *idx)

    def eval(self, seq, db):
        s, cut = split(seq)
        db = db.replace("&", "")
        return lib().orc_eval(self.P, s.encode(), len(s), cut, db.encode())

    def mfe(self, seq, nopair=None, counts=False):
        s, cut = split(seq)
        out = C.create_string_buffer(len(s) + 1)
        cnt = (C.c_longlong * 4)()
        mask = bytes(nopair) if nopair is not None else None
        e = lib().orc_mfe(self.P, s.encode(), len(s), cut, mask, out, cnt)
        if counts:
            return e, out.value.decode(), list(cnt)
        return e, out.value.decode()

    def pf(self, seq, bpp=False):
        import numpy as np
        s, cut = split(seq)
        out = (C.c_double * 5)()
        if bpp:
            P = np.zeros((len(s), len(s)), dtype=np.float64)
            lib().orc_pf(self.P, s.encode(), len(s), cut, out, P.ctypes.data_as(C.POINTER(C.c_double)))
            return list(out), P
        lib().orc_pf(self.P, s.encode(), len(s), cut, out, None)
        return list(out)

    def ensemble_defect(self, bpp, db):
        import numpy as np
        bpp = np.ascontiguousarray(bpp, dtype=np.float64)
        return lib().orc_ensemble_defect(bpp.ctypes.data_as(C.POINTER(C.c_double)), bpp.shape[0], db.encode())

    def enumerate(self, seq, bpp=False):
        import numpy as np
        n = len(seq)
        e1, e2 = C.c_int(), C.c_int()
        P = np.zeros((n, n), dtype=np.float64) if bpp else None
        F = lib().orc_enumerate(self.P, seq.encode(), n, P.ctypes.data_as(C.POINTER(C.c_double)) if bpp else None,
                                C.byref(e1), C.byref(e2))
        return F, P, e1.value, e2.value

    def enumerate_band(self, seq, bound, cap=20000):
        """brute force: sorted [(energy_dcal, structure)] of every structure with energy <= bound"""
        import numpy as np
        n = len(seq)
        en = np.zeros(cap, np.int32)
        ss = C.create_string_buffer(cap * (n + 1))
        cnt = lib().orc_enumerate_band(self.P, seq.encode(), n, int(bound), cap, en.ctypes.data, C.addressof(ss))
        assert cnt <= cap, "band larger than cap"
        raw = ss.raw
        return sorted((int(en[k]), raw[k * (n + 1): k * (n + 1) + n].decode()) for k in range(cnt))

    def mfe_fast(self, seq):
        out = C.create_string_buffer(len(seq) + 1)
        e = lib().orc_mfe_fast(self.P, seq.encode(), len(seq), out)
        return e, out.value.decode()

    def pf_fast(self, seq):
        return lib().orc_pf_fast(self.P, seq.encode(), len(seq))

    def fold_batch(self, seqs, targets=None, nthreads=1, fast=False):
        """seqs: list of equal-length strings -> (mfe[int32], ss[list], epf[f64], ed[int32]); fast: the tuned CPU arm"""
        import numpy as np
        B, n = len(seqs), len(seqs[0])
        mfe = np.zeros(B, np.int32); ed = np.zeros(B, np.int32); epf = np.zeros(B, np.float64)
        ss = C.create_string_buffer(B * (n + 1))
        tg = "".join(targets).encode() if targets is not None else None
        (lib().orc_fold_batch_fast if fast else lib().orc_fold_batch)(self.P, "".join(seqs).encode(), tg, B, n, nthreads,
                             mfe.ctypes.data, C.addressof(ss), epf.ctypes.data, ed.ctypes.data)
        raw = ss.raw
        return mfe, [raw[b * (n + 1): b * (n + 1) + n].decode() for b in range(B)], epf, ed
