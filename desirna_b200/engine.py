"""ctypes binding of libb200fold.so (C-ABI: include/b200fold.h).

The library is the product: there is no CPU fallback.  Importing this module without the
built shared object raises; calling into it without a B200 returns BF_ERR_CUDA, surfaced
as EngineError.
"""
import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libb200fold.so")
PARAMS_DIR = os.path.join(_PKG, "params")

WANT_MFE, WANT_SS, WANT_PF, WANT_EVAL, WANT_BPP, WANT_DEFECT = 1, 2, 4, 8, 16, 32
INF_DCAL = 10000000


class EngineError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"b200fold error {code}: {msg}")
        self.code = code


class bf_batch_t(C.Structure):
    _fields_ = [("B", C.c_int32), ("stride", C.c_int32), ("seq", C.c_void_p), ("len", C.c_void_p),
                ("cut", C.c_void_p), ("nopair", C.c_void_p), ("targets", C.c_void_p),
                ("n_targets", C.c_int32), ("want", C.c_uint32)]


class bf_result_t(C.Structure):
    _fields_ = [("mfe_dcal", C.c_void_p), ("mfe_ss", C.c_void_p), ("pf", C.c_void_p),
                ("eval_dcal", C.c_void_p), ("defect", C.c_void_p), ("bpp", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(the engine has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.bf_last_error.restype = C.c_char_p
        L.bf_init.argtypes = [C.c_int]
        L.bf_params_load.argtypes = [C.c_char_p]
        L.bf_params_builtin.argtypes = [C.c_int, C.c_char_p]
        L.bf_params_get.argtypes = [C.c_char_p] + [C.c_int] * 6 + [C.POINTER(C.c_int32)]
        L.bf_score_batch.argtypes = [C.POINTER(bf_batch_t), C.POINTER(bf_result_t)]
        L.bf_score_batch_device.argtypes = [C.POINTER(bf_batch_t), C.POINTER(bf_result_t), C.c_void_p]
        L.bf_kernel_launches.restype = C.c_int64
        L.bf_last_kernel_ms.argtypes = [C.POINTER(C.c_double)]
        L.bf_microbench.argtypes = [C.POINTER(C.c_double)]
        L.bf_second_best.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.bf_subopt.argtypes = [C.c_char_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.bf_set_option.argtypes = [C.c_char_p, C.c_int]
        L.bf_debug_copy_table.argtypes = [C.c_int, C.c_int32, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise EngineError(rc, lib().bf_last_error().decode())


_state = {"inited": False, "params": None}


def init(device=None):
    """Select the GPU (default: LOCAL_RANK or 0) and create the engine."""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    _check(lib().bf_init(int(device)))
    _state["inited"] = True


def shutdown():
    lib().bf_shutdown()
    _state["inited"] = False


def params_load(path):
    """RNA.params_load(path) (DesiRNA.py:456)."""
    _check(lib().bf_params_load(os.fsencode(path)))
    _state["params"] = path


def params_builtin(year=1999):
    """`-p 1999|2004` selection of DesiRNA.py:455; 2004 raises EngineError(BF_ERR_UNAVAILABLE)."""
    _check(lib().bf_params_builtin(int(year), os.fsencode(PARAMS_DIR)))
    _state["params"] = f"builtin:{year}"


def params_get(name, *idx):
    idx = list(idx) + [0] * (6 - len(idx))
    out = C.c_int32()
    _check(lib().bf_params_get(name.encode(), *idx, C.byref(out)))
    return out.value


def ensure_ready(device=None):
    if not _state["inited"]:
        init(device)
    if _state["params"] is None:
        params_builtin(1999)


def kernel_launches():
    return int(lib().bf_kernel_launches())


def last_kernel_ms():
    """(mfe_ms, pf_ms, eval_ms) of the most recent call, CUDA-event timed on the launch stream."""
    out = (C.c_double * 3)()
    _check(lib().bf_last_kernel_ms(out))
    return list(out)


def subopt(seq, delta_dcal, nopair=None, max_out=4096):
    """bf_subopt: every structure of one single-strand sequence within delta_dcal of the MFE, sorted by energy.
    Returns (list of (structure, energy_dcal), truncated)."""
    ensure_ready()
    s = seq.upper().replace("T", "U").encode("ascii")
    n = len(s)
    ss = np.zeros((max_out, n + 1), np.uint8)
    en = np.zeros(max_out, np.int32)
    cnt, trunc = C.c_int32(0), C.c_int32(0)
    mask = None
    if nopair is not None:
        mask = np.ascontiguousarray(nopair, np.uint8)
        assert mask.shape == (n,)
    _check(lib().bf_subopt(s, n, mask.ctypes.data if mask is not None else None, int(delta_dcal), max_out, ss.ctypes.data, en.ctypes.data,
                           C.byref(cnt), C.byref(trunc)))
    return [(bytes(ss[k, :n]).decode("ascii"), int(en[k])) for k in range(cnt.value)], bool(trunc.value)


def second_best(seqs, nopair=None):
    """bf_second_best: (e1, e2) int32 arrays, dcal/mol -- energies of the best and the second-best structure of every sequence
    (single strands); e2 >= 10000000 where a sequence has no second structure."""
    ensure_ready()
    buf, lens, cuts = pack([s.upper().replace("T", "U") for s in seqs])
    if cuts.any():
        raise EngineError(3, "second_best: single-strand sequences only")
    b = bf_batch_t()
    b.B, b.stride = len(seqs), buf.shape[1]
    b.seq, b.len = buf.ctypes.data, lens.ctypes.data
    keep = [buf, lens]
    if nopair is not None:
        nopair = np.ascontiguousarray(nopair, np.uint8)
        assert nopair.shape == buf.shape
        b.nopair = nopair.ctypes.data
        keep.append(nopair)
    e1, e2 = np.zeros(len(seqs), np.int32), np.zeros(len(seqs), np.int32)
    _check(lib().bf_second_best(C.byref(b), e1.ctypes.data, e2.ctypes.data))
    return e1, e2


def set_option(key, value):
    """Tuning / test hook (bf_set_option): sets the engine's BF_<KEY> variable, e.g. set_option("cl", 0) switches the
    cluster-per-sequence fill kernels off, set_option("cl", -1) restores the default rule."""
    ensure_ready()
    _check(lib().bf_set_option(key.encode(), int(value)))


def debug_table(which, n_seq):
    """Test hook: the packed diagonal-major DP table (0 = c, 1 = fML, 2 = qb) of the most recent call, [n_seq, slot]."""
    slot = C.c_size_t(0)
    _check(lib().bf_debug_copy_table(which, n_seq, None, 0, C.byref(slot)))
    out = np.zeros((n_seq, slot.value), np.float64 if which == 2 else np.int32)
    _check(lib().bf_debug_copy_table(which, n_seq, out.ctypes.data, out.nbytes, C.byref(slot)))
    return out


def microbench():
    """{'int32_ops', 'fp64_flops', 'smem_bytes'} per second, measured on the current GPU."""
    ensure_ready()
    out = (C.c_double * 3)()
    _check(lib().bf_microbench(out))
    return {"int32_ops_per_s": out[0], "fp64_flops_per_s": out[1], "smem_bytes_per_s": out[2]}


def score_batch_device(seq, lens, want, cut=None, nopair=None, targets=None, mfe=None, ss=None, pf=None, ev=None, stream=0, defect=None,
                       bpp=None):
    """Device-resident entry point (bf_score_batch_device).  Arguments are torch CUDA tensors:
    seq uint8[B,stride], lens int32[B], cut int32[B], nopair uint8[B,stride], targets uint8[B,T,stride];
    outputs mfe int32[B], ss uint8[B,stride+1], pf float64[B,5], ev int32[B,T].  Enqueues on `stream`
    (a raw cudaStream_t handle, e.g. torch.cuda.current_stream().cuda_stream) without synchronising."""
    ensure_ready()
    b, r = bf_batch_t(), bf_result_t()
    b.B, b.stride = int(seq.shape[0]), int(seq.shape[1])
    b.seq, b.len = seq.data_ptr(), lens.data_ptr()
    b.cut = cut.data_ptr() if cut is not None else None
    b.nopair = nopair.data_ptr() if nopair is not None else None
    if targets is not None:
        b.targets, b.n_targets = targets.data_ptr(), int(targets.shape[1])
    b.want = want
    r.mfe_dcal = mfe.data_ptr() if mfe is not None else None
    r.mfe_ss = ss.data_ptr() if ss is not None else None
    r.pf = pf.data_ptr() if pf is not None else None
    r.eval_dcal = ev.data_ptr() if ev is not None else None
    r.defect = defect.data_ptr() if defect is not None else None   # float64[B]
    r.bpp = bpp.data_ptr() if bpp is not None else None            # float64[B,stride,stride]
    _check(lib().bf_score_batch_device(C.byref(b), C.byref(r), C.c_void_p(stream)))


def sm_count():
    return int(lib().bf_sm_count())


def pack(seqs, stride=None):
    """list of str (optionally 'A&B') -> (chars[B,stride] uint8, len[B] int32, cut[B] int32)."""
    B = len(seqs)
    clean, cuts = [], np.zeros(B, np.int32)
    for k, s in enumerate(seqs):
        if "&" in s:
            a, b = s.split("&")[:2]
            cuts[k] = len(a) + 1
            s = a + b
        clean.append(s)
    lens = np.array([len(s) for s in clean], np.int32)
    if stride is None:
        stride = max(1, int(lens.max()) if B else 1)
    # one join + one frombuffer instead of a numpy assignment per row (rows padded with NUL to the stride)
    flat = "".join(s if len(s) == stride else s.ljust(stride, "\0") for s in clean).encode("ascii")
    buf = np.frombuffer(flat, np.uint8).reshape(B, stride).copy() if B else np.zeros((0, stride), np.uint8)
    return buf, lens, cuts


def pack_targets(targets, stride):
    """targets: list (per sequence) of list of dot-bracket strings ('&' dropped) -> uint8[B,T,stride]"""
    B = len(targets)
    T = len(targets[0]) if B else 0
    rows = []
    for row in targets:
        assert len(row) == T, "every sequence needs the same number of targets"
        for db in row:
            if "&" in db:
                db = db.replace("&", "")
            rows.append(db if len(db) == stride else db.ljust(stride, "."))
    if not rows:
        return np.full((B, T, stride), ord("."), np.uint8)
    return np.frombuffer("".join(rows).encode("ascii"), np.uint8).reshape(B, T, stride).copy()


def score_batch(seqs, targets=None, nopair=None, want=WANT_MFE | WANT_SS | WANT_PF):
    """Host-buffer entry point (bf_score_batch).  seqs: list of str; targets: list of lists of
    dot-bracket strings; nopair: optional uint8[B, stride] mask.  Returns a dict of numpy arrays."""
    ensure_ready()
    B = len(seqs)
    if B == 0:
        return {"len": np.zeros(0, np.int32), "cut": np.zeros(0, np.int32), "mfe_dcal": np.zeros(0, np.int32),
                "mfe_ss": [], "pf": np.zeros((0, 5)), "eval_dcal": np.zeros((0, 0), np.int32)}
    buf, lens, cuts = pack(seqs)
    stride = buf.shape[1]
    b, r = bf_batch_t(), bf_result_t()
    b.B, b.stride = B, stride
    b.seq, b.len = buf.ctypes.data, lens.ctypes.data
    b.cut = cuts.ctypes.data if cuts.any() else None
    keep = [buf, lens, cuts]
    if nopair is not None:
        nopair = np.ascontiguousarray(nopair, np.uint8)
        assert nopair.shape == (B, stride)
        b.nopair = nopair.ctypes.data
        keep.append(nopair)
    out = {}
    if targets is not None and (want & (WANT_EVAL | WANT_DEFECT)):
        tb = pack_targets(targets, stride)
        b.targets, b.n_targets = tb.ctypes.data, tb.shape[1]
        keep.append(tb)
        if want & WANT_EVAL:
            out["eval_dcal"] = np.zeros((B, tb.shape[1]), np.int32)
            r.eval_dcal = out["eval_dcal"].ctypes.data
    else:
        want &= ~(WANT_EVAL | WANT_DEFECT)
    if want & (WANT_BPP | WANT_DEFECT):
        want |= WANT_MFE | WANT_PF   # the MFE sets the partition function's scale (energy_scores.py:371-372)
    if want & WANT_DEFECT:
        out["defect"] = np.zeros(B, np.float64)
        r.defect = out["defect"].ctypes.data
    if want & WANT_BPP:
        out["bpp"] = np.zeros((B, stride, stride), np.float64)
        r.bpp = out["bpp"].ctypes.data
    if want & (WANT_MFE | WANT_SS):
        want |= WANT_MFE
        out["mfe_dcal"] = np.zeros(B, np.int32)
        r.mfe_dcal = out["mfe_dcal"].ctypes.data
    if want & WANT_SS:
        ss = np.zeros((B, stride + 1), np.uint8)
        r.mfe_ss = ss.ctypes.data
    if want & WANT_PF:
        out["pf"] = np.zeros((B, 5), np.float64)
        r.pf = out["pf"].ctypes.data
    b.want = want
    _check(lib().bf_score_batch(C.byref(b), C.byref(r)))
    if want & WANT_SS:
        raw, w = ss.tobytes().decode("ascii"), stride + 1
        out["mfe_ss"] = [raw[k * w:k * w + int(lens[k])] for k in range(B)]
    out["len"], out["cut"] = lens, cuts
    return out
