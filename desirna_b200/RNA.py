"""ViennaRNA-shaped front end of the B200 engine: the subset of the SWIG module `RNA` that DesiRNA calls.

DesiRNA has no FFI of its own; its hot path enters native code through `import RNA`
(utils/energy_scores.py:20, utils/dimer_multichain_energy.py:27, utils/sequence_utils.py:46,
DesiRNA.py:35).  A maintainer switches engines with `from desirna_b200 import RNA`.  Call sites covered:

    RNA.params_load(path)                         DesiRNA.py:456
    RNA.md(), md.compute_bpp                      energy_scores.py:27-28, 369
    RNA.fold_compound(seq[, md])                  energy_scores.py:147, 370, 416; dimer_multichain_energy.py:89, 104
    fc.pf() / fc.mfe() / fc.eval_structure(db)    energy_scores.py:75, 99, 150-151, 371-373
    fc.mfe_dimer() / fc.pf_dimer()                energy_scores.py:156-157; dimer_multichain_energy.py:47, 90, 105
    fc.hc_add_from_db(db)                         sequence_utils.py:1181, 1198, 1214
    fc.exp_params_rescale(mfe)                    energy_scores.py:372   (numerical no-op here: the kernels rescale per sequence)
    fc.ensemble_defect(db)                        energy_scores.py:374
    RNA.fold(seq) / RNA.co_pf_fold(seq)           energy_scores.py:354; dimer_multichain_energy.py:109, 113-114

Conventions kept: energies come back as Python floats holding float32-rounded kcal/mol (ViennaRNA returns C
float), structures from mfe_dimer() carry no '&', pf() returns (string, energy) and DesiRNA ignores element 0.
Results of one fold compound are cached, and a whole batch of compounds can be evaluated by ONE engine call
through `prefetch()` -- that is how the lock-step replica loop keeps the GPU busy while callers still see
the per-object API.  There is no CPU fallback: every number comes from libb200fold.so.
"""
import struct

import numpy as np

from . import engine as _eng


def f32(x):
    """round to C float, widen back to double (what SWIG hands to Python)"""
    return struct.unpack("f", struct.pack("f", float(x)))[0]


class _Cvar:
    uniq_ML = 0


cvar = _Cvar()


def params_load(path):
    _eng.params_load(path)
    return 1


class md:
    """model details: only the fields DesiRNA touches; the rest are ViennaRNA's defaults baked into the kernels
    (T=37C, dangles=2, noLP=0, noGU=0, special hairpins, TURN=3, MAXLOOP=30, pf_smooth=1)."""

    def __init__(self):
        self.compute_bpp = 1
        self.temperature = 37.0
        self.dangles = 2
        self.uniq_ML = 0


class fold_compound:
    def __init__(self, sequence, model_details=None, *_):
        self.sequence = sequence
        self.md = model_details if model_details is not None else md()
        self.length = len(sequence.replace("&", ""))
        self._nopair = None          # accumulated hard constraints ('x' positions)
        self._cache = {}

    # ---- batching hook -------------------------------------------------------------------------
    @staticmethod
    def prefetch(compounds, targets=None, want=None):
        """Evaluate many fold compounds with one engine call and fill their caches.
        targets: optional list (per compound) of lists of dot-bracket strings for eval_structure."""
        if not compounds:
            return
        E = _eng
        w = want if want is not None else (E.WANT_MFE | E.WANT_SS | E.WANT_PF)
        seqs = [c.sequence for c in compounds]
        tg = None
        if targets is not None:
            T = max(len(t) for t in targets)
            if T:
                tg = [list(t) + [t[-1] if t else "." * c.length] * (T - len(t)) for t, c in zip(targets, compounds)]
                w |= E.WANT_EVAL
        out = E.score_batch(seqs, tg, want=w)
        for k, c in enumerate(compounds):
            if "mfe_dcal" in out:
                c._cache["mfe"] = (out["mfe_ss"][k], int(out["mfe_dcal"][k])) if "mfe_ss" in out else (None, int(out["mfe_dcal"][k]))
            if "pf" in out:
                c._cache["pf"] = [float(x) for x in out["pf"][k]]
            if tg is not None:
                for t, db in enumerate(targets[k]):
                    c._cache[("eval", db.replace("&", ""))] = int(out["eval_dcal"][k, t])

    @staticmethod
    def prefetch_constrained(compounds):
        """One engine call for the constrained MFE folds (hc_add_from_db 'x' masks) of many compounds: the refolds of one
        round of the pseudoknot overlay (sequence_utils.py:1181-1216) across all mutants of a Monte-Carlo sub-step."""
        todo = [c for c in compounds if c._nopair is not None and ("mfe", bytes(c._nopair)) not in c._cache]
        if not todo:
            return
        E = _eng
        stride = max(c.length for c in todo)
        mask = np.zeros((len(todo), stride), np.uint8)
        for k, c in enumerate(todo):
            mask[k, :c.length] = c._nopair
        out = E.score_batch([c.sequence for c in todo], None, nopair=mask, want=E.WANT_MFE | E.WANT_SS)
        for k, c in enumerate(todo):
            c._cache[("mfe", bytes(c._nopair))] = (out["mfe_ss"][k], int(out["mfe_dcal"][k]))

    # ---- single-object API ---------------------------------------------------------------------
    def _run(self, want, targets=None):
        mask = None
        if self._nopair is not None:
            mask = np.asarray(self._nopair, np.uint8)[None, :]
        return _eng.score_batch([self.sequence], targets, nopair=mask, want=want)

    def _mfe(self):
        key = "mfe" if self._nopair is None else ("mfe", bytes(self._nopair))
        if key not in self._cache:
            E = _eng
            out = self._run(E.WANT_MFE | E.WANT_SS)
            self._cache[key] = (out["mfe_ss"][0], int(out["mfe_dcal"][0]))
        return self._cache[key]

    def _pf(self):
        if "pf" not in self._cache:
            if self._nopair is not None and any(self._nopair):
                # the partition-function kernels ignore hard constraints; DesiRNA always calls pf() before hc_add_from_db
                raise RuntimeError("fold_compound.pf() after hc_add_from_db(): constrained partition functions are not supported")
            out = self._run(_eng.WANT_PF)
            self._cache["pf"] = [float(x) for x in out["pf"][0]]
        return self._cache["pf"]

    def mfe(self):
        ss, e = self._mfe()
        return ss, f32(e / 100.0)

    def mfe_dimer(self):
        return self.mfe()

    def pf(self):
        p = self._pf()
        col = 3 if "&" in self.sequence else 4
        return "", f32(p[col])

    def pf_dimer(self):
        p = self._pf()
        if "&" not in self.sequence:
            return "", f32(p[4]), 0.0, 0.0, f32(p[4])
        return "", f32(p[0]), f32(p[1]), f32(p[2]), f32(p[3])

    def eval_structure(self, structure):
        db = structure.replace("&", "")
        key = ("eval", db)
        if key not in self._cache:
            out = self._run(_eng.WANT_EVAL, targets=[[db]])
            self._cache[key] = int(out["eval_dcal"][0, 0])
        return f32(self._cache[key] / 100.0)

    def hc_add_from_db(self, constraint, *_):
        """only 'x' (position may not pair) is used by DesiRNA; constraints accumulate on the compound"""
        db = constraint.replace("&", "")
        if self._nopair is None:
            self._nopair = [0] * self.length
        for i, ch in enumerate(db[:self.length]):
            if ch == "x":
                self._nopair[i] = 1
        return 1

    def exp_params_rescale(self, mfe=None):
        return None

    def bpp(self):
        if "bpp" not in self._cache:
            E = _eng
            out = self._run(E.WANT_MFE | E.WANT_PF | E.WANT_BPP)
            self._cache["bpp"] = out["bpp"][0]
            self._cache["pf"] = [float(x) for x in out["pf"][0]]
        return self._cache["bpp"]

    def ensemble_defect(self, structure):
        """vrna_ensemble_defect: (1/n) sum_i (1 - p_i,pt(i)) for paired i, sum_j p_ij for unpaired i"""
        db = structure.replace("&", "")
        key = ("defect", db)
        if key not in self._cache:
            E = _eng
            out = self._run(E.WANT_MFE | E.WANT_PF | E.WANT_DEFECT, targets=[[db]])
            self._cache[key] = float(out["defect"][0])
        return self._cache[key]

    def second_best_energy(self):
        """(MFE, energy of the second-best structure) in kcal/mol as C floats, or (MFE, None) when the sequence has a single
        structure: bf_second_best, a 2-best DP on the unambiguous grammar -- the number DesiRNA reads off subopt_cb's enumeration
        (energy_scores.py:453-488) without the enumeration.  Extension of the shim; ViennaRNA has no such call."""
        if "&" in self.sequence:
            raise NotImplementedError("second_best_energy on two-strand compounds is not part of the accelerated path")
        nopair = np.asarray(self._nopair, np.uint8)[None, :] if self._nopair is not None and any(self._nopair) else None
        e1, e2 = _eng.second_best([self.sequence], nopair)
        return f32(int(e1[0]) / 100.0), (f32(int(e2[0]) / 100.0) if int(e2[0]) < 10000000 else None)

    def subopt_cb(self, delta, cb, data=None):
        """vrna_subopt_cb as DesiRNA calls it (energy_scores.py:465-474, uniq_ML = 1): cb(structure, energy, data) for every
        structure within `delta` dcal/mol of the MFE, then once more with structure None."""
        if "&" in self.sequence:
            raise NotImplementedError("subopt_cb on two-strand compounds is not part of the accelerated path")
        nopair = np.asarray(self._nopair, np.uint8) if self._nopair is not None and any(self._nopair) else None
        # the walk stops at max_out structures in search order, not energy order: a truncated band is not a band.  Grow the
        # buffer until the whole band fits (ViennaRNA enumerates it completely, whatever its size).
        cap = 4096
        while True:
            found, truncated = _eng.subopt(self.sequence, int(delta), nopair, max_out=cap)
            if not truncated:
                break
            if cap >= (1 << 22):
                raise RuntimeError("subopt_cb: more than %d structures within %d dcal/mol of the MFE" % (cap, int(delta)))
            cap *= 8
        for structure, e in found:
            cb(structure, f32(e / 100.0), data)
        cb(None, 0.0, data)


def fold(sequence):
    return fold_compound(sequence).mfe()


def co_pf_fold(sequence):
    return fold_compound(sequence).pf_dimer()
