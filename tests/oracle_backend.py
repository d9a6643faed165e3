"""TEST INFRASTRUCTURE: an engine-shaped checker backend built on the CPU oracle.

oracle_backend.install(OracleBackend(...)) swaps the module attribute desirna_b200.RNA._eng (the product has no backend
switch of its own) so that the CPU test-suite can exercise the host-side mirror
(score arithmetic, similarity scores, pseudoknot overlay, batching, replica loop) without a GPU, and gives
the GPU tests a second implementation of the same score_batch() contract to compare with.  Never imported
by the product package."""
import numpy as np

from oracle.pyoracle import Oracle


def current():
    from desirna_b200 import RNA
    return RNA._eng


def install(backend):
    """test seam: replace the engine module the RNA shim calls; returns the previous one"""
    from desirna_b200 import RNA
    old = RNA._eng
    RNA._eng = backend
    return old


class OracleBackend:
    WANT_MFE, WANT_SS, WANT_PF, WANT_EVAL, WANT_BPP, WANT_DEFECT = 1, 2, 4, 8, 16, 32

    def __init__(self, par_path):
        self.O = Oracle(par_path)
        self.calls = 0
        self.folds = 0

    def params_load(self, path):
        self.O = Oracle(path)

    def score_batch(self, seqs, targets=None, nopair=None, want=7):
        self.calls += 1
        self.folds += len(seqs)
        B = len(seqs)
        out = {"len": np.array([len(s.replace("&", "")) for s in seqs], np.int32)}
        if want & (self.WANT_MFE | self.WANT_SS):
            res = [self.O.mfe(s, nopair=None if nopair is None else nopair[k][:len(s.replace("&", ""))]) for k, s in enumerate(seqs)]
            out["mfe_dcal"] = np.array([r[0] for r in res], np.int32)
            out["mfe_ss"] = [r[1] for r in res]
        if want & (self.WANT_PF | self.WANT_BPP | self.WANT_DEFECT):
            out["pf"] = np.zeros((B, 5))
            if want & (self.WANT_BPP | self.WANT_DEFECT):
                out["bpp"], out["defect"] = [], np.zeros(B)
            for k, s in enumerate(seqs):
                if want & (self.WANT_BPP | self.WANT_DEFECT):
                    pf, bpp = self.O.pf(s, bpp=True)
                    out["bpp"].append(bpp)
                    if targets is not None:
                        out["defect"][k] = self.O.ensemble_defect(bpp, targets[k][0].replace("&", ""))
                else:
                    pf = self.O.pf(s)
                out["pf"][k] = pf
        if targets is not None and (want & self.WANT_EVAL):
            out["eval_dcal"] = np.array([[self.O.eval(s, t) for t in targets[k]] for k, s in enumerate(seqs)], np.int32)
        return out

    def subopt(self, seq, delta_dcal, nopair=None, max_out=4096):
        """brute force (short sequences only): same contract as engine.subopt"""
        s = seq.upper().replace("T", "U")
        assert len(s) <= 20 and nopair is None
        mfe = self.O.mfe(s)[0]
        band = self.O.enumerate_band(s, mfe + int(delta_dcal))
        return [(ss, e) for e, ss in band[:max_out]], len(band) > max_out

    def second_best(self, seqs, nopair=None):
        """brute force (short sequences only): same contract as engine.second_best"""
        assert nopair is None
        e1, e2 = np.zeros(len(seqs), np.int32), np.zeros(len(seqs), np.int32)
        for k, seq in enumerate(seqs):
            s = seq.upper().replace("T", "U")
            assert len(s) <= 20
            band = self.O.enumerate_band(s, 10 ** 6)
            e1[k] = band[0][0]
            e2[k] = band[1][0] if len(band) > 1 else 10000000
        return e1, e2
