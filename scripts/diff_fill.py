#!/usr/bin/env python3
"""GPU debugging aid: run the diagonal-major and the tile-wavefront fill kernels on the same batch and
compare their DP tables cell by cell (first mismatches per table, decoded to (i,j))."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from desirna_b200 import engine as eng

def tri_off(n, d): return (d - 4) * n - (d * (d - 1) // 2 - 6)

def decode(n, o):
    for d in range(4, n):
        if tri_off(n, d) <= o < tri_off(n, d) + n - d:
            i = o - tri_off(n, d) + 1
            return i, i + d
    return None

def run(seqs, want, fill):
    eng.set_option("fill", fill)
    out = eng.score_batch(seqs, want=want)
    print("fill", fill, "kernel ms (mfe, pf, eval):", [round(x, 3) for x in eng.last_kernel_ms()])
    tabs = [eng.debug_table(w, len(seqs)) for w in ((0, 1) if not (want & eng.WANT_PF) else (0, 1, 2))]
    return out, tabs

def main():
    rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
    lens = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [5, 8, 9, 12, 13, 17, 23, 30, 37, 41, 50, 64, 77, 100, 100, 100]
    want = eng.WANT_MFE | eng.WANT_SS
    if len(sys.argv) > 3 and sys.argv[3] == "pf": want |= eng.WANT_PF
    seqs = ["".join("ACGU"[k] for k in rng.integers(0, 4, L)) for L in lens]
    ref, rt = run(seqs, want, 0)
    new, nt = run(seqs, want, 1)
    bad = 0
    nmax = max(lens)
    for k, s in enumerate(seqs):
        n = len(s)
        for w, name in enumerate(["c", "fML", "qb"][:len(rt)]):
            # tables are laid out for the sequence's own n
            size = tri_off(n, n) if n >= 5 else 0
            a, b = rt[w][k, :size], nt[w][k, :size]
            if name == "qb":
                diff = np.nonzero(~np.isclose(a, b, rtol=1e-11, atol=0))[0]
            else:
                diff = np.nonzero(a != b)[0]
            if len(diff):
                bad += 1
                print(f"seq {k} n={n} table {name}: {len(diff)} of {size} cells differ; first:")
                for o in diff[:6]:
                    i, j = decode(n, int(o))
                    print(f"   (i={i}, j={j}, d={j-i}) tile I={(i-1)//4} J={(j-1)//4} a={(i-1)%4} b={(j-1)%4}: diag={a[o]} tile={b[o]}")
        if ref["mfe_dcal"][k] != new["mfe_dcal"][k] or ref["mfe_ss"][k] != new["mfe_ss"][k]:
            print(f"seq {k} n={n}: mfe {ref['mfe_dcal'][k]} vs {new['mfe_dcal'][k]}")
        if want & eng.WANT_PF and abs(ref["pf"][k, 4] - new["pf"][k, 4]) > 1e-9 * max(1, abs(ref["pf"][k, 4])):
            print(f"seq {k} n={n}: F {ref['pf'][k, 4]!r} vs {new['pf'][k, 4]!r}")
    print("tables compared:", len(seqs), "sequences; mismatching tables:", bad)
    eng.set_option("fill", 0)

if __name__ == "__main__":
    main()
