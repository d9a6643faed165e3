#!/usr/bin/env python3
"""GPU debugging aid: blocked split on/off (BF_BLK) on the same batch; compares c / fML / qb tables and the scalar results."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from desirna_b200 import engine as eng

def tri_off(n, d): return (d - 4) * n - (d * (d - 1) // 2 - 6)

def run(seqs, blk):
    os.environ["BF_BLK"] = "1" if blk else "0"
    os.environ["BF_BLK_MIN"] = "1"; os.environ["BF_BLK_MIN_PF"] = "1"
    out = eng.score_batch(seqs, want=eng.WANT_MFE | eng.WANT_SS | eng.WANT_PF)
    print("blk", blk, "kernel ms:", [round(x, 3) for x in eng.last_kernel_ms()])
    return out, [eng.debug_table(w, len(seqs)) for w in (0, 1, 2)]

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
lens = [int(x) for x in sys.argv[2].split(",")]
seqs = ["".join("ACGU"[k] for k in rng.integers(0, 4, L)) for L in lens]
eng.set_option("fill", 0)
a, ta = run(seqs, False)
b, tb = run(seqs, True)
bad = 0
for k, s in enumerate(seqs):
    n = len(s); size = tri_off(n, n) if n >= 5 else 0
    for w, name in enumerate(["c", "fML", "qb"]):
        x, y = ta[w][k, :size], tb[w][k, :size]
        diff = np.nonzero(x != y)[0] if w < 2 else np.nonzero(~np.isclose(x, y, rtol=1e-12, atol=0))[0]
        if len(diff):
            bad += 1
            o = int(diff[0])
            d = next(d for d in range(4, n) if tri_off(n, d) <= o < tri_off(n, d) + n - d)
            i = o - tri_off(n, d) + 1
            print(f"seq {k} n={n} {name}: {len(diff)} differ; first (i={i}, j={i+d}, d={d}) I={(i-1)//4} J={(i+d-1)//4}: {x[o]!r} vs {y[o]!r}")
    if a["mfe_dcal"][k] != b["mfe_dcal"][k] or a["mfe_ss"][k] != b["mfe_ss"][k]: print("mfe differs", k); bad += 1
    if abs(a["pf"][k, 4] - b["pf"][k, 4]) > 1e-10 * max(1, abs(a["pf"][k, 4])): print("F differs", k, a["pf"][k, 4], b["pf"][k, 4]); bad += 1
print("sequences:", len(seqs), "mismatches:", bad)
