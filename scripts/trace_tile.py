#!/usr/bin/env python3
"""Per-phase, per-warp timeline of the tile kernel (needs a build with -DBF_TILE_TRACE)."""
import sys, os, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from desirna_b200 import engine as eng
L = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(3)
seqs = ["".join("ACGU"[k] for k in rng.integers(0, 4, L)) for _ in range(4096)]
eng.set_option("fill", 1)
for _ in range(2):
    eng.score_batch(seqs, want=eng.WANT_MFE | eng.WANT_SS)
buf = np.zeros(128 * 16 * 4, np.int64)
rc = eng.lib().bf_tile_trace(C.c_void_p(buf.ctypes.data)); assert rc == 0
t = buf.reshape(128, 16, 4)
NT = (L + 3) // 4
nw = int(os.environ.get("BF_TILE_NW", 8))
print("phase  tiles | step1 per warp (cycles)          | wait1 max | step2 per warp | phase total")
for D in range(1, NT):
    x = t[D, :nw]
    t0 = x[:, 0].min()
    s1 = x[:, 1] - x[:, 0]; s2 = x[:, 3] - x[:, 2]
    print(f"{D:3d} {NT-D:4d} | " + " ".join(f"{v:6d}" for v in s1) + " | " + " ".join(f"{v:6d}" for v in s2) + f" | {x[:,3].max()-t0:7d}")
tot = t[NT-1, :nw, 3].max() - t[1, :nw, 0].min()
print("whole sequence cycles:", tot)
