"""device-timed step with the outside pass (CUDA events; inputs resident): python scripts/out_time_dev.py <tag>"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from desirna_b200 import engine as eng
tag = sys.argv[1]; B = int(os.environ.get("OUT_B", "4096"))
eng.init(0); eng.params_builtin(1999)
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st)
lut = torch.tensor(list(b"ACGU"), dtype=torch.uint8, device=dev)
for L in [int(x) for x in os.environ.get("OUT_LS", "50,75,100").split(",")]:
    rng = np.random.default_rng(20240000 + L)
    seq = lut[torch.from_numpy(rng.integers(0, 4, (B, L))).to(dev)].contiguous()
    lens = torch.full((B,), L, dtype=torch.int32, device=dev)
    mfe = torch.zeros(B, dtype=torch.int32, device=dev); ss = torch.zeros((B, L + 1), dtype=torch.uint8, device=dev)
    pf = torch.zeros((B, 5), dtype=torch.float64, device=dev); ev = torch.zeros((B, 1), dtype=torch.int32, device=dev)
    tg = torch.full((B, 1, L), ord("."), dtype=torch.uint8, device=dev); dfc = torch.zeros(B, dtype=torch.float64, device=dev)
    W0 = eng.WANT_MFE | eng.WANT_SS | eng.WANT_PF | eng.WANT_EVAL
    def step(w, d=None):
        eng.score_batch_device(seq, lens, w, targets=tg, mfe=mfe, ss=ss, pf=pf, ev=ev, stream=st.cuda_stream, defect=d)
    step(W0); torch.cuda.synchronize(); tg[:, 0, :] = ss[:, :L]
    res = []
    for w, d in ((W0 | eng.WANT_DEFECT, dfc), (W0, None)):
        for _ in range(2): step(w, d)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): step(w, d)
        e1.record(); e1.synchronize()
        res.append(e0.elapsed_time(e1) / 3)
    print(f"{tag} L={L} B={B}: with defect {res[0]:.3f} ms, without {res[1]:.3f} ms, outside pass {res[0] - res[1]:.3f} ms, defect mean {float(dfc.mean()):.12f}", flush=True)
