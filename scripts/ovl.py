"""step time of one big batch through bf_score_batch_device: python scripts/ovl.py <tag> [L] [B]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from desirna_b200 import engine as eng
tag = sys.argv[1]; L = int(sys.argv[2]) if len(sys.argv) > 2 else 100; B = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
eng.init(0); eng.params_builtin(1999)
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st)
rng = np.random.default_rng(20240000 + L)
codes = rng.integers(0, 4, (B, L))
lut = torch.tensor(list(b"ACGU"), dtype=torch.uint8, device=dev)
seq = lut[torch.from_numpy(codes).to(dev)].contiguous()
lens = torch.full((B,), L, dtype=torch.int32, device=dev)
mfe = torch.zeros(B, dtype=torch.int32, device=dev); ss = torch.zeros((B, L + 1), dtype=torch.uint8, device=dev)
pf = torch.zeros((B, 5), dtype=torch.float64, device=dev); ev = torch.zeros((B, 1), dtype=torch.int32, device=dev)
tg = torch.full((B, 1, L), ord("."), dtype=torch.uint8, device=dev)
WANT = eng.WANT_MFE | eng.WANT_SS | eng.WANT_PF | eng.WANT_EVAL
def step():
    eng.score_batch_device(seq, lens, WANT, targets=tg, mfe=mfe, ss=ss, pf=pf, ev=ev, stream=st.cuda_stream)
for _ in range(3): step()
torch.cuda.synchronize()
ref = (mfe.clone(), pf.clone())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 5 if L <= 200 else 2
e0.record()
for _ in range(n): step()
e1.record(); e1.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"{tag} L={L} B={B}: {ms:.3f} ms per step = {B / ms * 1e3 / 1e3:.1f} k folds/s  kernel_ms {[round(x, 3) for x in eng.last_kernel_ms()]} mfe_sum {int(mfe.sum())} epf_sum {float(pf[:, 4].sum()):.6f}", flush=True)
