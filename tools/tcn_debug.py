"""GPU diagnostic for the tcgen05 TCN block kernel: error structure of one block vs the CPU oracle.
   python tools/tcn_debug.py [block] [L] [B]"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_helpers import models, state_dicts
from oracle import fixtures, networks_oracle as O

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
L = int(sys.argv[2]) if len(sys.argv) > 2 else 512
B = int(sys.argv[3]) if len(sys.argv) > 3 else 1
_, tcn = models()
_, tsd = state_dicts()
g = torch.Generator(); g.manual_seed(1)
x = torch.randn(B, 128 if n else 2, L, generator=g) * 0.5
cond = fixtures.make_cond(1, 2)
with torch.no_grad():
    ref = O.tcn_block(x, cond, tsd, f"blocks.{n}", 15, 2 ** n).numpy()
    got = tcn.blocks[n](x.cuda(), cond.cuda()).cpu().numpy()
d = got - ref
print(f"block {n} L={L} B={B}: rms err {np.sqrt((d**2).mean()):.3e} max {np.abs(d).max():.3e} ref rms {np.sqrt((ref**2).mean()):.3e}")
print("nan/inf:", np.isnan(got).sum(), np.isinf(got).sum())
print("err rms per 16-channel group:", np.round(np.sqrt((d**2).mean(axis=(0, 2)).reshape(8, 16).mean(1)), 6))
T = min(L, 512)
e_t = np.sqrt((d[0, :, :T]**2).mean(0))
print("err rms per 32-row group (first 512 rows):", np.round(e_t.reshape(-1, 32).mean(1) if T % 32 == 0 else e_t[:T//32*32].reshape(-1, 32).mean(1), 6))
print("err rms by row mod 8:", np.round([np.sqrt((d[0, :, r::8]**2).mean()) for r in range(8)], 6))
print("got[0,0:4,0:6]\n", got[0, 0:4, 0:6], "\nref[0,0:4,0:6]\n", ref[0, 0:4, 0:6])
print("got[0,64:68,0:6]\n", got[0, 64:68, 0:6], "\nref\n", ref[0, 64:68, 0:6])
# correlation helps spot a transposed / permuted result
print("corr(got, ref) =", float(np.corrcoef(got.ravel(), ref.ravel())[0, 1]))
