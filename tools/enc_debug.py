import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_helpers import models, state_dicts, err_stats
from oracle import networks_oracle as O, weights as W
torch.set_num_threads(32)
enc, _ = models(); esd, _ = state_dicts()
B, L = int(sys.argv[1]), int(sys.argv[2])
x = W.synthetic_audio(B, L, seed=52)
with torch.no_grad():
    ref = O.fxencoder_forward(x, esd, W.ENC_KERNELS, W.ENC_STRIDES)
    got = enc(x.cuda()).cpu()
e = err_stats(got, ref)
d = (got - ref).numpy()
print(f"first_umma={os.environ.get('MST_ENC_FIRST_UMMA')} B={B} L={L}: rms {e['rms']:.3e} max {e['max']:.3e} rel {e['rel']:.3e}; err rms per batch row {np.sqrt((d**2).mean(1))}")
