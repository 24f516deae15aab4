"""RMS error of the TCN against the committed golden vectors of the reference (tests/golden/) for both operand formats:
the small batches and the full-length (262144) segment.  Prints one line per fixture and format."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_helpers import models, err_stats
from oracle import fixtures, weights as W
_, tcn = models()
for mode in ("f16f8", "bf16x3"):
  tcn.precision = mode
  with torch.no_grad():
      for name, B, L, seed, cseed, ncond in (("tcn_small.npz", 2, 8191, 12, 21, 1), ("tcn_percond.npz", 3, 4099, 13, 22, 3)):
          got = tcn(W.synthetic_audio(B, L, seed=seed).cuda(), fixtures.make_cond(ncond, cseed).cuda()).cpu().numpy()
          e = err_stats(got, fixtures.load_golden(name)["y"])
          print(f"{mode}: {name}: rms {e['rms']:.3e} max {e['max']:.3e} ref ac-rms {e['ref_ac_rms']:.3f}")
      y = tcn(W.synthetic_audio(1, 262144, seed=14).cuda(), fixtures.make_cond(1, 23).cuda())[0].cpu().numpy()
      g = fixtures.load_golden("tcn_full.npz")
      win = np.stack([y[:, s:s + n] for s, n in fixtures.FULL_WINDOWS])
      e1, e2 = err_stats(win, g["windows"]), err_stats(y[:, ::fixtures.FULL_STRIDE], g["strided"])
      print(f"{mode}: tcn_full.npz (L=262144): windows rms {e1['rms']:.3e} max {e1['max']:.3e}; strided rms {e2['rms']:.3e} max {e2['max']:.3e}; ac-rms {float(g['ac_rms']):.3f}")
