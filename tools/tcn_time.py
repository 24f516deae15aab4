"""Time the 13 dilated TCN launches (config 2 shapes) for the current MST_* env; prints ms per launch."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_helpers import models
from oracle import fixtures, weights as W
_, tcn = models()
x = W.synthetic_audio(32, 262144, seed=3).cuda()
c = fixtures.make_cond(1, 4).cuda()
ev = []
def note(name, phase):
    if name == "tcn_block_umma_kernel":
        e = torch.cuda.Event(enable_timing=True); e.record(); ev.append(e)
with torch.no_grad():
    tcn.forward_layers(x, c)
    ev.clear()
    for _ in range(2):
        tcn.forward_layers(x, c, note)
torch.cuda.synchronize()
d = [ev[i].elapsed_time(ev[i + 1]) for i in range(0, len(ev), 2)]
print({k: os.environ.get(k) for k in ("MST_TCN_PRECISION", "MST_TCN_PIPE", "MST_TCN_MULTICAST", "MST_TCN_DBG", "MST_TCN_PAIRED", "MST_TCN_LOOKAHEAD")}, "ms/launch mean %.3f min %.3f max %.3f" % (sum(d) / len(d), min(d), max(d)))
n = len(d) // 2
print("per block (dilation 2^n, n=1..13), ms:", " ".join("%.2f" % ((d[i] + d[i + n]) / 2) for i in range(n)))
