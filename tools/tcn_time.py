"""Time the 13 dilated TCN launches (config 2 shapes); prints ms per launch.  usage: tcn_time.py [f16f8|bf16x3] [B]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_helpers import models
from music_mixing_style_transfer_b200 import _cabi
if os.environ.get("MST_DEV_LIB"):   # development builds of build.py --variant (ablation studies)
    _cabi.LIB_PATH = os.environ["MST_DEV_LIB"]
from oracle import fixtures, weights as W
_, tcn = models()
tcn.precision = sys.argv[1] if len(sys.argv) > 1 else "f16f8"
BATCH = int(sys.argv[2]) if len(sys.argv) > 2 else 32
x = W.synthetic_audio(BATCH, 262144, seed=3).cuda()
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
print(tcn.precision, "B", BATCH, os.environ.get("MST_DEV_LIB", ""), "ms/launch mean %.3f min %.3f max %.3f" % (sum(d) / len(d), min(d), max(d)))
n = len(d) // 2
print("per block (dilation 2^n, n=1..13), ms:", " ".join("%.2f" % ((d[i] + d[i + n]) / 2) for i in range(n)))
