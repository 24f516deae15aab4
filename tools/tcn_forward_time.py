"""Sustained whole-TCN forward (block 0 + 13 dilated blocks through mst_tcn_forward) at config 2 shapes: ms per forward over a
few seconds, with the SM clock / board power nvidia-smi saw.  The step is energy-bound under the 1000 W cap, so A/B comparisons
of a sub-kernel have to be made on the whole forward, alternating libraries inside one gpurun call.
usage: tcn_forward_time.py [reps] [batch] [length];  MST_DEV_LIB=<build.py --variant library> selects a side build."""
import os, statistics, subprocess, sys, threading, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from music_mixing_style_transfer_b200 import _cabi
if os.environ.get("MST_DEV_LIB"):
    _cabi.LIB_PATH = os.environ["MST_DEV_LIB"]
from gpu_helpers import models
from oracle import fixtures, weights as W
_, tcn = models()
tcn.precision = "f16f8"
REPS = int(sys.argv[1]) if len(sys.argv) > 1 else 30
BATCH = int(sys.argv[2]) if len(sys.argv) > 2 else 32
LEN = int(sys.argv[3]) if len(sys.argv) > 3 else 262144
x = W.synthetic_audio(BATCH, LEN, seed=3).cuda()
c = fixtures.make_cond(1, 4).cuda()
out = torch.empty_like(x)
samples = []
proc = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "50", "-i", "0"],
                        stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
threading.Thread(target=lambda: [samples.append((time.time(), l)) for l in proc.stdout], daemon=True).start()
with torch.no_grad():
    for _ in range(5):
        tcn(x, c, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time(); e0.record()
    for _ in range(REPS):
        tcn(x, c, out=out)
    e1.record(); torch.cuda.synchronize(); t1 = time.time()
time.sleep(0.1); proc.terminate()
clk, pw = [], []
for ts, l in samples:
    if t0 + 0.05 <= ts <= t1:
        f = l.split(",")
        try: clk.append(float(f[0])); pw.append(float(f[1]))
        except ValueError: pass
print("%s: B %d L %d: %.2f ms per TCN forward (%d reps) | SM %.0f MHz, %.0f W" % (os.environ.get("MST_DEV_LIB", "product") or "product", BATCH, LEN,
      e0.elapsed_time(e1) / REPS, REPS, statistics.median(clk) if clk else float("nan"), statistics.median(pw) if pw else float("nan")))
