"""Time the 13 dilated TCN launches (config 2 shapes) with CUDA events; prints ms per launch, per block, and the SM clock /
board power nvidia-smi saw meanwhile (the kernel runs at the 1000 W power cap, so ms alone does not separate pipeline
efficiency from clock).  usage: tcn_time.py [f16f8|bf16x3] [B] [reps];  MST_DEV_LIB=<build.py --variant library> for A/B."""
import os, statistics, subprocess, sys, threading, time, torch
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
REPS = int(sys.argv[3]) if len(sys.argv) > 3 else 4
x = W.synthetic_audio(BATCH, 262144, seed=3).cuda()
c = fixtures.make_cond(1, 4).cuda()
ev = []
def note(name, phase):
    if name == "tcn_block_umma_kernel":
        e = torch.cuda.Event(enable_timing=True); e.record(); ev.append(e)
samples = []
proc = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "50", "-i", "0"],
                        stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
threading.Thread(target=lambda: [samples.append((time.time(), l)) for l in proc.stdout], daemon=True).start()
with torch.no_grad():
    tcn.forward_layers(x, c)
    tcn.forward_layers(x, c)
    torch.cuda.synchronize()
    ev.clear()
    t0 = time.time()
    for _ in range(REPS):
        tcn.forward_layers(x, c, note)
    torch.cuda.synchronize()
    t1 = time.time()
time.sleep(0.1); proc.terminate()
d = [ev[i].elapsed_time(ev[i + 1]) for i in range(0, len(ev), 2)]
clk, pw = [], []
for ts, l in samples:
    if t0 + 0.05 <= ts <= t1:
        f = l.split(",")
        try: clk.append(float(f[0])); pw.append(float(f[1]))
        except ValueError: pass
mhz = statistics.median(clk) if clk else float("nan")
print(tcn.precision, "B", BATCH, os.environ.get("MST_DEV_LIB", "product"), "ms/launch mean %.3f min %.3f max %.3f | SM %.0f MHz, %.0f W (median of %d samples) | Mcycles/launch %.2f"
      % (sum(d) / len(d), min(d), max(d), mhz, statistics.median(pw) if pw else float("nan"), len(clk), sum(d) / len(d) * mhz / 1e3))
n = 13
print("per block (dilation 2^n, n=1..13), ms:", " ".join("%.2f" % (sum(d[i::n]) / len(d[i::n])) for i in range(n)))
