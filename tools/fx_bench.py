"""Config 3 of BASELINE.json: FX chain only, batch=256 random-parameter segments of 262144 stereo samples.
Prints one JSON line with the whole-chain time, per-kernel times (CUDA events via three single-purpose calls are not
possible -- the chain is one C-ABI call -- so per-kernel numbers come from the ncu launch list) and the HBM roofline
fraction (algorithmic bytes = read once + write once = 16 B per stereo frame)."""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from music_mixing_style_transfer_b200.mixing_manipulator import fx_chain_forward
from oracle import fx_oracle

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
L = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
g = torch.Generator(device="cuda"); g.manual_seed(1234)
x = (torch.randn(B, 2, L, generator=g, device="cuda") * 0.1).clamp_(-1, 1)
P = torch.from_numpy(fx_oracle.random_params(B, seed=1234)).cuda()
y = torch.empty_like(x)
for _ in range(3):
    fx_chain_forward(x, P, out=y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    fx_chain_forward(x, P, out=y)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
bytes_alg = 16.0 * B * L
peak = 6570.9
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
gbs = bytes_alg / (ms * 1e-3) / 1e9
print(json.dumps({"workload": f"configs[2]: FX chain B={B} L={L}", "ms": ms, "audio_s_per_s": B * L / 44100 / (ms * 1e-3),
                  "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                               "algorithmic_bytes": bytes_alg}}))
from music_mixing_style_transfer_b200.mixing_manipulator import common_audioeffects as _ca
for ws in _ca._ws_cache.values():
    st = ws[:B * 16 * 8].view(torch.float64).reshape(B, 16).cpu().numpy()
    tiles = (L + 4095) // 4096
    print("compressor smoother rounds per tile: mean %.2f max %.2f" % (st[:, 9].mean() / tiles, st[:, 9].max() / tiles))
# parity spot check on 2 segments at full length against the CPU oracle
for i in (0, B - 1):
    ref = fx_oracle.fx_chain(np.ascontiguousarray(x[i].cpu().numpy().T), P[i].cpu().numpy()).T
    d = y[i].cpu().numpy().astype(np.float64) - ref
    print(f"segment {i}: rms err {np.sqrt((d**2).mean()):.3e}  ref rms {np.sqrt((ref.astype(np.float64)**2).mean()):.3e}")
