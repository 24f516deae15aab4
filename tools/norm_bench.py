"""SURVEY 8f-2 / 8f-4 kernels on one 3-minute stereo stem (7,938,000 frames + the normaliser's 2 x 65,536 padding): time of each
device piece (CUDA events, 5 repetitions after a warm-up) and of the whole input FX normaliser for the effects that need no
host-side detector, plus both reverbs on a 262,144-frame segment batch.  Prints one JSON line."""
import json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from music_mixing_style_transfer_b200 import _cabi
from music_mixing_style_transfer_b200.mixing_manipulator import data_normalization as dn, AlgorithmicReverb, ConvolutionalReverb

def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

T = 180 * 44100 + 2 * 65536
g = torch.Generator(device="cuda"); g.manual_seed(7)
x = (torch.randn(2, T, generator=g, device="cuda") * 0.1).clamp_(-1, 1)
out = {"stem_frames": T}
out["stft_mag_mean_ms"] = timed(lambda: dn.stft_mag_mean(x))
taps = np.random.RandomState(0).randn(2, 1001) / 30.0
out["fir_filtfilt_ms"] = timed(lambda: dn.fir_filtfilt(x, taps))
out["fir_filtfilt_gdfma_per_s"] = 2 * 2 * 1001 * (T + 3003) / (out["fir_filtfilt_ms"] * 1e-3) / 1e9
out["kweighted_block_sums_ms"] = timed(lambda: dn.kweighted_block_sums(x))
f = np.arange(32769) / 32768.0
spec = (30.0 / (1.0 + 200.0 * f) + 0.02).astype(np.float32)
feats = {"eq": {"drums": spec}, "loudness": {"drums": np.array([-28.9])}, "imager": {"drums": np.float32(0.94)}}
norm = dn.Audio_Effects_Normalizer(feats, STEMS=["drums"], EFFECTS=["loudness", "eq", "imager", "loudness"])
stem = x[:, 65536:-65536].contiguous()
t0 = time.time(); norm.normalize_audio(stem, src="drums"); torch.cuda.synchronize()
t0 = time.time(); norm.normalize_audio(stem, src="drums"); torch.cuda.synchronize()
out["normalizer_loudness_eq_imager_loudness_wall_s"] = time.time() - t0
out["normalizer_audio_s_per_s"] = 180.0 / out["normalizer_loudness_eq_imager_loudness_wall_s"]
# reverbs through the C ABI on a batch of segments
lib = _cabi.lib()
B, L = 64, 262144
xb = (torch.randn(B, 2, L, generator=g, device="cuda") * 0.1)
p = torch.tensor([[0.6, 0.3, 0.8, 0.35, 0.6]] * B, dtype=torch.float32, device="cuda")
ws = torch.empty(lib.mst_algo_reverb_workspace_bytes(B, L), dtype=torch.uint8, device="cuda")
yb = torch.empty_like(xb)
out["algo_reverb_B64_ms"] = timed(lambda: _cabi.check(lib.mst_algo_reverb(_cabi.ptr(xb), _cabi.ptr(p), _cabi.ptr(yb), B, L, _cabi.ptr(ws), ws.numel(), _cabi.current_stream()), "rvb"))
out["algo_reverb_audio_s_per_s"] = B * L / 44100 / (out["algo_reverb_B64_ms"] * 1e-3)
M = 2 * 44100
h = (torch.randn(2, M, generator=g, device="cuda") * torch.exp(-torch.arange(M, device="cuda") / 15000.0)).contiguous()
ws2 = torch.empty(lib.mst_fft_convolve_workspace_bytes(L, M), dtype=torch.uint8, device="cuda")
y1 = torch.empty(2, L, device="cuda")
out["fft_convolve_262144x88200_ms"] = timed(lambda: _cabi.check(lib.mst_fft_convolve(_cabi.ptr(xb[0]), L, L, _cabi.ptr(h), M, M, 2, 100, 0.3, 0.7, _cabi.ptr(y1), L, _cabi.ptr(ws2), ws2.numel(), _cabi.current_stream()), "conv"))
print(json.dumps(out))
