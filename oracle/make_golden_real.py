"""TEST INFRASTRUCTURE ONLY -- real-audio golden vectors: the UNMODIFIED reference networks and FX chain run on windows
of the reference's own sample stems (samples/style_transfer/#0), decoded by the reference's own loader
(data_loader/loader_utils.py:47-70).

    python oracle/make_golden_real.py        (this container only; /root/reference does not exist on the GPU box)

Unlike the seeded fixtures of make_golden.py the INPUTS cannot be regenerated on the GPU box, so the int16 windows are
stored next to the reference outputs (tests/golden/real_audio.npz, ~3 MB):
  drums / vocals   65,536-frame windows of the input stems: transients; 64 % digital silence followed by an onset
  drums_fs         the drums window scaled to a full-scale peak of 1.0 (clip-free)
  mix_full         262,144 frames (the BASELINE segment length) of the input mixture
For every window: `emb_*` = reference FXencoder output, `y_*` = reference TCNModel output conditioned on `cond_*` = the
reference encoder's embedding of the SAME window of the style-reference song's stem (the real conditioning flow,
inference/style_transfer.py:144-162); the full-length case stores windows + strided samples of y like tcn_full.npz.
`fx_*` = reference AugmentationChain (eq, comp, imager, gain; EQ biquads restated, UNPINNED) on 16,000 frames of vocals.
Weights: the seeded state_dicts of oracle/weights.py (the public checkpoints are not part of the reference repo).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import fx_oracle, ref_import, weights  # noqa: E402
from oracle.fixtures import FULL_STRIDE, FULL_WINDOWS, GOLDEN_DIR as OUT  # noqa: E402
from oracle.make_golden import reference_fx_chain  # noqa: E402

SONG = os.path.join(ref_import.REFERENCE_ROOT, "samples", "style_transfer", "#0")
WIN = 65536
CASES = {"drums": ("drums", 2 * WIN, WIN), "vocals": ("vocals", 3 * WIN, WIN)}


def stem_path(which, inst):
    return os.path.join(SONG, "separated", "mdx_extra", which, f"{inst}.wav")


def main():
    lu = ref_import.import_reference_loader_utils()
    torch.set_num_threads(os.cpu_count())
    esd, tsd = weights.make_encoder_state_dict(0), weights.make_tcn_state_dict(0)
    enc, tcn = ref_import.build_reference_models(esd, tsd)

    def load(path, start, n):          # the reference's decode: float64 [2, n] in [-1, 1)
        return lu.load_wav_segment(path, start_point=start, duration=n, axis=0)

    def to_i16(x):                     # exact inverse of the reference's x / 2^15
        q = np.rint(x * 32768.0)
        assert np.array_equal(q / 32768.0, x)
        return q.astype(np.int16).T.copy()      # [n, 2] interleaved like the file

    out = {}
    with torch.no_grad():
        for name, (inst, start, n) in CASES.items():
            x = load(stem_path("input", inst), start, n)
            r = load(stem_path("reference", inst), start, n)
            xt, rt = torch.from_numpy(x).float()[None], torch.from_numpy(r).float()[None]
            cond = enc(rt)
            out[f"x_{name}"], out[f"ref_{name}"] = to_i16(x), to_i16(r)
            out[f"emb_{name}"], out[f"cond_{name}"] = enc(xt)[0].numpy(), cond[0].numpy()
            out[f"y_{name}"] = tcn(xt, cond)[0].numpy()
            if name == "drums":
                peak = float(np.abs(x).max())
                xf = torch.from_numpy(x / peak).float()[None]        # float64 divide, then the float32 cast the loader's caller does
                out["drums_fs_peak"] = np.float64(peak)
                out["emb_drums_fs"] = enc(xf)[0].numpy()
                out["y_drums_fs"] = tcn(xf, cond)[0].numpy()
        # BASELINE segment length on the mixture
        x = load(os.path.join(SONG, "input.wav"), 2 * WIN, 262144)
        r = load(os.path.join(SONG, "reference.wav"), 2 * WIN, 262144)
        xt, rt = torch.from_numpy(x).float()[None], torch.from_numpy(r).float()[None]
        cond = enc(rt)
        y = tcn(xt, cond)[0].numpy()
        out["x_mix_full"] = to_i16(x)
        out["emb_mix_full"], out["cond_mix_full"] = enc(xt)[0].numpy(), cond[0].numpy()
        out["y_mix_full_windows"] = np.stack([y[:, s:s + n] for s, n in FULL_WINDOWS])
        out["y_mix_full_strided"] = y[:, ::FULL_STRIDE]
        out["y_mix_full_ac_rms"] = np.float64(np.sqrt(np.mean((y - y.mean(-1, keepdims=True)) ** 2)))

    ca = ref_import.import_reference_fx()
    P = fx_oracle.random_params(2, seed=78)
    xv = load(stem_path("input", "vocals"), 4 * WIN, 16000).T.astype(np.float32)    # [16000, 2] like the chain's arrays
    xd = load(stem_path("input", "drums"), 2 * WIN + 16000, 16000).T.astype(np.float32)
    out["fx_params"] = P
    out["fx_x0"], out["fx_x1"] = to_i16(xv.T.astype(np.float64)), to_i16(xd.T.astype(np.float64))
    out["fx_y0"] = reference_fx_chain(ca, P[0])([xv.copy()])[0].astype(np.float32)
    out["fx_y1"] = reference_fx_chain(ca, P[1])([xd.copy()])[0].astype(np.float32)

    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, "real_audio.npz")
    np.savez_compressed(path, **out)
    for k, v in out.items():
        a = np.asarray(v)
        print(k, a.dtype, a.shape, "" if a.ndim == 0 else f"rms {np.sqrt(np.mean(a.astype(np.float64) ** 2)):.4g}")
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
