"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz by running the UNMODIFIED reference from /root/reference.

    python oracle/make_golden.py          (this container only; /root/reference does not exist on the GPU box)

Inputs and weights are regenerated from seeds (oracle/weights.py), so only the reference OUTPUTS are stored:
  enc_small.npz   FXencoder(x[2,2,32768])                      -> emb [2,2048]
  tcn_small.npz   TCNModel(x[2,2,8191], cond[1,2048])          -> y [2,2,8191]       (odd length, all taps clipped)
  tcn_percond.npz TCNModel(x[3,2,4099], cond[3,2048])          -> y                   (per-segment conditioning)
  tcn_full.npz    TCNModel(x[1,2,262144], cond[1,2048])        -> windows + strided samples of y (BASELINE length)
  tcn_blocks.npz  TCNBlock n=0,1,4,9,13 on x[1,C,3000]         -> y[0, ::16, :] (8 of the 128 channels) each
  fx_chain.npz    reference AugmentationChain(eq,comp,imager,gain) on 3 x [16000,2], params from fx_oracle.random_params
                  (compressor / imager / gain / chain logic = reference code; EQ biquads = restated IIRfilter, UNPINNED)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import fx_oracle, ref_import, weights  # noqa: E402

from oracle.fixtures import (BLOCK_CH_STRIDE, BLOCK_LEN, FULL_STRIDE, FULL_WINDOWS, GOLDEN_DIR as OUT,  # noqa: E402
                             fx_input, make_cond)


def reference_fx_chain(ca, p):
    eq = ca.Equaliser(n_channels=2, sample_rate=44100)
    names = ['low_shelf_gain', 'low_shelf_freq', 'first_band_gain', 'first_band_freq', 'first_band_q',
             'second_band_gain', 'second_band_freq', 'second_band_q', 'third_band_gain', 'third_band_freq',
             'third_band_q', 'high_shelf_gain', 'high_shelf_freq']
    for j, n in enumerate(names):
        getattr(eq.parameters, n).value = float(p[j])
    comp = ca.Compressor(sample_rate=44100)
    for j, n in enumerate(['threshold', 'attack_time', 'release_time', 'ratio']):
        getattr(comp.parameters, n).value = float(p[13 + j])
    im = ca.MidSideImager()
    im.parameters.bal.value = float(p[17])
    g = ca.Gain()
    g.parameters.gain.value = float(p[18])
    g.parameters.invert.value = bool(p[19] >= 0.5)
    return ca.AugmentationChain([(eq, 1, True), (comp, 1, True), (im, 1, True), (g, 1, False)],
                                randomize_param_value=False)


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    esd, tsd = weights.make_encoder_state_dict(0), weights.make_tcn_state_dict(0)
    enc, tcn = ref_import.build_reference_models(esd, tsd)
    with torch.no_grad():
        x = weights.synthetic_audio(2, 32768, seed=11)
        np.savez_compressed(os.path.join(OUT, "enc_small.npz"), emb=enc(x).numpy())

        x = weights.synthetic_audio(2, 8191, seed=12)
        np.savez_compressed(os.path.join(OUT, "tcn_small.npz"), y=tcn(x, make_cond(1, 21)).numpy())

        x = weights.synthetic_audio(3, 4099, seed=13)
        np.savez_compressed(os.path.join(OUT, "tcn_percond.npz"), y=tcn(x, make_cond(3, 22)).numpy())

        x = weights.synthetic_audio(1, 262144, seed=14)
        y = tcn(x, make_cond(1, 23))[0].numpy()
        np.savez_compressed(os.path.join(OUT, "tcn_full.npz"),
                            windows=np.stack([y[:, s:s + n] for s, n in FULL_WINDOWS]),
                            strided=y[:, ::FULL_STRIDE],
                            ac_rms=np.float64(np.sqrt(np.mean((y - y.mean(-1, keepdims=True)) ** 2))))

        blocks = {}
        cond = make_cond(1, 24)
        for n in (0, 1, 4, 9, 13):
            cin = 2 if n == 0 else 128
            g = torch.Generator()
            g.manual_seed(300 + n)
            xb = torch.randn(1, cin, BLOCK_LEN, generator=g) * 0.5
            blocks[f"b{n}"] = tcn.blocks[n](xb, cond)[0, ::BLOCK_CH_STRIDE].numpy()
        np.savez_compressed(os.path.join(OUT, "tcn_blocks.npz"), **blocks)

    ca = ref_import.import_reference_fx()
    P = fx_oracle.random_params(3, seed=77)
    outs = {}
    for i in range(3):
        outs[f"y{i}"] = reference_fx_chain(ca, P[i])([fx_input(i, 16000)])[0].astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "fx_chain.npz"), params=P, **outs)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
