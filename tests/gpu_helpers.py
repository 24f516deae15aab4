"""Shared helpers for the -m gpu parity tests: seeded models on cuda:0 + oracle on CPU + error metrics."""
import os

import numpy as np
import torch
import yaml

from oracle import networks_oracle as O, weights as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RMS_TOL = 1e-4     # BASELINE north_star: output waveforms within 1e-4 RMS of the reference CPU forward
_cache = {}


def state_dicts():
    if "sd" not in _cache:
        _cache["sd"] = (W.make_encoder_state_dict(0), W.make_tcn_state_dict(0))
    return _cache["sd"]


def models():
    """Our FXencoder / TCNModel on cuda:0 with the seeded reference-layout state_dicts loaded."""
    if "models" not in _cache:
        from music_mixing_style_transfer_b200.networks import FXencoder, TCNModel
        cfg = yaml.full_load(open(os.path.join(ROOT, "music_mixing_style_transfer_b200", "inference", "configs.yaml")))
        c = cfg["TCN"]["default"]
        enc = FXencoder(cfg["Effects_Encoder"]["default"])
        tcn = TCNModel(nparams=c["condition_dimension"], ninputs=2, noutputs=2, nblocks=c["nblocks"],
                       dilation_growth=c["dilation_growth"], kernel_size=c["kernel_size"],
                       channel_width=c["channel_width"], stack_size=c["stack_size"],
                       cond_dim=c["condition_dimension"], causal=c["causal"])
        esd, tsd = state_dicts()
        enc.load_state_dict(esd)
        tcn.load_state_dict(tsd)
        _cache["models"] = (enc.cuda().eval(), tcn.cuda().eval())
    return _cache["models"]


def err_stats(got, ref):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    d = got - ref
    ac = ref - ref.mean(axis=-1, keepdims=True)
    return {"rms": float(np.sqrt(np.mean(d ** 2))), "max": float(np.abs(d).max()),
            "ref_rms": float(np.sqrt(np.mean(ref ** 2))), "ref_ac_rms": float(np.sqrt(np.mean(ac ** 2))),
            "rel": float(np.sqrt(np.mean(d ** 2)) / max(1e-30, np.sqrt(np.mean(ref ** 2))))}


def oracle_threads():
    # torch's CPU convolutions are fastest at ~16 intra-op threads on the 128-core GPU box (bench.py calibrates the same)
    torch.set_num_threads(max(1, min(16, os.cpu_count() or 1)))
