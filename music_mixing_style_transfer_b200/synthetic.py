"""Seeded random weights in the reference's `state_dict` key layout + seeded synthetic audio (bench.py, smoke(), tests).

No compute lives here: these are the inputs BOTH arms (the CUDA path and the CPU oracle) are fed with.

The pretrained checkpoints (FXencoder_ps.pt / MixFXcloner_ps.pt) are NOT in the reference repo
(README.md:15-16, inference/style_transfer.py:340-341), so parity is established on seeded random weights loaded
through the same state_dict keys.  Every tensor is drawn from its own torch.Generator seeded by
crc32(key) ^ seed, so the values do not depend on module construction order and are identical here and on the
GPU box.  BatchNorm running stats / affine are randomised so eval-BN folding is really exercised, FiLM biases put
gamma near 1, and `output.weight` carries a fixed gain so the TCN output has AC-RMS ~0.1 (a 1e-4 absolute
tolerance would be vacuous on the ~3e-3 AC-RMS output of the default init).

Key layout (probed from the reference modules):
  encoder.{i}.conv{1,2}.conv1d.conv1d.{weight,bias}
  encoder.{i}.conv{1,2}.conv1d.batch_norm.{weight,bias,running_mean,running_var,num_batches_tracked}
  blocks.{n}.conv1.weight / blocks.{n}.film.film_fc.{weight,bias} / blocks.{n}.bn.* / blocks.{n}.res.weight
  output.{weight,bias}
"""
import math
import zlib
from collections import OrderedDict

import torch

ENC_CHANNELS = [2, 16, 32, 64, 128, 256, 256, 512, 512, 1024, 1024, 2048, 2048]  # inference/configs.yaml:8 (+2 in)
ENC_KERNELS = [25, 25, 15, 15, 10, 10, 10, 10, 5, 5, 5, 5]                      # configs.yaml:9
ENC_STRIDES = [4, 4, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1]                              # configs.yaml:10
TCN_NBLOCKS, TCN_K, TCN_CH, TCN_COND = 14, 15, 128, 2048                         # configs.yaml:19-30

# fixed gain on output.weight: makes the reference TCN output AC-RMS ~0.1 for N(0,0.1) input (see make_golden.py)
TCN_OUTPUT_GAIN = 0.6


def _gen(key: str, seed: int) -> torch.Generator:
    g = torch.Generator()
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def _uniform(key, seed, shape, lo, hi):
    return torch.rand(shape, generator=_gen(key, seed), dtype=torch.float32) * (hi - lo) + lo


def _bn(sd, prefix, ch, seed):
    sd[prefix + ".weight"] = _uniform(prefix + ".weight", seed, (ch,), 0.5, 1.5)
    sd[prefix + ".bias"] = _uniform(prefix + ".bias", seed, (ch,), -0.1, 0.1)
    sd[prefix + ".running_mean"] = _uniform(prefix + ".running_mean", seed, (ch,), -0.1, 0.1)
    sd[prefix + ".running_var"] = _uniform(prefix + ".running_var", seed, (ch,), 0.5, 1.5)
    sd[prefix + ".num_batches_tracked"] = torch.tensor(100, dtype=torch.long)


def make_encoder_state_dict(seed: int = 0, channels=None, kernels=None) -> "OrderedDict[str, torch.Tensor]":
    channels = ENC_CHANNELS if channels is None else channels
    kernels = ENC_KERNELS if kernels is None else kernels
    sd = OrderedDict()
    for i, k in enumerate(kernels):
        cin, cout = channels[i], channels[i + 1]
        for name, co in (("conv1", cin), ("conv2", cout)):
            p = f"encoder.{i}.{name}.conv1d"
            bound = 1.0 / math.sqrt(cin * k)
            # a little hotter than torch's default so the 24-layer ReLU stack keeps a healthy signal level
            sd[p + ".conv1d.weight"] = _uniform(p + ".conv1d.weight", seed, (co, cin, k), -1.7 * bound, 1.7 * bound)
            sd[p + ".conv1d.bias"] = _uniform(p + ".conv1d.bias", seed, (co,), -bound, bound)
            _bn(sd, p + ".batch_norm", co, seed)
    return sd


def make_tcn_state_dict(seed: int = 0, nblocks=TCN_NBLOCKS, ch=TCN_CH, k=TCN_K, cond=TCN_COND,
                        ninputs=2, noutputs=2, output_gain=None) -> "OrderedDict[str, torch.Tensor]":
    sd = OrderedDict()
    for n in range(nblocks):
        cin = ninputs if n == 0 else ch
        p = f"blocks.{n}"
        bound = 1.0 / math.sqrt(cin * k)
        sd[p + ".conv1.weight"] = _uniform(p + ".conv1.weight", seed, (ch, cin, k), -1.7 * bound, 1.7 * bound)
        fb = 1.0 / math.sqrt(cond)
        sd[p + ".film.film_fc.weight"] = _uniform(p + ".film.film_fc.weight", seed, (2 * ch, cond), -fb, fb)
        fbias = torch.empty(2 * ch)
        fbias[:ch] = _uniform(p + ".film.film_fc.bias.g", seed, (ch,), 0.7, 1.3)     # gamma ~ 1
        fbias[ch:] = _uniform(p + ".film.film_fc.bias.b", seed, (ch,), -0.1, 0.1)    # beta ~ 0
        sd[p + ".film.film_fc.bias"] = fbias
        _bn(sd, p + ".bn", ch, seed)
        sd[p + ".res.weight"] = _uniform(p + ".res.weight", seed, (ch, 1, 1), -1.0, 1.0)
    g = TCN_OUTPUT_GAIN if output_gain is None else output_gain
    ob = 1.0 / math.sqrt(ch)
    sd["output.weight"] = _uniform("output.weight", seed, (noutputs, ch, 1), -ob, ob) * g
    sd["output.bias"] = _uniform("output.bias", seed, (noutputs,), -0.01, 0.01)
    return sd


def synthetic_audio(batch: int, length: int, seed: int = 1234, std: float = 0.1) -> torch.Tensor:
    """Seeded N(0, std^2) stereo waveforms clamped to [-1, 1] (the reference clamps inputs,
    data_loader.py:589-590); SURVEY.md 8d 'Synthetic values'."""
    g = torch.Generator()
    g.manual_seed(seed)
    return (torch.randn(batch, 2, length, generator=g, dtype=torch.float32) * std).clamp_(-1.0, 1.0)
