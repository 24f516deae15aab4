"""BASELINE.json sizes (config 2: 32 x 262144, config 3: 256 x 262144) through size-independent properties -- no CPU
oracle at these sizes: determinism, segment independence, bounds, RMS conservation of the re-normalised FX stages,
linearity of the gain stage, batch-permutation equivariance."""
import numpy as np
import pytest
import torch

from gpu_helpers import models
from oracle import fixtures, fx_oracle, weights as W

pytestmark = pytest.mark.gpu
L = 262144


def test_tcn_config2_properties():
    enc, tcn = models()
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    x = (torch.randn(32, 2, L, generator=g, device="cuda") * 0.1).clamp_(-1, 1)
    with torch.no_grad():
        emb = enc(x).mean(dim=0)
        y = tcn(x, emb.unsqueeze(0))
        y2 = tcn(x, emb.unsqueeze(0))
        assert torch.equal(y, y2), "not bit-reproducible"
        assert y.shape == (32, 2, L) and bool(torch.isfinite(y).all()) and float(y.abs().max()) <= 1.0
        # segments are independent: a slice of the batch gives the same bits (zero padding per segment, no cross-batch op)
        for i in (0, 17, 31):
            assert torch.equal(y[i:i + 1], tcn(x[i:i + 1].contiguous(), emb.unsqueeze(0))), i
        # per-segment conditioning with identical rows == broadcast conditioning
        y3 = tcn(x[:4].contiguous(), emb.unsqueeze(0).repeat(4, 1))
        assert torch.equal(y3, y[:4])
        # the forward_layers path (one C-ABI call per launch, used by bench.py's roofline leg) is the same computation
        assert torch.equal(tcn.forward_layers(x[:2].contiguous(), emb.unsqueeze(0)), y[:2])
        # encoder: batch-permutation equivariance at full length
        perm = torch.randperm(32, device="cuda")
        e1, e2 = enc(x), enc(x[perm].contiguous())
        assert torch.equal(e1[perm], e2)
    # one segment of the batch against the committed full-length golden vector is in test_gpu_tcn.py


def test_fx_config3_properties():
    from music_mixing_style_transfer_b200.mixing_manipulator import (FX_COMP, FX_EQ, FX_GAIN, FX_IMAGER, FX_RMSNORM,
                                                                     fx_chain_forward)
    B = 256
    g = torch.Generator(device="cuda")
    g.manual_seed(6)
    x = (torch.randn(B, 2, L, generator=g, device="cuda") * 0.1).clamp_(-1, 1)
    x[:, 1] = 0.6 * x[:, 0] + 0.4 * x[:, 1]
    P = torch.from_numpy(fx_oracle.random_params(B, seed=1234)).cuda()
    rms = lambda t: t.double().pow(2).mean(dim=(1, 2)).sqrt()  # noqa: E731
    # EQ / compressor / imager each re-normalise to their input RMS (common_audioeffects.py:142-145)
    y = fx_chain_forward(x, P, FX_EQ | FX_COMP | FX_IMAGER | FX_RMSNORM)
    assert bool(torch.isfinite(y).all())
    assert float((rms(y) / rms(x) - 1).abs().max()) <= 2e-5
    # gain stage: exact scalar multiple 10^(g/20), sign flipped when `invert`
    yg = fx_chain_forward(x, P, FX_GAIN)
    gain = torch.pow(torch.tensor(10.0, device="cuda", dtype=torch.float64), P[:, 18].double() / 20.0)
    gain = torch.where(P[:, 19] >= 0.5, -gain, gain).float()
    assert torch.allclose(yg, x * gain[:, None, None], rtol=2e-6, atol=0)
    # full chain = the three normalised stages followed by the gain
    # (fx2.cu folds RMS factor, imager and gain into one 2x2 matrix per segment: out_L = m0 l + m1 r can cancel, so the
    # bound is float32 rounding relative to the segment's peak, not to each sample)
    yall = fx_chain_forward(x, P)
    want = y * gain[:, None, None]
    peak = want.abs().amax(dim=(1, 2), keepdim=True)
    assert float(((yall - want).abs() / peak).max()) <= 5e-7
    # batch-permutation equivariance and determinism
    perm = torch.randperm(B, device="cuda")
    assert torch.equal(fx_chain_forward(x[perm].contiguous(), P[perm].contiguous()), yall[perm])
    # first and last segment against the CPU oracle at full length
    for i in (0, B - 1):
        ref = fx_oracle.fx_chain(np.ascontiguousarray(x[i].cpu().numpy().T), P[i].cpu().numpy()).T
        err = float(np.sqrt(np.mean((yall[i].cpu().numpy().astype(np.float64) - ref) ** 2)))
        assert err <= 1e-5, (i, err)
