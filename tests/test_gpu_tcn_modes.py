"""TCN operand formats (include/mst_b200.h: MST_TCN_F16F8 / MST_TCN_BF16X3) and the f16f8 operand-range guard.

The precision is a per-call argument of the C ABI (`TCNModel.precision` on the module), so both formats are exercised in
one process.  The range tests scale FiLM of block 0 so that the inter-block activations reach 1e2 / 1e3 / 1e4 -- the
reference's FiLM gamma is unbounded (networks/network_utils.py:180-182) -- and undo the scale in `output.weight`, so the
north_star tolerance (1e-4 RMS on the output waveform) stays meaningful: either f16f8 holds, or the guard fires and
the forward is repeated in bf16x3."""
import numpy as np
import pytest
import torch

from gpu_helpers import RMS_TOL, err_stats, models, oracle_threads, state_dicts
from oracle import fixtures, networks_oracle as O, weights as W

pytestmark = pytest.mark.gpu


@pytest.fixture()
def bf16x3_tcn():
    _, tcn = models()
    tcn.precision = "bf16x3"
    yield tcn
    tcn.precision = "auto"


@pytest.mark.parametrize("n", list(range(1, 14)))
def test_bf16x3_every_block_matches_oracle(bf16x3_tcn, n):
    oracle_threads()
    _, tsd = state_dicts()
    g = torch.Generator()
    g.manual_seed(1700 + n)
    x = torch.randn(2, 128, 4099, generator=g) * 0.5
    cond = fixtures.make_cond(1, 1701 + n)
    with torch.no_grad():
        ref = O.tcn_block(x, cond, tsd, f"blocks.{n}", 15, 2 ** n)
        got = bf16x3_tcn.blocks[n](x.cuda(), cond.cuda()).cpu()
    e = err_stats(got, ref)
    assert e["rms"] <= 3e-5 * max(1.0, e["ref_rms"]) and e["max"] <= 5e-4 * max(1.0, e["ref_rms"]), (n, e)


@pytest.mark.parametrize("n", [1, 3, 6])
def test_bf16x3_interleaved_pairing(bf16x3_tcn, n):
    oracle_threads()
    _, tsd = state_dicts()
    g = torch.Generator()
    g.manual_seed(1800 + n)
    x = torch.randn(2, 128, 4224, generator=g) * 0.5
    cond = fixtures.make_cond(1, 1801 + n)
    with torch.no_grad():
        ref = O.tcn_block(x, cond, tsd, f"blocks.{n}", 15, 2 ** n)
        got = bf16x3_tcn.blocks[n](x.cuda(), cond.cuda()).cpu()
    e = err_stats(got, ref)
    assert e["rms"] <= 3e-5 * max(1.0, e["ref_rms"]) and e["max"] <= 5e-4 * max(1.0, e["ref_rms"]), (n, e)


@pytest.mark.parametrize("name,B,L,seed,cseed,ncond", [("tcn_small.npz", 2, 8191, 12, 21, 1),
                                                        ("tcn_percond.npz", 3, 4099, 13, 22, 3)])
def test_bf16x3_golden_full_tcn(bf16x3_tcn, name, B, L, seed, cseed, ncond):
    x = W.synthetic_audio(B, L, seed=seed)
    with torch.no_grad():
        got = bf16x3_tcn(x.cuda(), fixtures.make_cond(ncond, cseed).cuda()).cpu().numpy()
    e = err_stats(got, fixtures.load_golden(name)["y"])
    assert e["rms"] <= 2e-5, e


def _scaled_model(scale):
    from music_mixing_style_transfer_b200.networks import TCNModel
    tsd = W.make_tcn_state_dict(5)
    tsd["blocks.0.film.film_fc.weight"] = tsd["blocks.0.film.film_fc.weight"] * scale
    tsd["blocks.0.film.film_fc.bias"] = tsd["blocks.0.film.film_fc.bias"] * scale
    tsd["output.weight"] = tsd["output.weight"] / scale
    m = TCNModel(nparams=2048, ninputs=2, noutputs=2, nblocks=14, dilation_growth=2, kernel_size=15, channel_width=128,
                 stack_size=15, cond_dim=2048, causal=False)
    m.load_state_dict(tsd)
    return m.cuda().eval(), tsd


@pytest.mark.parametrize("scale", [1e2, 1e3, 1e4])
def test_activation_range_guard(scale):
    """auto precision: parity at the north_star tolerance whatever the activation scale; the guard must fire (and the
    bf16x3 repeat run) once activations pass the e4m3 range of the f16f8 format (448)."""
    oracle_threads()
    m, tsd = _scaled_model(scale)
    x = W.synthetic_audio(2, 6000, seed=41)
    cond = fixtures.make_cond(1, 42)
    with torch.no_grad():
        ref = O.tcn_forward(x, cond, tsd)
        got = m(x.cuda(), cond.cuda()).cpu()
    e = err_stats(got, ref)
    assert e["ref_ac_rms"] > 0.01, e            # the fixture carries a real signal
    assert e["rms"] <= RMS_TOL, (scale, m.last_range_excess, e)
    if scale >= 1e3:
        assert m.last_range_excess > 448.0, (scale, m.last_range_excess)
    # deferred form: explicit f16f8 + a caller-owned flag, no host read-back inside forward
    m.precision = "f16f8"
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    with torch.no_grad():
        y = m(x.cuda(), cond.cuda(), range_flag=flag)
        excess = float(flag.view(torch.float32).item())
        assert (excess > 448.0) == (m.last_range_excess > 448.0), (excess, m.last_range_excess)
        if excess > 0:
            y = m.rerun_bf16x3(x.cuda(), cond.cuda(), y)
    assert err_stats(y.cpu(), ref)["rms"] <= RMS_TOL


def test_in_range_forward_does_not_fire():
    _, tcn = models()
    x = W.synthetic_audio(2, 5000, seed=43)
    with torch.no_grad():
        tcn(x.cuda(), fixtures.make_cond(1, 44).cuda())
    assert tcn.last_range_excess == 0.0
