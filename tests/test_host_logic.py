"""Host-side logic: module surfaces / state_dict layout, segmentation, shard partition, 2-rank gloo exchange (CPU)."""
import copy
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp
import yaml

from music_mixing_style_transfer_b200 import shard
from music_mixing_style_transfer_b200.networks import FXencoder, TCNModel, TCNBlock
from oracle import weights as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_models():
    cfg = yaml.full_load(open(os.path.join(ROOT, "music_mixing_style_transfer_b200", "inference", "configs.yaml")))
    enc_cfg = cfg["Effects_Encoder"]["default"]
    before = copy.deepcopy(enc_cfg)
    enc = FXencoder(enc_cfg)
    assert enc_cfg == before, "constructor must not mutate the caller's config"
    FXencoder(enc_cfg)  # second construction from the same dict works (the reference's would not)
    c = cfg["TCN"]["default"]
    tcn = TCNModel(nparams=c["condition_dimension"], ninputs=2, noutputs=2, nblocks=c["nblocks"],
                   dilation_growth=c["dilation_growth"], kernel_size=c["kernel_size"], channel_width=c["channel_width"],
                   stack_size=c["stack_size"], cond_dim=c["condition_dimension"], causal=c["causal"])
    return enc, tcn


def test_state_dict_layout_matches_reference_keys():
    enc, tcn = build_models()
    esd, tsd = W.make_encoder_state_dict(0), W.make_tcn_state_dict(0)
    assert list(enc.state_dict().keys()) == list(esd.keys())
    assert list(tcn.state_dict().keys()) == list(tsd.keys())
    for k, v in enc.state_dict().items():
        assert v.shape == esd[k].shape, k
    for k, v in tcn.state_dict().items():
        assert v.shape == tsd[k].shape, k
    enc.load_state_dict(esd)
    tcn.load_state_dict(tsd)
    # DDP checkpoints carry a 7-char `module.` prefix that reload_weights strips (inference/style_transfer.py:102)
    ddp = {"module." + k: v for k, v in tsd.items()}
    tcn.load_state_dict({k[7:]: v for k, v in ddp.items()})
    assert tcn.compute_receptive_field() == 229363
    assert tcn.hparams.kernel_size == 15 and tcn.hparams["nblocks"] == 14
    assert isinstance(tcn.blocks[3], TCNBlock) and tcn.blocks[3].dilation == 8 and tcn.blocks[3].pad_length == 56


def test_no_cpu_fallback():
    enc, tcn = build_models()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        enc(torch.zeros(1, 2, 4096))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tcn(torch.zeros(1, 2, 4096), torch.zeros(1, 2048))
    from music_mixing_style_transfer_b200.mixing_manipulator import fx_chain_forward
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fx_chain_forward(torch.zeros(1, 2, 64), torch.zeros(1, 20))


def test_unsupported_variants_raise():
    with pytest.raises(NotImplementedError):
        TCNModel(nparams=2048, causal=True)
    with pytest.raises(NotImplementedError):
        TCNModel(nparams=2048, grouped=True)
    from music_mixing_style_transfer_b200.mixing_manipulator import create_effects_augmentation_chain
    with pytest.raises(NotImplementedError):
        create_effects_augmentation_chain(["eq", "expander"])     # the reference's factory dies with NameError here
    chain = create_effects_augmentation_chain(["eq", ("reverb", 0.3), "algorithmic"])
    assert [(f.name, p, n) for f, p, n in chain.fxs] == [("Equaliser", 1, True), ("algoreverb", 0.3, True), ("algoreverb", 1, True)]
    with pytest.raises(ValueError):
        create_effects_augmentation_chain(["wobble"])
    chain = create_effects_augmentation_chain(["eq", ("comp", 0.5), "imager", "gain"])
    assert [(f.name, p, n) for f, p, n in chain.fxs] == [("Equaliser", 1, True), ("Compressor", 0.5, True),
                                                          ("IMAGER", 1, True), ("Gain", 1, False)]
    assert chain.param_tensor(3).shape == (3, 20)


def test_entry_point_segmentation_matches_oracle():
    from music_mixing_style_transfer_b200.inference.style_transfer import Mixing_Style_Transfer_Inference, build_parser
    from oracle import networks_oracle as O
    args = build_parser().parse_args(["--segment_length", "250", "--batch_size", "3"])
    assert args.segment_length_ref == 2 ** 19 and args.interpolate_segments == 30 and args.normalize_input is True
    obj = Mixing_Style_Transfer_Inference.__new__(Mixing_Style_Transfer_Inference)
    obj.args = args
    song = torch.randn(2, 1000)
    ours = obj.batchwise_segmentization(song, "s", segment_length=250)
    ref = O.batchwise_segmentization(song, 250, 3)
    assert len(ours) == len(ref) and all(torch.equal(a, b) for a, b in zip(ours, ref))
    with pytest.raises(AssertionError):
        obj.batchwise_segmentization(torch.randn(2, 100), "s", segment_length=250)


def test_song_row_table_equals_reference_segmentation():
    """The engine cuts ALL stems of a song into one row table (inference/style_transfer.py: plan_cut / cut_rows / join_rows):
    the rows must be the reference's segments (oracle restatement of batchwise_segmentization, incl. the extra all-zero
    segment of an exact multiple), in stem-major order, and join_rows must undo the cut."""
    from music_mixing_style_transfer_b200.inference.style_transfer import (Mixing_Style_Transfer_Inference, build_parser,
                                                                           cut_rows, join_rows)
    from oracle import networks_oracle as O
    obj = Mixing_Style_Transfer_Inference.__new__(Mixing_Style_Transfer_Inference)
    obj.args = build_parser().parse_args(["--segment_length", "250", "--segment_length_ref", "300"])
    for T in (1000, 1001, 777, 251):
        stems = torch.randn(4, 2, T)
        seg, n = obj.plan_cut(T, 250, 250)
        rows = cut_rows(stems, seg, n)
        ref = torch.cat([torch.cat(O.batchwise_segmentization(stems[i], 250, 3), 0) for i in range(4)], 0)
        assert (seg, n) == (250, T // 250 + 1) and torch.equal(rows, ref), T
        assert torch.equal(join_rows(rows, 4, T), stems)
    # short stems pass through as one odd-length segment (style_transfer.py:131-132); the style reference is only cut above
    # TWICE the segment length (:133) and then at segment_length_ref (:136)
    assert obj.plan_cut(250, 250, 250) == (250, 1) and obj.plan_cut(199, 250, 250) == (199, 1)
    assert obj.plan_cut(500, 300, 500) == (500, 1) and obj.plan_cut(501, 300, 500) == (300, 2)
    assert torch.equal(cut_rows(torch.ones(3, 2, 199), 199, 1), torch.ones(3, 2, 199))
    with pytest.raises(AssertionError):     # cut requested but shorter than args.segment_length (:275-279)
        obj.plan_cut(200, 100, -1)


def test_stem_directory_layout_follows_do_not_separate():
    """data_loader/data_loader.py:555-556: --do_not_separate True drops the separation-model path component."""
    from music_mixing_style_transfer_b200.inference.style_transfer import Song_Dataset_Inference, build_parser
    a = build_parser().parse_args(["--target_dir", "/data/", "--do_not_separate", "True"])
    assert Song_Dataset_Inference(a).stem_level_directory_name == "separated"
    a = build_parser().parse_args(["--target_dir", "/data/"])
    assert Song_Dataset_Inference(a).stem_level_directory_name == os.path.join("separated", "mdx_extra")


def test_shard_bounds_cover_everything_once():
    for n in (0, 1, 7, 32, 33, 512):
        for ws in (1, 2, 3, 4, 8):
            spans = [shard.shard_bounds(n, ws, r) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(ws - 1))
            counts = shard.shard_counts(n, ws)
            assert sum(counts) == n and max(counts) - min(counts) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, n_total, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        torch.manual_seed(0)
        full = torch.randn(n_total, 2, 37)                       # every rank can rebuild the full batch
        ref = torch.randn(5, 2, 64)
        enc = lambda x: x.mean(dim=-1).repeat(1, 8)             # noqa: E731  stand-in encoder -> [B, 16]
        conv = lambda x, c: x * 2.0 + c[0, :1]                   # noqa: E731  stand-in converter
        lo, hi = shard.shard_bounds(n_total, ws, rank)
        emb, out = shard.sharded_style_transfer(enc, conv, ref if rank == 0 else None, full[lo:hi], n_total, cond_dim=16)
        expect_emb = enc(ref).mean(dim=0)
        ok = torch.allclose(emb, expect_emb) and torch.equal(out, full * 2.0 + expect_emb[:1])
        # sharded reference batch: every rank encodes its slice, one all-reduce of the partial sums
        rlo, rhi = shard.shard_bounds(ref.shape[0], ws, rank)
        for batch in (ref, ref[rlo:rhi]):                        # the whole batch, or just this rank's slice of it
            emb2, out2 = shard.sharded_style_transfer(enc, conv, batch, full[lo:hi], n_total, cond_dim=16,
                                                      shard_reference=True, n_reference=ref.shape[0], gather_chunks=2)
            ok = ok and torch.allclose(emb2, expect_emb, atol=1e-6) and torch.allclose(out2, full * 2.0 + expect_emb[:1], atol=1e-6)
        q.put((rank, bool(ok), tuple(out.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [8, 7, 1])
def test_two_rank_gloo_broadcast_and_allgather(n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert [r[1] for r in res] == [True, True], res
    assert all(r[2] == (n_total, 2, 37) for r in res)


def _worker_interp(rank, ws, port, n_total, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        torch.manual_seed(1)
        full = torch.randn(n_total, 2, 29)
        ref_a, ref_b = torch.randn(3, 2, 40), torch.randn(4, 2, 40)
        enc = lambda x: x.mean(dim=-1).repeat(1, 8)              # noqa: E731  stand-in encoder -> [B, 16]
        conv = lambda x, c: x + c[:, :1, None]                    # noqa: E731  stand-in converter, per-row conditioning
        S = 4
        w = shard.interpolation_weights(n_total, S)
        lo, hi = shard.shard_bounds(n_total, ws, rank)
        embs, out = shard.sharded_interpolation(enc, conv, ref_a if rank == 0 else None, ref_b if rank == 0 else None,
                                                full[lo:hi], n_total, w, cond_dim=16)
        ea, eb = enc(ref_a).mean(0), enc(ref_b).mean(0)
        cond = w[:, None] * ea[None] + (1 - w[:, None]) * eb[None]
        ok = torch.allclose(embs, torch.stack([ea, eb])) and torch.allclose(out, full + cond[:, :1, None])
        ok = ok and torch.allclose(w[:5], torch.tensor([1.0, 2 / 3, 1 / 3, 0.0, 1.0]))
        q.put((rank, bool(ok), tuple(out.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [8, 5])
def test_two_rank_gloo_interpolation(n_total):
    """BASELINE config 5 host logic: one broadcast of both reference embeddings, per-segment conditioning rows built per
    shard, all-gather of the outputs (world size 2 over gloo)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_interp, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert [r[1] for r in res] == [True, True], res
    assert all(r[2] == (n_total, 2, 29) for r in res)


def test_wav_header_checks_match_reference_loader(tmp_path):
    """wav_io.read_wav_pcm keeps load_wav_segment's checks (loader_utils.py:52-53, 62-63): ValueError for a wrong sample rate
    and for bit depths other than 16 / 32, raw interleaved PCM otherwise (no conversion on the host)."""
    import wave
    import numpy as np
    from music_mixing_style_transfer_b200 import wav_io
    x = (np.arange(2000, dtype=np.int64) * 37 % 65536 - 32768).astype("<i2").reshape(-1, 2)
    p16 = str(tmp_path / "a16.wav")
    with wave.open(p16, "wb") as w:
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(44100); w.writeframes(x.tobytes())
    got = wav_io.read_wav_pcm(p16)
    assert got.dtype == np.dtype("<i2") and got.shape == (1000, 2) and np.array_equal(got, x)
    assert np.array_equal(wav_io.read_wav_pcm(p16, start_point=10, duration=5), x[10:15])
    with pytest.raises(ValueError, match="sample rate"):
        wav_io.read_wav_pcm(p16, sample_rate=48000)
    p8 = str(tmp_path / "a8.wav")
    with wave.open(p8, "wb") as w:
        w.setnchannels(1); w.setsampwidth(1); w.setframerate(44100); w.writeframes(bytes(range(200)))
    with pytest.raises(ValueError, match="bit depth"):
        wav_io.read_wav_pcm(p8)
    p32 = str(tmp_path / "a32.wav")
    y = (np.arange(300, dtype=np.int64) * 7919 % (2 ** 32) - 2 ** 31).astype("<i4").reshape(-1, 1)
    with wave.open(p32, "wb") as w:
        w.setnchannels(1); w.setsampwidth(4); w.setframerate(44100); w.writeframes(y.tobytes())
    got = wav_io.read_wav_pcm(p32)
    assert got.dtype == np.dtype("<i4") and np.array_equal(got, y)
    # no CPU path for the conversion itself
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="CUDA"):
            wav_io.decode_pcm(x)


def test_feature_extraction_segmentation_matches_oracle():
    """inference/feature_extraction.py:114-140 -- same padding quirk as the style-transfer entry (a full extra zero segment
    when the length is an exact multiple), checked on the host against the oracle restatement."""
    from music_mixing_style_transfer_b200.inference.feature_extraction import FXencoder_Inference
    from oracle import networks_oracle as O
    obj = FXencoder_Inference.__new__(FXencoder_Inference)
    obj.segment_length, obj.batch_size = 250, 3
    for T in (250, 1000, 1001, 1749):
        song = torch.randn(2, T)
        ours = obj.batchwise_segmentization(song, "s")
        ref = O.batchwise_segmentization(song, 250, 3)
        assert len(ours) == len(ref) and all(torch.equal(a, b) for a, b in zip(ours, ref)), T
        assert sum(b.shape[0] for b in ours) == T // 250 + 1
    with pytest.raises(AssertionError):
        obj.batchwise_segmentization(torch.randn(2, 100), "s")


def test_interpolation_weights():
    w = shard.interpolation_weights(64, 16)
    assert w.shape == (64,) and float(w[0]) == 1.0 and float(w[15]) == 0.0 and float(w[16]) == 1.0
    assert torch.allclose(w[:16], (15 - torch.arange(16).float()) / 15)      # style_transfer.py:250 per segment
    with pytest.raises(ValueError):
        shard.interpolation_weights(4, 1)


def _chain_structure(chain):
    """(class name, processor name, probability, normalise flag) tree of an AugmentationChain, with the chain-level switches."""
    out = {"shuffle": bool(chain.shuffle), "parallel": bool(chain.parallel), "weight": chain.parallel_weight_factor, "fxs": []}
    for fx, p, norm in chain.fxs:
        if hasattr(fx, "fxs"):
            out["fxs"].append((_chain_structure(fx), p, norm))
        else:
            params = sorted((q.name, q.value) for q in fx.parameters) if type(fx).__name__ == "Equaliser" and len(fx.bands) == 1 else None
            out["fxs"].append((type(fx).__name__, fx.name, list(getattr(fx, "bands", [])), params, p, norm))
    return out


def test_instrument_chain_factory_structure():
    """create_inst_effects_augmentation_chain (audio_effects_chain.py:99-164): nested / parallel structure per instrument."""
    from music_mixing_style_transfer_b200.mixing_manipulator import create_inst_effects_augmentation_chain
    prob = {"eq": 0.9, "comp": 0.8, "pan": 0.7, "imager": 0.6, "reverb": 0.5, "gain": 1.0}
    s = _chain_structure(create_inst_effects_augmentation_chain("vocals", prob, algorithmic=True))
    assert [type(f[0]) for f in s["fxs"]] == [dict, dict, dict, str] and s["fxs"][3][:2] == ("Gain", "Gain") and s["fxs"][3][-1] is False
    assert s["fxs"][0][0]["shuffle"] and [f[0] for f in s["fxs"][0][0]["fxs"]] == ["Equaliser", "Compressor"]
    assert [f[0] for f in s["fxs"][1][0]["fxs"]] == ["Panner", "MidSideImager"]
    rv = s["fxs"][2][0]
    assert rv["parallel"] and rv["weight"] is None and [(f[0], f[-2]) for f in rv["fxs"]] == [("AlgorithmicReverb", 0.5)]
    d = _chain_structure(create_inst_effects_augmentation_chain("drums", prob, algorithmic=True))
    low, high = d["fxs"][2][0]["fxs"][0][0], d["fxs"][2][0]["fxs"][1][0]
    assert (low["parallel"], low["weight"], high["weight"]) == (True, 0.8, 0.6)
    assert low["fxs"][0][2] == ["high_shelf"] and high["fxs"][0][2] == ["low_shelf"]
    assert low["fxs"][0][3] == [("high_shelf_freq", 100.0), ("high_shelf_gain", -50.0)]
    assert abs(low["fxs"][1][-2] - 0.005) < 1e-12 and high["fxs"][1][-2] == 0.5


def test_normalizer_host_logic_matches_oracle():
    import numpy as np
    """The host-side pieces of the input FX normaliser (no GPU involved): inter-onset peak statistics, BS.1770 gating and the
    imager's three balance steps folded into one matrix, against oracle/norm_oracle.py (itself pinned to the reference)."""
    from music_mixing_style_transfer_b200.mixing_manipulator import data_normalization as dn
    from oracle import fixtures, norm_oracle as N
    x = fixtures.fx_input(4, 60000)
    got = dn.mean_peak(x[:, 0], N.stub_onsets, 44100, 75)
    ref = N.get_mean_peak(x[:, :1])
    assert got is not None and np.allclose(got, ref, rtol=0, atol=1e-9)
    assert dn.mean_peak(np.zeros(5000, np.float32), N.stub_onsets) is None                  # no onset -> the reference's TypeError path
    lo, hi = dn.gating_block_bounds(200000)
    lo_o, hi_o = N.gating_block_bounds(200000)
    assert np.array_equal(lo, lo_o) and np.array_equal(hi, hi_o)
    z = np.abs(np.random.RandomState(0).randn(2, len(lo))) * 50.0 + 1e-3
    assert abs(dn.gated_loudness(z) - N.gated_loudness(z)) < 1e-12
    assert abs(dn.gated_loudness(z[:1]) - N.gated_loudness(z[:1])) < 1e-12                  # mono meter of the EQ matching
    x64 = x.astype(np.float64)
    x64[:, 1] = 0.3 * x64[:, 1] + 0.5 * np.roll(x64[:, 0], 100)
    ll, rr, lr = np.sum(x64[:, 0] ** 2), np.sum(x64[:, 1] ** 2), np.sum(x64[:, 0] * x64[:, 1])
    M = dn.imager_matrix(ll, rr, lr, 0.93)
    ref = N.normalize_imager(x64, 0.93, 0.999)
    assert np.abs(x64 @ M.T - ref).max() <= 1e-9 * np.abs(ref).max()


def test_entry_flags_that_take_names():
    """--normalization_order / --instruments: the reference declares them with type=str2bool (only the defaults are usable
    there); here they also accept names, and the defaults are the reference's."""
    from music_mixing_style_transfer_b200.inference.style_transfer import build_parser
    a = build_parser().parse_args([])
    assert a.normalization_order == ['loudness', 'eq', 'compression', 'imager', 'loudness'] and a.normalize_input is True
    assert a.instruments == ["drums", "bass", "other", "vocals"]
    a = build_parser().parse_args(["--normalization_order", "loudness", "imager", "--instruments", "drums", "bass"])
    assert a.normalization_order == ['loudness', 'imager'] and a.instruments == ["drums", "bass"]
