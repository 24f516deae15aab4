"""End-to-end: segment-batched style transfer (encoder -> mean embedding -> TCN -> concat/crop) vs the oracle, the
public entry class on real-format WAV files, smoke()."""
import os
import wave

import numpy as np
import pytest
import torch

from gpu_helpers import RMS_TOL, err_stats, models, oracle_threads, state_dicts
from oracle import networks_oracle as O, weights as W

pytestmark = pytest.mark.gpu


def test_stem_style_transfer_matches_oracle():
    oracle_threads()
    enc, tcn = models()
    esd, tsd = state_dicts()
    seg, B = 16384, 3
    inp = W.synthetic_audio(1, 5 * seg, seed=61)[0]            # exact multiple -> the extra zero segment (quirk q1)
    ref = W.synthetic_audio(1, 3 * seg + 777, seed=62)[0]
    with torch.no_grad():
        y_ref, emb_ref = O.style_transfer_stem(inp, ref, esd, tsd, seg, seg, B, W.ENC_KERNELS, W.ENC_STRIDES)
        in_b = O.batchwise_segmentization(inp, seg, B)
        ref_b = O.batchwise_segmentization(ref, seg, B)
        emb = torch.cat([enc(b.cuda()) for b in ref_b], 0).mean(0)
        outs = [tcn(b.cuda(), emb.unsqueeze(0)).cpu() for b in in_b]
    y = torch.cat([torch.cat(torch.unbind(o, 0), -1) for o in outs], -1)[:, :inp.shape[-1]]
    ee, ey = err_stats(emb.cpu(), emb_ref), err_stats(y, y_ref)
    assert ee["max"] <= 1e-4 and ee["rel"] <= 2e-5, ee
    print("e2e stem parity", ee, ey)
    assert ey["rms"] <= RMS_TOL and ey["rms"] <= 2e-5, ey


def _write_wav(path, x):
    pcm = np.clip(np.rint(x.T * 32768.0), -32768, 32767).astype("<i2")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with wave.open(path, "wb") as w:
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(44100)
        w.writeframes(pcm.tobytes())


@pytest.mark.parametrize("interp", [False, True])
def test_public_entry_on_wav_files(tmp_path, interp):
    """inference/style_transfer.py surface: checkpoints with `module.` prefix, stems on disk, PCM_16 mixture out."""
    oracle_threads()
    from music_mixing_style_transfer_b200.inference import style_transfer as st
    esd, tsd = state_dicts()
    torch.save({"model": {"module." + k: v for k, v in esd.items()}}, tmp_path / "enc.pt")
    torch.save({"model": {"module." + k: v for k, v in tsd.items()}}, tmp_path / "tcn.pt")
    song = tmp_path / "data" / "song0"
    insts = ["drums", "bass", "other", "vocals"]
    seg = 16384   # >= 3 * 4096: the encoder's last k=5 layers need T > 2 for their reflection padding
    audio = {}
    for name, L in (("input", 3 * seg + 100), ("reference", 4 * seg + 5), ("reference_B", 2 * seg + 9)):
        for i, inst in enumerate(insts):
            x = W.synthetic_audio(1, L, seed=500 + 10 * len(name) + i)[0].numpy()
            x = np.clip(np.rint(x * 32768.0), -32768, 32767) / 32768.0        # what survives PCM_16
            audio[(name, inst)] = torch.from_numpy(x).float()
            _write_wav(str(song / "separated" / name / f"{inst}.wav"), x)
    argv = ["--target_dir", str(tmp_path / "data") + "/", "--output_dir", str(tmp_path / "out") + "/",
            "--ckpt_path_enc", str(tmp_path / "enc.pt"), "--ckpt_path_conv", str(tmp_path / "tcn.pt"),
            "--segment_length", str(seg), "--segment_length_ref", str(seg), "--batch_size", "2",
            "--normalize_input", "False", "--do_not_separate", "True"]
    if interp:
        argv += ["--interpolation", "True", "--interpolate_segments", "4"]
    st.main(argv)
    tag = "output_notnormed_interpolation" if interp else "output_notnormed"
    out_path = tmp_path / "out" / "song0" / f"mixture_{tag}.wav"
    got = st.load_wav_segment(str(out_path), axis=0)
    # oracle restatement of the same flow
    mix = 0
    with torch.no_grad():
        for inst in insts:
            inp, ref = audio[("input", inst)], audio[("reference", inst)]
            if not interp:
                y, _ = O.style_transfer_stem(inp, ref, esd, tsd, seg, seg, 2, W.ENC_KERNELS, W.ENC_STRIDES)
            else:
                S = 4
                iseg = inp.shape[-1] // S + 1
                in_b = O.batchwise_segmentization(inp, iseg, 2, min_length=seg)
                embs = []
                for r, cut in ((ref, seg), (audio[("reference_B", inst)], seg)):
                    rb = O.batchwise_segmentization(r, cut, 2) if r.shape[-1] > seg else [r.unsqueeze(0)]
                    embs.append(torch.cat([O.fxencoder_forward(b, esd, W.ENC_KERNELS, W.ENC_STRIDES) for b in rb], 0).mean(0))
                outs = []
                for idx, b in enumerate(in_b):
                    w = (S - 1 - idx) / (S - 1)
                    outs.append(O.tcn_forward(b, (w * embs[0] + (1 - w) * embs[1]).unsqueeze(0), tsd))
                y = torch.cat([torch.cat(torch.unbind(o, 0), -1) for o in outs], -1)[:, :inp.shape[-1]]
            mix = mix + y.numpy()
    ref_pcm = np.clip(np.rint(mix * 32768.0), -32768, 32767) / 32768.0
    assert got.shape == ref_pcm.shape
    # one PCM_16 LSB is 3e-5: allow isolated 1-LSB rounding flips on top of the 1e-4 RMS budget
    assert np.abs(got - ref_pcm).max() <= 2.0 / 32768.0 + 4e-4 and np.sqrt(np.mean((got - ref_pcm) ** 2)) <= RMS_TOL
    assert os.path.exists(tmp_path / "out" / "style_transfer_inference_configurations.txt")


def test_public_entry_with_input_normalizer(tmp_path):
    """The reference's default `--normalize_input True` path (data_loader.py:586-590): the INPUT stems go through the FX
    normaliser (here loudness -> eq -> imager -> loudness; 'compression' needs the aubio package) and are clamped, the
    reference stems are not.  Entry output vs the oracle normaliser + oracle networks on the same PCM."""
    oracle_threads()
    from music_mixing_style_transfer_b200.inference import style_transfer as st
    from oracle import norm_oracle as N
    esd, tsd = state_dicts()
    torch.save({"model": {"module." + k: v for k, v in esd.items()}}, tmp_path / "enc.pt")
    torch.save({"model": {"module." + k: v for k, v in tsd.items()}}, tmp_path / "tcn.pt")
    insts, seg = ["drums", "bass"], 16384
    f = np.arange(32769) / 32768.0
    feats = {"eq": {}, "loudness": {}, "imager": {}}
    for i, inst in enumerate(insts):
        feats["eq"][inst] = (30.0 / (1.0 + (150.0 + 50.0 * i) * f) + 0.02).astype(np.float32)
        feats["loudness"][inst] = np.array([-20.0 - i])
        feats["imager"][inst] = np.float32(0.9 + 0.02 * i)
    np.save(tmp_path / "feats.npy", feats, allow_pickle=True)
    song = tmp_path / "data" / "song0"
    audio = {}
    for name, L in (("input", 2 * seg + 501), ("reference", 2 * seg)):
        for i, inst in enumerate(insts):
            x = W.synthetic_audio(1, L, seed=700 + 10 * len(name) + i)[0].numpy()
            x[1] = 0.4 * x[1] + 0.5 * np.roll(x[0], 300)                      # wide enough to stay off the randomised Haas branch
            x = np.clip(np.rint(x * 32768.0), -32768, 32767) / 32768.0
            audio[(name, inst)] = x.astype(np.float32)
            _write_wav(str(song / "separated" / name / f"{inst}.wav"), x)
    order = ["loudness", "eq", "imager", "loudness"]
    st.main(["--target_dir", str(tmp_path / "data") + "/", "--output_dir", str(tmp_path / "out") + "/",
             "--ckpt_path_enc", str(tmp_path / "enc.pt"), "--ckpt_path_conv", str(tmp_path / "tcn.pt"),
             "--segment_length", str(seg), "--segment_length_ref", str(seg), "--batch_size", "2", "--instruments", *insts,
             "--normalize_input", "True", "--precomputed_normalization_feature", str(tmp_path / "feats.npy"),
             "--normalization_order", *order, "--do_not_separate", "True"])
    got = st.load_wav_segment(str(tmp_path / "out" / "song0" / "mixture_output.wav"), axis=0)
    smooth = dict(feats, eq={k: N.smooth_eq_feature(v, k) for k, v in feats["eq"].items()})
    mix = 0
    with torch.no_grad():
        for inst in insts:
            xin = N.normalize_audio(np.ascontiguousarray(audio[("input", inst)].T), order, smooth, src=inst).T
            inp = torch.from_numpy(np.clip(xin, -1.0, 1.0).astype(np.float32))
            y, _ = O.style_transfer_stem(inp, torch.from_numpy(audio[("reference", inst)]), esd, tsd, seg, seg, 2,
                                         W.ENC_KERNELS, W.ENC_STRIDES)
            mix = mix + y.numpy()
    ref_pcm = np.clip(np.rint(mix * 32768.0), -32768, 32767) / 32768.0
    assert got.shape == ref_pcm.shape
    assert np.abs(got - ref_pcm).max() <= 2.0 / 32768.0 + 4e-4 and np.sqrt(np.mean((got - ref_pcm) ** 2)) <= RMS_TOL


def test_smoke_entry():
    import __graft_entry__
    __graft_entry__.smoke()


def test_device_io_is_bit_identical_to_host_io(tmp_path):
    """--device_io True (PCM decode / remix / PCM_16 on the GPU, csrc/pcm.cu) must write the same bytes as the host path."""
    from music_mixing_style_transfer_b200.inference import style_transfer as st
    esd, tsd = state_dicts()
    torch.save({"model": {"module." + k: v for k, v in esd.items()}}, tmp_path / "enc.pt")
    torch.save({"model": {"module." + k: v for k, v in tsd.items()}}, tmp_path / "tcn.pt")
    song = tmp_path / "data" / "song0"
    insts = ["drums", "bass", "other", "vocals"]
    seg = 16384
    for name, L in (("input", 2 * seg + 77), ("reference", 3 * seg + 1)):
        for i, inst in enumerate(insts):
            x = W.synthetic_audio(1, L, seed=900 + 10 * len(name) + i)[0].numpy() * 3.0   # loud: the remix clips
            _write_wav(str(song / "separated" / name / f"{inst}.wav"), x)
    outs = {}
    for mode in ("True", "False"):
        out_dir = tmp_path / f"out_{mode}"
        st.main(["--target_dir", str(tmp_path / "data") + "/", "--output_dir", str(out_dir) + "/",
                 "--ckpt_path_enc", str(tmp_path / "enc.pt"), "--ckpt_path_conv", str(tmp_path / "tcn.pt"),
                 "--segment_length", str(seg), "--segment_length_ref", str(seg), "--batch_size", "3",
                 "--normalize_input", "False", "--do_not_separate", "True", "--save_each_inst", "True",
                 "--device_io", mode])
        files = sorted(os.listdir(out_dir / "song0"))
        assert "mixture_output_notnormed.wav" in files and "drums_output_notnormed.wav" in files
        outs[mode] = {f: open(out_dir / "song0" / f, "rb").read() for f in files if f.endswith(".wav")}
    assert outs["True"].keys() == outs["False"].keys()
    for f in outs["True"]:
        assert outs["True"][f] == outs["False"][f], f


def test_feature_extraction_entry(tmp_path):
    """inference/feature_extraction.py surface (BASELINE config 1 semantics): one stereo and one mono WAV, DDP checkpoint,
    `<name>_fx_embedding.npy` = mean embedding over the zero-padded segments, vs the oracle encoder on the same PCM."""
    oracle_threads()
    from music_mixing_style_transfer_b200.inference import feature_extraction as fe
    esd, _ = state_dicts()
    torch.save({"model": {"module." + k: v for k, v in esd.items()}}, tmp_path / "enc.pt")
    seg = 16384
    clips = {}
    x = W.synthetic_audio(1, 2 * seg + 311, seed=71)[0].numpy()
    x = np.clip(np.rint(x * 32768.0), -32768, 32767) / 32768.0
    clips["a/stereo.wav"] = x
    _write_wav(str(tmp_path / "data" / "a" / "stereo.wav"), x)
    m = W.synthetic_audio(1, seg, seed=72)[0, :1].numpy()                      # exact multiple: extra all-zero segment
    m = np.clip(np.rint(m * 32768.0), -32768, 32767) / 32768.0
    os.makedirs(tmp_path / "data" / "b", exist_ok=True)
    with wave.open(str(tmp_path / "data" / "b" / "mono.wav"), "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(44100)
        w.writeframes(np.rint(m[0] * 32768.0).astype("<i2").tobytes())
    clips["b/mono.wav"] = np.concatenate([m, m], axis=0)
    fe.main(["--target_dir", str(tmp_path / "data") + "/", "--output_dir", str(tmp_path / "emb") + "/",
             "--ckpt_path_enc", str(tmp_path / "enc.pt"), "--segment_length", str(seg), "--batch_size", "2"])
    assert os.path.exists(tmp_path / "emb" / "feature_extraction_inference_configurations.txt")
    for rel, audio in clips.items():
        got = np.load(tmp_path / "emb" / rel.replace(".wav", "_fx_embedding.npy"))
        song = torch.from_numpy(audio).float()
        batches = O.batchwise_segmentization(song, seg, 2)
        with torch.no_grad():
            ref = torch.cat([O.fxencoder_forward(b, esd, W.ENC_KERNELS, W.ENC_STRIDES) for b in batches], 0).mean(0).numpy()
        assert got.shape == (2048,) and got.dtype == np.float32
        assert np.abs(got - ref).max() <= 1e-4 and np.linalg.norm(got - ref) / np.linalg.norm(ref) <= 2e-5, rel
    with pytest.raises(RuntimeError):
        fe.main(["--target_dir", str(tmp_path / "data") + "/", "--ckpt_path_enc", str(tmp_path / "enc.pt"),
                 "--inference_device", "cpu"])
