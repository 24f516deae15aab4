"""Host I/O around the forward with the sample-format work on the device.

Mirrors, with the reference's semantics (paths relative to /root/reference/):
  * `load_wav_segment`   mixing_style_transfer/data_loader/loader_utils.py:47-70  -- sample-rate / bit-depth checks and
    `ValueError`s on the host (stdlib `wave`), the int -> float conversion and de-interleave on the GPU (`mst_pcm_decode`)
  * stem clamp           mixing_style_transfer/data_loader/data_loader.py:589-590
  * mono duplication     inference/feature_extraction.py:87-89
  * remix + PCM_16 file  inference/style_transfer.py:165-177 (`sum(inst_outputs)`, `sf.write(..., 'PCM_16')`) via
    `mst_pcm_encode_mix`: one int16 mixture crosses PCIe instead of one fp32 waveform per instrument.
Raw PCM is staged through pinned host memory; the copies are asynchronous on the current stream.
`soundfile` / libsndfile are not in this image: the PCM_16 quantisation (scale 2^15, round half to even, clip) restates
what `write_wav_pcm16` of the inference entry already did on the host, and the GPU path is bit-identical to it.
"""
import wave

import numpy as np
import torch

from . import _cabi


def read_wav_pcm(audio_path, start_point=None, duration=None, sample_rate=44100):
    """Header checks of load_wav_segment (loader_utils.py:47-63); returns the raw interleaved PCM as a numpy view
    [n_frames, n_channels] of int16 / int32 without converting it."""
    start_point = 0 if start_point is None else start_point
    with wave.open(audio_path, 'r') as pt_wav:
        duration = pt_wav.getnframes() if duration is None else duration
        if pt_wav.getframerate() != sample_rate:
            raise ValueError(f"ValueError: input audio's sample rate should be {sample_rate}")
        pt_wav.setpos(start_point)
        raw = pt_wav.readframes(duration)
        width, n_ch = pt_wav.getsampwidth(), pt_wav.getnchannels()
    if width == 2:
        x = np.frombuffer(raw, dtype='<i2')
    elif width == 4:
        x = np.frombuffer(raw, dtype='<i4')
    else:
        raise ValueError("ValueError: input audio's bit depth should be 16 or 32-bit")
    if n_ch not in (1, 2):
        raise ValueError(f"ValueError: {n_ch}-channel audio is not supported (mono or stereo)")
    return x.reshape(-1, n_ch)


def pin_pcm(pcm):
    """numpy int16 / int32 [n_frames, n_channels] (the read-only file buffer) -> the same samples in a pinned torch tensor,
    ready for an asynchronous H2D copy.  Safe to call from a loader thread."""
    if pcm.dtype not in (np.dtype('<i2'), np.dtype('<i4')):
        raise ValueError("ValueError: input audio's bit depth should be 16 or 32-bit")
    tdt = torch.int16 if pcm.dtype == np.dtype('<i2') else torch.int32
    host = torch.empty(pcm.shape, dtype=tdt, pin_memory=bool(pcm.size) and torch.cuda.is_available())
    if pcm.size:
        np.copyto(host.numpy(), pcm)
    return host


def decode_pcm(pcm, device=None, out=None):
    """pcm: numpy or torch int16 / int32 [n_frames, n_channels] (host or device) -> float32 [2, n_frames] on the GPU:
    x / 2^15 (or 2^31), clamped to [-1, 1], de-interleaved; mono is duplicated into both channels."""
    if not torch.cuda.is_available():
        raise RuntimeError("wav_io.decode_pcm needs a CUDA device (this engine has no CPU path)")
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if isinstance(pcm, np.ndarray):
        if pcm.dtype not in (np.dtype('<i2'), np.dtype('<i4')):
            raise ValueError("ValueError: input audio's bit depth should be 16 or 32-bit")
        # one host copy: the (read-only) file buffer straight into a pinned staging tensor, then an async H2D copy
        pcm = pin_pcm(pcm)
    if pcm.dtype not in (torch.int16, torch.int32):
        raise ValueError("ValueError: input audio's bit depth should be 16 or 32-bit")
    if pcm.dim() == 1:
        pcm = pcm.unsqueeze(1)
    pcm = pcm.to(device, non_blocking=True).contiguous()
    n_frames, n_ch = int(pcm.shape[0]), int(pcm.shape[1])
    if out is None:
        out = torch.empty(2, n_frames, dtype=torch.float32, device=device)
    if tuple(out.shape) != (2, n_frames) or out.dtype != torch.float32 or out.stride(1) != 1:
        raise RuntimeError(f"decode_pcm: out must be float32 [2, {n_frames}] with unit stride in time")
    if n_frames:
        _cabi.check(_cabi.lib().mst_pcm_decode(_cabi.ptr(pcm), pcm.element_size(), n_ch, n_frames, _cabi.ptr(out),
                                               out.stride(0), _cabi.current_stream()), "pcm_decode")
    return out


def load_wav_to_device(audio_path, device=None, sample_rate=44100):
    """load_wav_segment(path, axis=0) + clamp + mono duplication, as one H2D copy of raw PCM and one kernel."""
    return decode_pcm(read_wav_pcm(audio_path, sample_rate=sample_rate), device)


def encode_mix_pcm16(stems, n_frames=None):
    """stems: float32 CUDA tensor [n_stems, 2, T] (or [2, T]) -> int16 CUDA tensor [n_frames, 2]:
    the float32 sum over the stems in order, then clip(rint(x * 32768), -32768, 32767)."""
    if stems.dim() == 2:
        stems = stems.unsqueeze(0)
    if stems.dim() != 3 or stems.shape[1] != 2 or stems.dtype != torch.float32:
        raise RuntimeError(f"encode_mix_pcm16 expects float32 [n_stems, 2, T], got {tuple(stems.shape)} {stems.dtype}")
    stems = stems.contiguous()
    T = int(stems.shape[2])
    n_frames = T if n_frames is None else int(n_frames)
    if n_frames > T:
        raise RuntimeError("encode_mix_pcm16: n_frames exceeds the stem length")
    pcm = torch.empty(n_frames, 2, dtype=torch.int16, device=stems.device)
    if n_frames:
        _cabi.check(_cabi.lib().mst_pcm_encode_mix(_cabi.ptr(stems), int(stems.shape[0]), T, n_frames, _cabi.ptr(pcm),
                                                   _cabi.current_stream()), "pcm_encode_mix")
    return pcm


def write_wav_pcm16_from_device(path, stems, sample_rate, n_frames=None):
    """Remix + quantise on the GPU, copy the int16 frames back and write the RIFF file (style_transfer.py:174-177)."""
    pcm = encode_mix_pcm16(stems, n_frames)
    host = torch.empty(pcm.shape, dtype=torch.int16).pin_memory() if pcm.numel() else torch.empty(pcm.shape, dtype=torch.int16)
    host.copy_(pcm, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    with wave.open(path, 'wb') as w:
        w.setnchannels(2)
        w.setsampwidth(2)
        w.setframerate(sample_rate)
        w.writeframes(host.numpy().astype('<i2', copy=False).tobytes())
