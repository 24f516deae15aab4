#!/usr/bin/env python
"""bench.py -- BASELINE metric: stereo-audio seconds per second through the full style-transfer forward.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1], per GPU): reference batch [32, 2, 262144] -> FXencoder -> mean embedding;
input batch [32, 2, 262144] -> MixFXcloner TCN conditioned on it -> clamp.  Synthetic seeded stereo 44.1 kHz audio,
seeded random weights loaded through the reference's state_dict layout (no checkpoints ship with the reference).
N > 1 (weak scaling): every rank converts its own 32 input segments (32*N in total).  The reference batch is sharded too
(every rank encodes 32/N of its segments, ONE 8 KiB all-reduce of the partial sums -- SURVEY.md 8e; `--ref-mode broadcast`
selects rank-0 encode + broadcast instead) and the output segments are all-gathered in sub-batches that overlap the TCN of
the next sub-batch (shard.py).  One "step" = one full forward over that batch.  `value` = audio seconds of input converted
by ALL ranks per second, inputs resident in HBM; `e2e` = same through the public host-to-host call
(pipeline.StyleTransferPipeline): pinned-host inputs, H2D of every step's inputs and D2H of its output waveforms inside the
timed region, copies on side streams.
`extra` carries the other BASELINE configs as short legs: configs[2] (FX chain), configs[3]'s 64 segments per GPU, configs[4]
(interpolation, per-row conditioning, one odd length), the sample-format kernels, a file-to-file run of the inference entry,
and -- at N = 1 -- the reference's own modules on this GPU through cuDNN (what inference/style_transfer.py:29-32 runs as
shipped) as a secondary baseline.

--impl reference: the reference's own CPU implementation of the path (oracle port of its torch modules -- a Python
reference cannot travel to the GPU box as source) on the host cores, a bounded sample per step.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SR = 44100
SEG_LEN = 262144
BATCH_PER_GPU = 32
METRIC = "stereo-audio sec/sec through full style-transfer forward"
UNIT = "audio_s/s"
# algorithmic FLOPs of one dilated block n>=1 per (segment, sample): 2 * 128 * (128*15)   
FLOP_PER_ROW_UMMA = 2 * 128 * 128 * 15
TCN_FLOP_PER_SAMPLE = 6397952        # whole TCN, SURVEY.md 8d
ENC_FLOP_PER_SEG = 28.59e9           # encoder at L = 2^18
TRAFFIC_PER_LAUNCH = 9.98e9          # dram bytes per tcn_block_umma_kernel launch at configs[1], from the committed ncu capture (r02e)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d["bf16_tflops"]),
                    "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "source": "measured"}
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, line in self.lines:
            if ts < t0 or ts > t1 + 0.2:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_models(device):
    import torch
    import yaml
    from music_mixing_style_transfer_b200.networks import FXencoder, TCNModel
    from music_mixing_style_transfer_b200 import synthetic as W
    cfg = yaml.full_load(open(os.path.join(ROOT, "music_mixing_style_transfer_b200", "inference", "configs.yaml")))
    c = cfg["TCN"]["default"]
    enc = FXencoder(cfg["Effects_Encoder"]["default"])
    tcn = TCNModel(nparams=c["condition_dimension"], ninputs=2, noutputs=2, nblocks=c["nblocks"],
                   dilation_growth=c["dilation_growth"], kernel_size=c["kernel_size"], channel_width=c["channel_width"],
                   stack_size=c["stack_size"], cond_dim=c["condition_dimension"], causal=c["causal"])
    enc.load_state_dict(W.make_encoder_state_dict(0))
    tcn.load_state_dict(W.make_tcn_state_dict(0))
    return enc.to(device).eval(), tcn.to(device).eval()


def cpu_reference_step(state, n_seg=1, length=SEG_LEN):
    """One bounded CPU sample of the same forward: encoder on n_seg reference segments + TCN on n_seg input segments,
    through the oracle port of the reference's torch modules.  Returns audio seconds converted."""
    import torch
    from oracle import networks_oracle as O, weights as W
    if "sd" not in state:
        state["sd"] = (W.make_encoder_state_dict(0), W.make_tcn_state_dict(0))
        state["ref"] = W.synthetic_audio(n_seg, length, seed=1234)
        state["inp"] = W.synthetic_audio(n_seg, length, seed=1235)
    esd, tsd = state["sd"]
    with torch.no_grad():
        emb = O.fxencoder_forward(state["ref"], esd, W.ENC_KERNELS, W.ENC_STRIDES).mean(dim=0)
        out = O.tcn_forward(state["inp"], emb.unsqueeze(0), tsd)
    state["last"] = float(out.abs().mean())
    state["out"] = out
    return n_seg * length / SR


def pick_cpu_threads():
    """The reference arm may use every host thread it can USE: torch's CPU convolutions do not scale to 128 threads on
    this box, so time a short TCN forward at a few intra-op thread counts and keep the fastest."""
    import torch
    from oracle import networks_oracle as O, weights as W
    cores = os.cpu_count() or 1
    tsd = W.make_tcn_state_dict(0)
    x = W.synthetic_audio(1, 16384, seed=7)
    cond = torch.zeros(1, 2048)
    best, best_t = cores, float("inf")
    for n in sorted({cores, 64, 32, 16, 8}, reverse=True):
        if n > cores:
            continue
        torch.set_num_threads(n)
        with torch.no_grad():
            O.tcn_forward(x, cond, tsd)
            t0 = time.perf_counter()
            O.tcn_forward(x, cond, tsd)
            dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = n, dt
    torch.set_num_threads(best)
    return best


def run_reference(args, rank):
    import torch
    if rank != 0:
        return
    cores = pick_cpu_threads()
    state = {}
    for _ in range(args.warmup):
        cpu_reference_step(state)
    t0 = time.perf_counter()
    secs = 0.0
    for _ in range(args.steps):
        secs += cpu_reference_step(state)
    dt = time.perf_counter() - t0
    value = secs / dt
    sample = f"per step: 1 reference + 1 input segment of {SEG_LEN} stereo samples (of the {BATCH_PER_GPU}-segment batch)"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, args.steps), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: FXencoder+MixFXcloner full forward, segments of 262144 stereo samples",
                       "segment_length": SEG_LEN, "batch_per_gpu": BATCH_PER_GPU, "cpu_sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "host_cores": os.cpu_count(), "kind": "port",
                             "sample": sample, "torch": torch.__version__},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def reference_modules_on_gpu(device, B, L, cpu_out=None, cpu_in=None):
    """Secondary baseline (N = 1, rank 0): the oracle port of the reference's torch modules on THIS GPU through cuDNN -- what
    inference/style_transfer.py:29-32 runs as shipped when CUDA is available.  Twice: fp32 with TF32 off, and with torch's
    defaults for convolutions (TF32 on).  Its error against the CPU forward is measured on the cpu_baseline sample."""
    import torch
    from oracle import networks_oracle as O, weights as W
    esd = {k: v.to(device) for k, v in W.make_encoder_state_dict(0).items()}
    tsd = {k: v.to(device) for k, v in W.make_tcn_state_dict(0).items()}
    B = min(B, 8)            # bounded: throughput is per audio second, eager fp32 keeps ~5 activation tensors of B x 134 MB alive
    ref = W.synthetic_audio(B, L, seed=1234).to(device)
    inp = W.synthetic_audio(B, L, seed=2000).to(device)
    out = {"batch": B}
    for name, tf32 in (("fp32", False), ("tf32_default", True)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32

        def step(r, x):
            with torch.no_grad():
                emb = O.fxencoder_forward(r, esd, W.ENC_KERNELS, W.ENC_STRIDES).mean(dim=0)
                return O.tcn_forward(x, emb.unsqueeze(0), tsd)
        try:
            step(ref[:2], inp[:2])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            step(ref, inp)                      # warm (cuDNN algorithm selection)
            e0.record()
            for _ in range(2):
                y = step(ref, inp)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 2
            rec = {"ms_per_step": ms, "audio_s_per_s": B * L / SR / (ms * 1e-3)}
            if cpu_out is not None:
                y1 = step(cpu_in[0].to(device), cpu_in[1].to(device)).cpu()
                rec["rms_err_vs_cpu_forward"] = float((y1 - cpu_out).pow(2).mean().sqrt())
            del y
            out[name] = rec
        except Exception as exc:
            out[name] = {"error": repr(exc)}
        torch.cuda.empty_cache()
    torch.backends.cudnn.allow_tf32 = True
    return out


def file_to_file_leg(device, tmp_root):
    """The inference entry on a synthetic song directory in the reference's layout (2 songs x 4 stems x 60 s, PCM_16):
    WAV files in, mixture WAV out; audio seconds of song per wall second, file reads and writes included."""
    import shutil
    import wave
    import numpy as np
    import torch
    from music_mixing_style_transfer_b200 import synthetic as W
    from music_mixing_style_transfer_b200.inference import style_transfer as st
    root = os.path.join(tmp_root, "mst_bench_songs")
    shutil.rmtree(root, ignore_errors=True)
    n_frames = 60 * SR
    for song in ("song0", "song1"):
        for name in ("input", "reference"):
            d = os.path.join(root, "data", song, "separated", name)
            os.makedirs(d)
            for i, inst in enumerate(("drums", "bass", "other", "vocals")):
                x = W.synthetic_audio(1, n_frames, seed=100 * int(song[-1]) + 10 * len(name) + i)[0].numpy()
                with wave.open(os.path.join(d, inst + ".wav"), "wb") as w:
                    w.setnchannels(2); w.setsampwidth(2); w.setframerate(SR)
                    w.writeframes(np.clip(np.rint(x.T * 32768.0), -32768, 32767).astype("<i2").tobytes())
    torch.save({"model": {"module." + k: v for k, v in W.make_encoder_state_dict(0).items()}}, os.path.join(root, "enc.pt"))
    torch.save({"model": {"module." + k: v for k, v in W.make_tcn_state_dict(0).items()}}, os.path.join(root, "tcn.pt"))
    argv = ["--target_dir", os.path.join(root, "data") + "/", "--output_dir", os.path.join(root, "out") + "/",
            "--ckpt_path_enc", os.path.join(root, "enc.pt"), "--ckpt_path_conv", os.path.join(root, "tcn.pt"),
            "--segment_length", str(SEG_LEN), "--segment_length_ref", str(SEG_LEN), "--do_not_separate", "True"]
    # the input FX normaliser on the GPU with synthetic per-stem targets; without 'compression', whose onset
    # detector (aubio, on the host as in the reference) is not installed on the bench box
    f = np.arange(32769) / 32768.0
    feats = {"eq": {}, "loudness": {}, "imager": {}}
    for i, inst in enumerate(("drums", "bass", "other", "vocals")):
        feats["eq"][inst] = (30.0 / (1.0 + (150.0 + 50.0 * i) * f) + 0.02).astype(np.float32)
        feats["loudness"][inst] = np.array([-28.0 - i])
        feats["imager"][inst] = np.float32(0.93 + 0.01 * i)
    np.save(os.path.join(root, "feats.npy"), feats, allow_pickle=True)
    argv_norm = argv + ["--normalize_input", "True", "--precomputed_normalization_feature", os.path.join(root, "feats.npy"),
                        "--normalization_order", "loudness", "eq", "imager", "loudness"]
    argv = argv + ["--normalize_input", "False"]
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        st.main(argv)                       # warm: weight packing, allocator
        stats = dict(st.main(argv))
        try:
            st.main(argv_norm)
            stats_norm = dict(st.main(argv_norm))
            normalized = {"audio_s_per_s": stats_norm["audio_seconds"] / stats_norm["wall_seconds"], "wall_s": stats_norm["wall_seconds"],
                          "normalization_order": ["loudness", "eq", "imager", "loudness"]}
        except Exception as exc:
            normalized = {"error": repr(exc)}
    shutil.rmtree(root, ignore_errors=True)
    return {"workload": "inference/style_transfer.py entry, 2 songs x 4 stems x 60 s PCM_16 WAV in -> mixture WAV out, "
                        f"segment_length {SEG_LEN}, default batch_size 1",
            "audio_s_per_s": stats["audio_seconds"] / stats["wall_seconds"], "wall_s": stats["wall_seconds"],
            "song_seconds": stats["audio_seconds"], "with_input_normalizer": normalized,
            "note": "song seconds (4 stems each) per wall second, file reads / decode / remix / PCM_16 / file writes included"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help="segments per GPU (default: BASELINE config 2)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the `extra` legs (profiling runs)")
    ap.add_argument("--precision", default="auto", choices=["auto", "f16f8", "bf16x3"],
                    help="TCN operand format (auto = f16f8 with the range guard and a bf16x3 repeat when it fires)")
    ap.add_argument("--ref-mode", default="allreduce", choices=["allreduce", "broadcast"],
                    help="N > 1: shard the reference batch (all-reduce of partial sums) or encode on rank 0 and broadcast")
    ap.add_argument("--gather-chunks", type=int, default=4, help="N > 1: sub-batches whose all-gather overlaps the next TCN")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from music_mixing_style_transfer_b200 import _cabi, shard
    from music_mixing_style_transfer_b200.pipeline import StyleTransferPipeline

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    _cabi.check(_cabi.lib().mst_device_check(local_rank), "device_check")

    B, L = args.batch, SEG_LEN
    total = B * world
    enc, tcn = build_models(device)
    tcn.precision = args.precision
    from music_mixing_style_transfer_b200 import synthetic as W
    shard_ref = world > 1 and args.ref_mode == "allreduce"
    ref_lo, ref_hi = shard.shard_bounds(B, world, rank) if shard_ref else (0, B)
    holds_ref = shard_ref or rank == 0
    # the reference batch is the same seeded [B, 2, L] everywhere; a rank keeps the segments it encodes
    ref_host = W.synthetic_audio(B, L, seed=1234)[ref_lo:ref_hi].contiguous().pin_memory() if holds_ref else None
    inp_host = W.synthetic_audio(B, L, seed=2000 + rank).pin_memory()
    ref_dev = ref_host.to(device) if holds_ref else None
    inp_dev = inp_host.to(device)
    chunks = args.gather_chunks if (world > 1 and B % max(1, args.gather_chunks) == 0) else 1
    step_kw = dict(gather=True, shard_reference=shard_ref, n_reference=B, gather_chunks=chunks)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        with torch.no_grad():
            return shard.sharded_style_transfer(enc, tcn, ref_dev, inp_dev, total, **step_kw)

    pipe = StyleTransferPipeline(enc, tcn, device, total, depth=2, **step_kw)

    def run_e2e(steps):
        """`steps` host-to-host steps through the public pipeline; every step's inputs come from pinned host memory and its
        waveforms end in pinned host memory before this returns."""
        for i in range(steps):
            if i >= pipe.depth:
                pipe.collect()
            pipe.submit(ref_host, inp_host)
        pipe.drain()

    def timed(fn, steps, whole=False):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        ev0.record()
        if whole:
            fn(steps)
        else:
            for _ in range(steps):
                fn()
        ev1.record()
        barrier()
        t1 = time.time()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), t0, t1

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    ms_total, t0, t1 = timed(step_resident, args.steps)
    clocks = sampler.stop(t0, t1)
    ms_step = ms_total / args.steps
    value = total * L / SR / (ms_step / 1e3)

    run_e2e(3)
    ms_e2e, _, _ = timed(run_e2e, args.steps, whole=True)
    e2e_value = total * L / SR / (ms_e2e / args.steps / 1e3)
    h2d = (inp_host.numel() + (ref_host.numel() if holds_ref else 0)) * 4
    d2h = B * 2 * L * 4

    # ---- roofline leg: per-launch CUDA events around the dominant kernel (tcn_block_umma_kernel) ----
    pk = peaks()
    cond = torch.zeros(1, 2048, device=device)
    events = []

    def on_launch(name, phase):
        if name != "tcn_block_umma_kernel":
            return
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        events.append((phase, ev))

    with torch.no_grad():
        tcn.forward_layers(inp_dev, cond)                    # warm
        events.clear()
        n_prof = 3
        for _ in range(n_prof):
            tcn.forward_layers(inp_dev, cond, on_launch)
    torch.cuda.synchronize()
    durs = [events[i][1].elapsed_time(events[i + 1][1]) for i in range(0, len(events), 2)]
    umma_ms = sum(durs) / max(1, len(durs))
    flops_per_launch = FLOP_PER_ROW_UMMA * B * L
    achieved = flops_per_launch / (umma_ms * 1e-3) / 1e12
    n_umma = 13
    share = n_umma * umma_ms / ms_step
    f8 = tcn.precision != "bf16x3"
    units = 2 if f8 else 3     # tensor-pipe time per algorithmic MMA in bf16-MMA equivalents (an e4m3 MMA runs at twice the rate)
    roofline = {"kernel": "tcn_block_umma_kernel", "bound": "tensor", "achieved": achieved,
                "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["bf16_tflops_sustained"],
                "peak_source": f"{pk['source']} bf16 dense (sustained: kernel timed inside a long step)",
                # dram__bytes_read.sum + dram__bytes_write.sum, mean over the 13 launches of a step, from the committed ncu capture
                # of this workload (profiles/r02e_tcn_ncu.csv: 6.04 GB read + 3.94 GB written = 1.16x the algorithmic bytes; the
                # small-dilation layers read 4.4-5.1 GB, the far-paired ones 6-9 GB depending on the box)
                "traffic": TRAFFIC_PER_LAUNCH if (B == BATCH_PER_GPU and L == SEG_LEN and f8) else None,
                "traffic_algorithmic": 2.0 * B * L * 512, "ms_per_launch": umma_ms, "launches_per_step": n_umma, "share_of_step": share,
                "algorithmic_flop_per_launch": flops_per_launch,
                "note": ("fp32-grade parity needs split operands: fp16 main product + two e4m3 correction products (each at twice "
                         "the bf16 rate) = 2 bf16-MMA equivalents per algorithmic MMA, so frac <= 0.5 by construction; "
                         "tensor_pipe_frac = 2*frac.  The kernel runs at the board's 1000 W cap: its time is energy per "
                         "launch / cap (profiles/r02_tcn_ablation.md)") if f8 else
                        ("fp32-grade parity needs the 3-product bf16 split: tensor-pipe work is 3x the algorithmic FLOPs, "
                         "so frac <= 0.333 by construction; tensor_pipe_frac = 3*frac"),
                "tensor_pipe_frac": units * achieved / pk["bf16_tflops_sustained"]}

    extra = {}
    if not args.no_extra:
        # ---- configs[3]: 64 input segments per GPU (512 over 8 GPUs), same reference batch, a few steps on every rank ----
        try:
            x64 = torch.cat([inp_dev, inp_dev.flip(0)], dim=0) if B == BATCH_PER_GPU else None
            if x64 is not None:
                def step64():
                    with torch.no_grad():
                        return shard.sharded_style_transfer(enc, tcn, ref_dev, x64, 2 * total, gather=True,
                                                            shard_reference=shard_ref, n_reference=B, gather_chunks=chunks)
                step64()
                ms64, _, _ = timed(step64, 3)
                extra["config4_64_per_gpu"] = {
                    "workload": f"configs[3]: full forward, {2 * total} input segments of 262144 sharded 64 per GPU over {world} GPU(s), "
                                "reference batch 32", "ms_per_step": ms64 / 3, "audio_s_per_s": 2 * total * L / SR / (ms64 / 3 / 1e3)}
            del x64
        except Exception as exc:
            extra["config4_64_per_gpu"] = {"error": repr(exc)}
        torch.cuda.empty_cache()
        # ---- configs[4]: interpolation, two reference embeddings, per-ROW conditioning, 16 rows per GPU; L = 2^18 and 82,412 ----
        try:
            rows, S = 16, 16
            legs = {}
            for L5 in (SEG_LEN, 82412):
                xi = inp_dev[:rows, :, :L5].contiguous()
                ra = ref_dev[:2] if (rank == 0 and ref_dev is not None) else (W.synthetic_audio(2, L, seed=1234).to(device) if rank == 0 else None)
                rb = W.synthetic_audio(2, L, seed=1300).to(device) if rank == 0 else None
                wts = shard.interpolation_weights(rows * world, S, device=device)

                def step5():
                    with torch.no_grad():
                        return shard.sharded_interpolation(enc, tcn, ra, rb, xi, rows * world, wts, gather=True)
                step5()
                ms5, _, _ = timed(step5, 3)
                legs[str(L5)] = {"ms_per_step": ms5 / 3, "audio_s_per_s": rows * world * L5 / SR / (ms5 / 3 / 1e3)}
            extra["config5_interpolation"] = {
                "workload": f"configs[4]: interpolation mode, 2 reference embeddings (1 broadcast of [2, 2048]), interpolate_segments {S}, "
                            f"{rows} rows per GPU x {world} GPU(s), cond [rows, 2048] per row; segment lengths 262144 and 82412",
                "by_segment_length": legs}
        except Exception as exc:
            extra["config5_interpolation"] = {"error": repr(exc)}
        torch.cuda.empty_cache()

    # ---- extra: BASELINE config 3 (FX chain only, B=256 random-parameter segments) on rank 0 ----
    if rank == 0 and not args.no_extra:
        try:
            import numpy as np
            from music_mixing_style_transfer_b200.mixing_manipulator import fx_chain_forward
            gen = torch.Generator(device=device)
            gen.manual_seed(1234)
            fb = 256
            fx_x = (torch.randn(fb, 2, L, generator=gen, device=device) * 0.1).clamp_(-1, 1)
            rng = np.random.RandomState(1234)
            lo = np.array([-15, 30, -15, 200, .1, -15, 1000, .1, -15, 3000, .1, -15, 5000, -80, 1, 50, 4, 0, -6, 0], np.float32)
            hi = np.array([15, 200, 15, 1000, 2, 15, 3000, 2, 15, 8000, 2, 15, 10000, -5, 20, 500, 40, 2, 9, 1], np.float32)
            fx_p = torch.from_numpy((lo + rng.rand(fb, 20).astype(np.float32) * (hi - lo))).to(device)
            fx_p[:, 19] = (fx_p[:, 19] >= 0.5).float()
            fx_y = torch.empty_like(fx_x)
            for _ in range(3):
                fx_chain_forward(fx_x, fx_p, out=fx_y)
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for _ in range(10):
                fx_chain_forward(fx_x, fx_p, out=fx_y)
            f1.record()
            torch.cuda.synchronize()
            fx_ms = f0.elapsed_time(f1) / 10
            fx_gbs = 16.0 * fb * L / (fx_ms * 1e-3) / 1e9
            extra["fx_chain_config3"] = {
                "workload": "configs[2]: FX chain (EQ+comp+imager+gain), batch=256 random-param segments of 262144",
                "ms": fx_ms, "audio_s_per_s": fb * L / SR / (fx_ms * 1e-3),
                "roofline": {"bound": "hbm", "achieved": fx_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s",
                             "frac": fx_gbs / pk["hbm_gbs"], "algorithmic_bytes": 16.0 * fb * L,
                             "traffic": 3.06e9, "traffic_source": "profiles/r01f_fx_v2_2_ncu_full.csv: three passes "
                             "(two whole-segment RMS barriers), the EQ and compressor passes are instruction-bound"}}
            del fx_x, fx_y
        except Exception as exc:  # the headline line must survive a failure of the extra leg
            extra["fx_chain_config3"] = {"error": repr(exc)}

    # ---- extra: the sample-format kernels either side of the forward, HBM-bound streams, on rank 0 ----
    if rank == 0 and not args.no_extra:
        try:
            from music_mixing_style_transfer_b200 import wav_io
            n_fr = 2 * B * L                                     # two batches worth of stereo frames: every buffer set > the 126 MB L2
            gen = torch.Generator(device=device)
            gen.manual_seed(4321)
            pcm = torch.randint(-32768, 32768, (n_fr, 2), generator=gen, device=device, dtype=torch.int32).to(torch.int16)
            dec = torch.empty(2, n_fr, dtype=torch.float32, device=device)
            stems = (torch.randn(4, 2, n_fr // 2, generator=gen, device=device) * 0.3)
            for _ in range(3):
                wav_io.decode_pcm(pcm, device, out=dec)
                wav_io.encode_mix_pcm16(stems)
            g0, g1, g2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            g0.record()
            for _ in range(10):
                wav_io.decode_pcm(pcm, device, out=dec)
            g1.record()
            for _ in range(10):
                wav_io.encode_mix_pcm16(stems)
            g2.record()
            torch.cuda.synchronize()
            dec_ms, enc_ms = g0.elapsed_time(g1) / 10, g1.elapsed_time(g2) / 10
            dec_bytes = n_fr * (4 + 8)                           # int16 stereo in, fp32 planar out
            enc_bytes = (n_fr // 2) * (4 * 8 + 4)                # 4 fp32 stereo stems in, int16 stereo out
            extra["wav_io"] = {
                "workload": "8f-1: PCM16 decode of 2 batches of stereo frames (201 MB in + out); remix of 4 stems of one batch "
                            "+ PCM_16 quantise (302 MB); buffers exceed the L2",
                "decode": {"ms": dec_ms, "GB/s": dec_bytes / (dec_ms * 1e-3) / 1e9,
                           "frac_of_hbm_peak": dec_bytes / (dec_ms * 1e-3) / 1e9 / pk["hbm_gbs"]},
                "encode_mix": {"ms": enc_ms, "GB/s": enc_bytes / (enc_ms * 1e-3) / 1e9,
                               "frac_of_hbm_peak": enc_bytes / (enc_ms * 1e-3) / 1e9 / pk["hbm_gbs"]}}
            del pcm, dec, stems
        except Exception as exc:
            extra["wav_io"] = {"error": repr(exc)}
        torch.cuda.empty_cache()

    # ---- extra: file -> file through the inference entry (single process only: the entry would start its own group) ----
    if rank == 0 and world == 1 and not args.no_extra:
        try:
            extra["file_to_file"] = file_to_file_leg(device, os.environ.get("TMPDIR", "/tmp"))
        except Exception as exc:
            extra["file_to_file"] = {"error": repr(exc)}

    cpu_baseline = None
    if rank == 0 and not args.no_cpu_baseline:
        cores = pick_cpu_threads()
        state = {}
        tc0 = time.perf_counter()
        secs = cpu_reference_step(state)
        dt = time.perf_counter() - tc0
        if dt < 8.0:  # fast host: take a second sample so the figure is not a cold-start artefact
            tc0 = time.perf_counter()
            secs = cpu_reference_step(state)
            dt = time.perf_counter() - tc0
        cpu_baseline = {"value": secs / dt, "unit": UNIT, "cores": cores, "host_cores": os.cpu_count(), "kind": "port",
                        "sample": f"1 reference + 1 input segment of {SEG_LEN} stereo samples through the oracle port of "
                                  f"the reference torch modules, {dt:.1f} s wall", "torch": torch.__version__}
        if world == 1 and not args.no_extra:
            # the same modules on this GPU through cuDNN, next to the CPU number (secondary baseline, not the target)
            try:
                extra["reference_modules_on_gpu"] = dict(
                    reference_modules_on_gpu(device, B, L, state.get("out"), (state["ref"], state["inp"])),
                    workload="oracle port of the reference torch modules on this GPU (cuDNN), configs[1] segments, batch 8",
                    note="what inference/style_transfer.py:29-32 runs when CUDA is available; rms_err_vs_cpu_forward on the "
                         "cpu_baseline sample (1 segment)")
            except Exception as exc:
                extra["reference_modules_on_gpu"] = {"error": repr(exc)}

    if rank == 0:
        n_enc = 27                                           # 24 conv + split + pool + batch reduction (ranks that encode)
        n_tcn = 15 * chunks                                  # film + block 0 + 13 tcgen05 launches per converted sub-batch
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None,
                "dtype": ("f16+2xe4m3 (TCN: fp16 main product + two e4m3 correction products on tcgen05, fp32 accumulate; encoder: "
                          "split-bf16 x3 on tcgen05, blocks 0-2 fp32)") if f8 else
                         "bf16x3 (split-bf16 operand pairs on tcgen05, fp32 accumulate; encoder blocks 0-2 fp32)",
                "data": "synthetic",
                "config": {"workload": "configs[1]: FXencoder+MixFXcloner full forward, batch=32 segments of 262144 "
                                       "stereo samples per GPU", "segment_length": L, "batch_per_gpu": B,
                           "global_batch": total, "reference_batch": B, "tcn_precision": args.precision,
                           "parallelism": (f"dp{world} (input segments sharded; reference batch sharded, 1 NCCL all-reduce of the "
                                           f"embedding sums; output all-gather in {chunks} sub-batches overlapping the TCN)"
                                           if shard_ref else f"dp{world} (segments sharded; 1 NCCL broadcast of the embedding + "
                                           f"all-gather of outputs in {chunks} sub-batches)") if world > 1 else "single GPU",
                           "l2": "inputs and activations (>= 67 MB per tensor, 4.3 GB per TCN activation) exceed the 126 MB L2",
                           "weights": "seeded random, reference state_dict layout"},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / args.steps,
                        "api": "pipeline.StyleTransferPipeline.submit / collect (copies on side streams, 2 buffer sets)"},
                "gpu_launches": args.steps * (n_enc + n_tcn),
                "gpu_launches_per_step": {"encoder (24 conv + split + pool + batch mean)": n_enc,
                                          f"tcn (film + block0 + 13 umma) x {chunks} sub-batch(es)": n_tcn},
                "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline, "extra": extra,
                "tflops_algorithmic": (TCN_FLOP_PER_SAMPLE * L * B + ENC_FLOP_PER_SEG * (ref_hi - ref_lo)) / (ms_step * 1e-3) / 1e12}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
