"""2+ ranks over NCCL (torchrun --nproc-per-node N tools/nccl_check.py): the shard layer, the host-to-host pipeline and the
inference entry against single-process results of the same kernels.
  A  ragged shards, rank-0 encode + broadcast, one all-gather                      -> bit-identical to the unsharded forward
  B  even shards, sharded reference batch (all-reduce), all-gather in 2 sub-batches -> TCN path bit-identical given the
     embedding, embedding within 1e-6 of the unsharded mean (different summation order)
  C  interpolation mode with per-row conditioning (BASELINE configs[4])
  D  pipeline.StyleTransferPipeline: 4 pipelined host-to-host steps == direct steps
  E  inference/style_transfer.py under torchrun on synthetic song files: rank 0's mixture == the rows converted in one process"""
import os, shutil, sys, wave
import numpy as np
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_helpers import models, state_dicts
from music_mixing_style_transfer_b200 import shard, wav_io
from music_mixing_style_transfer_b200.pipeline import StyleTransferPipeline
from oracle import weights as W

rank, ws, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
enc, tcn = models()
ok = {}
L, S = 32768, 4
with torch.no_grad():
    # ---- A ----
    total = 6 * ws + 1
    full = W.synthetic_audio(total, L, seed=11).cuda()
    ref_a, ref_b = W.synthetic_audio(4, L, seed=12).cuda(), W.synthetic_audio(2, L, seed=13).cuda()
    lo, hi = shard.shard_bounds(total, ws, rank)
    emb, out = shard.sharded_style_transfer(enc, tcn, ref_a if rank == 0 else None, full[lo:hi].contiguous(), total)
    ea = enc.embed_mean(ref_a)
    ok["A broadcast + ragged all-gather"] = torch.equal(emb, ea) and torch.equal(out, tcn(full, ea.unsqueeze(0)))
    # ---- B ----
    total = 8 * ws
    full = W.synthetic_audio(total, L, seed=14).cuda()
    lo, hi = shard.shard_bounds(total, ws, rank)
    emb, out = shard.sharded_style_transfer(enc, tcn, ref_a, full[lo:hi].contiguous(), total, shard_reference=True,
                                            n_reference=4, gather_chunks=2)
    ok["B all-reduce + chunked all-gather"] = bool((emb - ea).abs().max() <= 1e-6) and torch.equal(out, tcn(full, emb.unsqueeze(0)))
    # ---- C ----
    total = 6 * ws + 1
    full = W.synthetic_audio(total, L, seed=11).cuda()
    lo, hi = shard.shard_bounds(total, ws, rank)
    w = shard.interpolation_weights(total, S, device="cuda")
    embs, out2 = shard.sharded_interpolation(enc, tcn, ref_a if rank == 0 else None, ref_b if rank == 0 else None,
                                             full[lo:hi].contiguous(), total, w)
    eb = enc.embed_mean(ref_b)
    cond = w[:, None] * ea[None] + (1 - w[:, None]) * eb[None]
    ok["C interpolation"] = torch.equal(out2, tcn(full, cond))
    # ---- D ----
    total = 4 * ws
    pipe = StyleTransferPipeline(enc, tcn, torch.device("cuda", lr), total, shard_reference=True, n_reference=4, gather_chunks=2)
    ref_h = W.synthetic_audio(4, L, seed=12).pin_memory()
    good = True
    outs, wants = [], []
    for i in range(4):
        inp_h = W.synthetic_audio(total, L, seed=20 + i)
        lo, hi = shard.shard_bounds(total, ws, rank)
        mine = inp_h[lo:hi].contiguous().pin_memory()
        if i >= pipe.depth:
            outs.append(pipe.collect().clone())
        pipe.submit(ref_h, mine)
        _, direct = shard.sharded_style_transfer(enc, tcn, ref_a, mine.cuda(), total, shard_reference=True, n_reference=4)
        wants.append(direct[lo:hi].cpu())
    outs += [o.clone() for o in pipe.drain()]
    ok["D pipeline"] = len(outs) == 4 and all(torch.equal(o, w_) for o, w_ in zip(outs, wants))

# ---- E ----
from music_mixing_style_transfer_b200.inference import style_transfer as st
root = f"/tmp/mst_nccl_check"
insts = ["drums", "bass", "other", "vocals"]
seg = 16384
if rank == 0:
    shutil.rmtree(root, ignore_errors=True)
    esd, tsd = state_dicts()
    os.makedirs(root)
    torch.save({"model": {"module." + k: v for k, v in esd.items()}}, f"{root}/enc.pt")
    torch.save({"model": {"module." + k: v for k, v in tsd.items()}}, f"{root}/tcn.pt")
    for name, n in (("input", 5 * seg + 123), ("reference", 7 * seg + 5)):
        for i, inst in enumerate(insts):
            x = W.synthetic_audio(1, n, seed=600 + 10 * len(name) + i)[0].numpy()
            d = f"{root}/data/song0/separated/{name}"
            os.makedirs(d, exist_ok=True)
            with wave.open(f"{d}/{inst}.wav", "wb") as f:
                f.setnchannels(2); f.setsampwidth(2); f.setframerate(44100)
                f.writeframes(np.clip(np.rint(x.T * 32768.0), -32768, 32767).astype("<i2").tobytes())
dist.barrier()
st.main(["--target_dir", f"{root}/data/", "--output_dir", f"{root}/out/", "--ckpt_path_enc", f"{root}/enc.pt",
         "--ckpt_path_conv", f"{root}/tcn.pt", "--segment_length", str(seg), "--segment_length_ref", str(seg),
         "--normalize_input", "False", "--do_not_separate", "True"])
dist.barrier()
if rank == 0:
    with torch.no_grad():
        got = wav_io.read_wav_pcm(f"{root}/out/song0/mixture_output_notnormed.wav").astype(np.int32)
        stems_in = torch.stack([wav_io.load_wav_to_device(f"{root}/data/song0/separated/input/{i}.wav") for i in insts])
        stems_ref = torch.stack([wav_io.load_wav_to_device(f"{root}/data/song0/separated/reference/{i}.wav") for i in insts])
        T = stems_in.shape[-1]
        rows_in = st.cut_rows(stems_in, seg, T // seg + 1)
        rows_ref = st.cut_rows(stems_ref, seg, stems_ref.shape[-1] // seg + 1)
        n_r = rows_ref.shape[0] // 4
        embs = torch.stack([enc.embed_mean(rows_ref[i * n_r:(i + 1) * n_r].contiguous()) for i in range(4)])
        n_s = rows_in.shape[0] // 4
        cond = embs.repeat_interleave(n_s, dim=0)
        y = st.join_rows(tcn(rows_in.contiguous(), cond), 4, T).contiguous()
        want = wav_io.encode_mix_pcm16(y).cpu().numpy().astype(np.int32)
    ok["E entry under torchrun"] = got.shape == want.shape and int(np.abs(got - want).max()) <= 1     # all-reduced embedding sums: +-1 LSB
    shutil.rmtree(root, ignore_errors=True)
print(f"rank {rank}/{ws}: " + ", ".join(f"{k}: {'ok' if v else 'FAILED'}" for k, v in ok.items()), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if all(ok.values()) else 1)
