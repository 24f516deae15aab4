"""2+ ranks over NCCL: shard.sharded_style_transfer and shard.sharded_interpolation (BASELINE configs 4 / 5 host logic) against
the single-process result of the same kernels.  torchrun --nproc-per-node N tools/nccl_interp_check.py"""
import os, sys
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_helpers import models
from music_mixing_style_transfer_b200 import shard
from oracle import weights as W

rank, ws, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
enc, tcn = models()
L, total, S = 32768, 6 * ws + 1, 4                       # ragged shards on purpose
full = W.synthetic_audio(total, L, seed=11).cuda()
ref_a, ref_b = W.synthetic_audio(3, L, seed=12).cuda(), W.synthetic_audio(2, L, seed=13).cuda()
lo, hi = shard.shard_bounds(total, ws, rank)
with torch.no_grad():
    emb, out = shard.sharded_style_transfer(enc, tcn, ref_a if rank == 0 else None, full[lo:hi].contiguous(), total)
    want = tcn(full, enc(ref_a).mean(0).unsqueeze(0))
    ok1 = torch.equal(out, want)
    w = shard.interpolation_weights(total, S, device="cuda")
    embs, out2 = shard.sharded_interpolation(enc, tcn, ref_a if rank == 0 else None, ref_b if rank == 0 else None,
                                             full[lo:hi].contiguous(), total, w)
    ea, eb = enc(ref_a).mean(0), enc(ref_b).mean(0)
    cond = w[:, None] * ea[None] + (1 - w[:, None]) * eb[None]
    want2 = tcn(full, cond)
    ok2 = torch.equal(out2, want2)
print(f"rank {rank}/{ws}: style_transfer bit-identical {ok1}, interpolation bit-identical {ok2}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if (ok1 and ok2) else 1)
