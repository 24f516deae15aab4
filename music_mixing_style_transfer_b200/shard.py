"""Multi-GPU shard layer for the segment-batched forward (one process per GPU, torch.distributed over NCCL/NVLink).

Segments are independent through the TCN (zero padding per segment, BatchNorm in eval mode, no cross-batch op --
reference architectures.py:222-234), so the path partitions over the batch with NO data-path collective inside the
network.  The only exchanges are the ones BASELINE's north_star names: one broadcast of the 2048-float reference
embedding and one all-gather of the output segments.  The reference itself is single-device
(inference/style_transfer.py:29-32,327); this layer is the B200 addition inside its orchestration level.

Works with any initialised process group: `nccl` on the GPU box, `gloo` in the CPU tests of the index logic.
"""
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def world(group=None) -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_bounds(n_items: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous partition of range(n_items); the first n_items % world_size ranks get one extra item."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    base, rem = divmod(n_items, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_counts(n_items: int, world_size: int) -> List[int]:
    return [shard_bounds(n_items, world_size, r)[1] - shard_bounds(n_items, world_size, r)[0] for r in range(world_size)]


def broadcast_embedding(emb: Optional[torch.Tensor], shape, device, src: int = 0, group=None) -> torch.Tensor:
    """Rank `src` holds `emb` ([2048] or [n, 2048]); every rank returns it.  8-16 KiB, latency bound."""
    rank, ws = world(group)
    if ws == 1:
        return emb
    if rank != src:
        emb = torch.empty(shape, dtype=torch.float32, device=device)
    else:
        emb = emb.contiguous()
    dist.broadcast(emb, src=src, group=group)
    return emb


def allgather_segments(local: torch.Tensor, counts: List[int], group=None) -> torch.Tensor:
    """local: this rank's [counts[rank], C, L] output segments -> [sum(counts), C, L] on every rank, in rank order.
    Even shards go through one all_gather_into_tensor (a single NCCL all-gather over NVLink); ragged shards are padded
    to the largest count and trimmed."""
    rank, ws = world(group)
    if ws == 1:
        return local
    if local.shape[0] != counts[rank]:
        raise RuntimeError(f"rank {rank}: local shard has {local.shape[0]} segments, expected {counts[rank]}")
    mx = max(counts)
    tail = tuple(local.shape[1:])
    if mx == 0:
        return local.new_empty((0,) + tail)
    if local.shape[0] < mx:
        pad = local.new_zeros((mx - local.shape[0],) + tail)
        local = torch.cat([local, pad], dim=0)
    local = local.contiguous()
    out = local.new_empty((ws * mx,) + tail)
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(out, local, group=group)
    else:  # gloo (CPU tests)
        parts = [local.new_empty(local.shape) for _ in range(ws)]
        dist.all_gather(parts, local, group=group)
        out = torch.cat(parts, dim=0)
    if all(c == mx for c in counts):
        return out
    return torch.cat([out[r * mx:r * mx + counts[r]] for r in range(ws)], dim=0)


def _embed(encoder, batch, mean=True):
    """Mean (or sum) of the batch's embeddings: the encoder's fused reduction when it has one (FXencoder.embed_mean)."""
    if hasattr(encoder, "embed_mean"):
        return encoder.embed_mean(batch, None if mean else 1.0)
    emb = encoder(batch)
    return emb.mean(dim=0) if mean else emb.sum(dim=0)


def mean_embedding(encoder, reference_batch: Optional[torch.Tensor], n_reference: int, device, cond_dim: int = 2048,
                   shard_reference: bool = False, group=None) -> torch.Tensor:
    """Mean FXencoder embedding of the reference batch on every rank (inference/style_transfer.py:144-153).
      shard_reference = False: rank 0 holds the batch, encodes it and the [cond_dim] mean is BROADCAST (the exchange
                               BASELINE's north_star names); the other ranks wait for it.
      shard_reference = True:  every rank holds the batch (or at least its own contiguous slice of it, passed as the slice),
                               encodes n_reference / world segments and the partial sums are ALL-REDUCED (8 KiB) -- no rank
                               idles behind rank 0's encoder pass."""
    rank, ws = world(group)
    if ws == 1:
        return _embed(encoder, reference_batch)
    if not shard_reference:
        emb = _embed(encoder, reference_batch) if rank == 0 else None
        return broadcast_embedding(emb, (cond_dim,), device, src=0, group=group)
    lo, hi = shard_bounds(n_reference, ws, rank)
    part = torch.zeros(cond_dim, dtype=torch.float32, device=device)
    if hi > lo:
        mine = reference_batch if reference_batch.shape[0] == hi - lo else reference_batch[lo:hi]
        part = _embed(encoder, mine.contiguous(), mean=False)
    dist.all_reduce(part, group=group)
    return part / float(n_reference)


def convert_and_gather(converter, input_shard: torch.Tensor, cond: torch.Tensor, total_segments: int, gather: bool = True,
                       chunks: int = 1, group=None) -> torch.Tensor:
    """TCN over this rank's shard and the all-gather of the output segments, overlapped: the shard is converted in `chunks`
    sub-batches and the NCCL all-gather of sub-batch k runs (on NCCL's own stream) while sub-batch k + 1 is being converted,
    so only the last sub-batch's gather is exposed.  cond: [1, cond_dim] or one row per local segment.
    Returns [total_segments, 2, L] in rank order (or the local shard when gather is False / single rank)."""
    rank, ws = world(group)
    n_local = input_shard.shape[0]
    counts = shard_counts(total_segments, ws)
    even = all(c == counts[0] for c in counts)
    overlapped = gather and ws > 1 and chunks > 1 and even and n_local % chunks == 0 and dist.get_backend(group) == "nccl"
    if not overlapped:
        out = converter(input_shard, cond)
        return allgather_segments(out, counts, group=group) if (gather and ws > 1) else out
    sb = n_local // chunks
    tail = tuple(input_shard.shape[1:-1]) + (input_shard.shape[-1],)
    gathered = input_shard.new_empty((chunks, ws, sb) + tail)          # [k][rank][row of sub-batch k]
    works, keep = [], []
    for k in range(chunks):
        c = cond if cond.shape[0] == 1 else cond[k * sb:(k + 1) * sb].contiguous()
        y = converter(input_shard[k * sb:(k + 1) * sb].contiguous(), c)
        keep.append(y)                                   # alive until its gather has completed
        works.append(dist.all_gather_into_tensor(gathered[k].view((ws * sb,) + tail), y, group=group, async_op=True))
    for w in works:
        w.wait()
    # rank-major order: out[r * n_local + k * sb + i] = gathered[k][r][i]   (one device copy of the gathered batch)
    return gathered.permute(1, 0, 2, *range(3, gathered.dim())).reshape((total_segments,) + tail)


def sharded_style_transfer(encoder, converter, reference_batch: Optional[torch.Tensor], input_shard: torch.Tensor,
                           total_segments: int, cond_dim: int = 2048, gather: bool = True, group=None,
                           shard_reference: bool = False, n_reference: Optional[int] = None, gather_chunks: int = 1):
    """One sharded forward step.
      reference_batch: [B_ref, 2, L_ref] on rank 0 (ignored elsewhere)  -> encoder -> mean embedding -> broadcast;
                       with shard_reference=True every rank passes the batch (or its slice) and the sums are all-reduced
      input_shard:     this rank's contiguous [B_local, 2, L] slice of the `total_segments` input segments
    Returns (embedding [cond_dim], output) where output is the gathered [total_segments, 2, L] (gather=True) or the
    local shard.  `encoder` / `converter` are callables with the FXencoder / TCNModel forward signatures."""
    if n_reference is None:
        n_reference = 0 if reference_batch is None else int(reference_batch.shape[0])
    emb = mean_embedding(encoder, reference_batch, n_reference, input_shard.device, cond_dim, shard_reference, group)
    out = convert_and_gather(converter, input_shard, emb.unsqueeze(0), total_segments, gather, gather_chunks, group)  # :161
    return emb, out


def interpolation_weights(total_segments: int, interpolate_segments: int, device=None) -> torch.Tensor:
    """Per-segment weight of embedding A for BASELINE config 5: the batch holds songs of `interpolate_segments` consecutive
    segments, row b gets w = (S - 1 - (b mod S)) / (S - 1)  (inference/style_transfer.py:250 with the index taken per
    SEGMENT -- the reference indexes it per batch, which is the same thing at its default batch_size = 1)."""
    S = int(interpolate_segments)
    if S < 2:
        raise ValueError("interpolate_segments must be >= 2")
    idx = torch.arange(total_segments, device=device) % S
    return (S - 1 - idx).to(torch.float32) / float(S - 1)


def sharded_interpolation(encoder, converter, reference_a: Optional[torch.Tensor], reference_b: Optional[torch.Tensor],
                          input_shard: torch.Tensor, total_segments: int, weights: torch.Tensor, cond_dim: int = 2048,
                          gather: bool = True, group=None, gather_chunks: int = 1):
    """Interpolation mode (inference/style_transfer.py:181-270) sharded over the ranks.
      reference_a / reference_b: [B_ref, 2, L_ref] on rank 0 -> two mean embeddings -> ONE broadcast of [2, cond_dim]
      input_shard: this rank's contiguous slice of the `total_segments` input segments
      weights:     [total_segments] weight of embedding A per segment (interpolation_weights); every rank holds all of it
    Each rank builds the per-segment conditioning rows of its own slice, cond = w A + (1 - w) B (:251), and runs the
    converter with cond [B_local, cond_dim] (FiLM broadcasts per row, network_utils.py:180-182)."""
    rank, ws = world(group)
    embs = None
    if rank == 0:
        embs = torch.stack([_embed(encoder, reference_a),
                            _embed(encoder, reference_b)], dim=0)
    embs = broadcast_embedding(embs, (2, cond_dim), input_shard.device, src=0, group=group)
    lo, hi = shard_bounds(total_segments, ws, rank)
    if hi - lo != input_shard.shape[0]:
        raise RuntimeError(f"rank {rank}: input shard has {input_shard.shape[0]} segments, expected {hi - lo}")
    w = weights[lo:hi].to(device=input_shard.device, dtype=torch.float32).unsqueeze(1)
    cond = w * embs[0].unsqueeze(0) + (1.0 - w) * embs[1].unsqueeze(0)
    out = convert_and_gather(converter, input_shard, cond, total_segments, gather, gather_chunks, group)
    return embs, out
