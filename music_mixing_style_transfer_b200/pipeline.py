"""Host-to-host style-transfer steps with the copies hidden behind compute.

`StyleTransferPipeline` is the public call for callers whose audio lives in (pinned) host memory: `submit(reference, input)`
enqueues one sharded step (shard.sharded_style_transfer) and returns at once; `collect()` hands back the oldest finished step's
waveforms in pinned host memory.  Three CUDA streams and two buffer sets: while step i computes, the inputs of step i + 1 are
copied in on the copy-in stream and the waveforms of step i - 1 are copied out on the copy-out stream, so per step PCIe moves
the same bytes as a plain `.to(device)` / `.cpu()` round trip but none of them on the compute stream's critical path.
The reference does this serially on the default stream (inference/style_transfer.py:145, 158, 162).
"""
from collections import deque

import torch

from . import shard


class StyleTransferPipeline:
    def __init__(self, encoder, converter, device, total_segments, depth=2, shard_reference=False, n_reference=None,
                 gather=True, gather_chunks=1, group=None):
        self.encoder, self.converter, self.device = encoder, converter, torch.device(device)
        self.total, self.depth = total_segments, depth
        self.kw = dict(gather=gather, group=group, shard_reference=shard_reference, n_reference=n_reference,
                       gather_chunks=gather_chunks)
        self.copy_in, self.copy_out = torch.cuda.Stream(self.device), torch.cuda.Stream(self.device)
        self.slots = [dict() for _ in range(depth)]
        self.pending = deque()
        self.n_submitted = 0
        self.rank, self.world = shard.world(group)

    def _buffer(self, slot, name, like, pinned=False):
        buf = slot.get(name)
        if buf is None or buf.shape != like.shape:
            buf = torch.empty(like.shape, dtype=like.dtype, pin_memory=True) if pinned \
                else torch.empty(like.shape, dtype=like.dtype, device=self.device)
            slot[name] = buf
        return buf

    def submit(self, reference_host, input_host):
        """reference_host: pinned [B_ref, 2, L_ref] (None on ranks that do not encode), input_host: pinned [B_local, 2, L]."""
        if len(self.pending) >= self.depth:
            raise RuntimeError("pipeline full: collect() a finished step first")
        slot = self.slots[self.n_submitted % self.depth]
        compute = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self.copy_in):
            # the slot's device inputs were last read by the step submitted `depth` steps ago
            if "computed" in slot:
                self.copy_in.wait_event(slot["computed"])
            ref = None
            if reference_host is not None:
                ref = self._buffer(slot, "ref", reference_host)
                ref.copy_(reference_host, non_blocking=True)
            x = self._buffer(slot, "inp", input_host)
            x.copy_(input_host, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self.copy_in)
        compute.wait_event(ready)
        flags = []
        with torch.no_grad():
            _, out = shard.sharded_style_transfer(self.encoder, self._converter(flags), ref, x, self.total, **self.kw)
        fired_dev = None
        if flags:
            # did any converter call of this step leave the f16f8 range?  Reduced on the compute stream now and copied out with
            # the waveforms, so that collect() never waits on the compute stream (which already holds the next step)
            # (the raw flags travel: float bits of the largest out-of-range activation, 0 = none; positive floats order like
            # their int32 bit patterns, so MAX over ranks is well defined and no device arithmetic is needed here)
            fired_dev = flags[0] if len(flags) == 1 else torch.cat(flags)
            if self.world > 1:      # the repeat in collect() runs collectives: every rank must take the same branch
                torch.distributed.all_reduce(fired_dev, op=torch.distributed.ReduceOp.MAX, group=self.kw["group"])
        slot["computed"] = torch.cuda.Event()
        slot["computed"].record(compute)
        n_local = x.shape[0]
        lo = shard.shard_bounds(self.total, self.world, self.rank)[0] if (self.world > 1 and out.shape[0] == self.total) else 0
        mine = out[lo:lo + n_local]
        with torch.cuda.stream(self.copy_out):
            self.copy_out.wait_event(slot["computed"])
            host = self._buffer(slot, "out", mine, pinned=True)
            host.copy_(mine, non_blocking=True)
            mine.record_stream(self.copy_out)
            fired_host = None
            if fired_dev is not None:
                fired_host = self._buffer(slot, "fired", fired_dev, pinned=True)
                fired_host.copy_(fired_dev, non_blocking=True)
                fired_dev.record_stream(self.copy_out)
            done = torch.cuda.Event()
            done.record(self.copy_out)
        self.pending.append((host, done, out, fired_host, (ref, x, lo, n_local)))
        self.n_submitted += 1

    def _converter(self, flags):
        """With TCNModel.precision == 'auto' the f16f8 range flag is NOT read back inside the step (that would serialise the
        host with the GPU and un-hide the copies): every converter call gets its own device flag, checked in collect()."""
        conv = self.converter
        if getattr(conv, "precision", None) != "auto":
            return conv

        def run(x, cond):
            flag = torch.empty(1, dtype=torch.int32, device=self.device)     # mst_tcn_forward clears it (cudaMemsetAsync)
            flags.append(flag)
            conv.precision = "f16f8"
            try:
                return conv(x, cond, range_flag=flag)
            finally:
                conv.precision = "auto"
        return run

    def collect(self):
        """Pinned host tensor with this rank's waveforms of the oldest submitted step (blocks until its copy-out is done).
        The buffer is reused by the step submitted `depth` submits later."""
        host, done, _, fired_host, (ref, x, lo, n_local) = self.pending.popleft()
        done.synchronize()
        if fired_host is not None and bool((fired_host != 0).any()):
            # an activation left the f16f8 operand range: repeat this step with fp32-range operands (its inputs are still in
            # their slot: at most depth - 1 later steps have been submitted)
            self.converter.precision = "bf16x3"
            try:
                with torch.no_grad():
                    _, out = shard.sharded_style_transfer(self.encoder, self.converter, ref, x, self.total, **self.kw)
            finally:
                self.converter.precision = "auto"
            host.copy_(out[lo:lo + n_local])
        return host

    def drain(self):
        outs = []
        while self.pending:
            outs.append(self.collect())
        return outs
