"""Helpers shared by the two inference entries of the B200 engine (style_transfer.py, feature_extraction.py)."""
import os
from collections import OrderedDict

import numpy as np
import torch


def checkpoint_state_dict(path, device, ddp=True):
    """`{"model": state_dict}` checkpoint -> state_dict with the 7-character `module.` prefix of DDP training removed
    (reference: inference/style_transfer.py:94-108, inference/feature_extraction.py:51-64)."""
    checkpoint = torch.load(path, map_location=device)
    return OrderedDict((k[7:] if ddp else k, v) for k, v in checkpoint["model"].items())


def segment_into_batches(song, segment_length, batch_size, min_length=None, name="<tensor>", discard_last=False):
    """[2, T] -> list of [<= batch_size, 2, segment_length] batches, on the device `song` lives on.
    Reference semantics (inference/feature_extraction.py:114-140, style_transfer.py:274-301): AssertionError when the song
    is shorter than `min_length` (default: one segment); the tail is padded with `segment_length - T % segment_length`
    zeros, i.e. with a FULL extra all-zero segment when T is an exact multiple; `discard_last` drops the tail instead."""
    T = song.shape[-1]
    min_length = segment_length if min_length is None else min_length
    assert T >= min_length, ("Error : Insufficient duration!\n\t Target song's length is shorter than segment length.\n\t "
                             f"Song name : {name}\n\t Consider changing the 'segment_length' or song with sufficient duration")
    if discard_last:
        song = song[:, :T - T % segment_length]
    else:
        song = torch.cat((song, song.new_zeros(2, segment_length - T % segment_length)), dim=-1)
    segments = song.unfold(-1, segment_length, segment_length).permute(1, 0, 2)        # [n_seg, 2, segment_length] (view)
    return [segments[i:i + batch_size].contiguous() for i in range(0, segments.shape[0], batch_size)]


def dump_arguments(params, record_path):
    """The `[args]` dump both reference entries write next to their outputs (style_transfer.py:304-322)."""
    parser = getattr(params, "_parser", None)
    lines = ["", "[args]"]
    for group in (parser._action_groups if parser is not None else []):
        if group.title in ("positional arguments", "optional arguments", "options"):
            continue
        lines.append(f"  {group.title} ({len(group._group_actions)})")
        lines += [f"      - {act.dest:20s}: {getattr(params, act.dest)}" for act in group._group_actions]
    os.makedirs(os.path.dirname(record_path) or ".", exist_ok=True)
    with open(record_path, "w") as f:
        np.savetxt(f, ["\n".join(lines) + "\n\n"], delimiter=" ", fmt="%s")
