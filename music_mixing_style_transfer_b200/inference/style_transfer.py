"""
    Music mixing style transfer on the B200 engine -- drop-in for the reference's `inference/style_transfer.py`
    (class Mixing_Style_Transfer_Inference :27, inference :112, inference_interpolation :181, batchwise_segmentization
    :274, flags :346-381): same flags, same directory layout, same output files.

        "path_to_data_directory"/"song_name_#1"/<stem dir>/input/{drums,bass,other,vocals}.wav
        "path_to_data_directory"/"song_name_#1"/<stem dir>/reference/{...}.wav        (+ reference_B for --interpolation)
      <stem dir> = `--stem_level_directory_name` with `--do_not_separate True`, else `separated/<separation_model>`
      (data_loader/data_loader.py:555-556).

    What is computed is the reference's: every stem is cut into zero-padded segments (a FULL extra zero segment when the
    length is an exact multiple, :287-288), the reference stem's segments go through the FXencoder and are averaged, the
    input stem's segments go through the TCN conditioned on that embedding (interpolation: w A + (1 - w) B with w indexed
    by the reference's BATCH index, :250), segments are concatenated, cropped, the stems summed and written as PCM_16.

    How it is scheduled is not the reference's per-stem / per-batch loop.  A song is one job:
      1. all stems are decoded on the device from raw PCM (wav_io.py / csrc/pcm.cu); the next song's files are read into
         pinned memory by a loader thread while this one computes;
      2. the segments of ALL stems form one row table [n_stems * n_segments, 2, segment]; under torchrun the rows are
         sharded over the ranks -- reference rows and input rows alike, so no GPU idles behind rank 0 and the result does
         not depend on `--batch_size` being large enough to shard (the reference's default is 1);
      3. encoder: per-stem embedding sums of the local rows, ONE all-reduce of [n_refs, n_stems, 2048] floats;
      4. TCN: local rows in launches bounded by activation memory, per-row conditioning (FiLM broadcasts per row,
         network_utils.py:180-182); the f16f8 range flags of all launches are read back once, flagged launches repeat in
         bf16x3; ONE all-gather of the output rows;
      5. rank 0 crops, remixes and quantises on the device; the int16 mixture is copied back asynchronously and a writer
         thread puts it on disk while the next song computes.
    `--batch_size` therefore only carries its reference semantics (the interpolation weight index) -- not memory, not speed.

    Outside this engine (raise, never silently skipped): the demucs subprocess (:77-90; pass --do_not_separate True with
    stems already separated).  All four normalisation effects run through mixing_manipulator/data_normalization.py; the
    'compression' effect needs the third-party `aubio` onset detector on the host exactly like the reference, and raises
    with that message when the package is missing.  CUDA only.
"""
import os
import queue
import sys
import threading
import time
import wave
from glob import glob

import numpy as np
import torch

currentdir = os.path.dirname(os.path.realpath(__file__))
sys.path.append(os.path.dirname(os.path.dirname(currentdir)))
from music_mixing_style_transfer_b200 import shard, wav_io  # noqa: E402
from music_mixing_style_transfer_b200.inference._common import (checkpoint_state_dict, dump_arguments,  # noqa: E402
                                                                 segment_into_batches)
from music_mixing_style_transfer_b200.networks import FXencoder, TCNModel  # noqa: E402

# time rows (segments x samples) one TCN launch chain may hold: 2 activation buffers of 512 B per row -> 8.6 GB
MAX_ROWS_PER_LAUNCH = 1 << 23


def load_wav_segment(audio_path, start_point=None, duration=None, axis=1, sample_rate=44100):
    """Host decode with the reference loader's result (data_loader/loader_utils.py:47-70): float64 in [-1, 1), channels
    along `axis`.  Only the host-I/O mode and the tests use it; the engine decodes on the device."""
    pcm = wav_io.read_wav_pcm(audio_path, start_point, duration, sample_rate)          # [n, ch] int16 / int32
    x = pcm / float(2 ** (8 * pcm.dtype.itemsize - 1))
    if pcm.shape[1] == 1:
        return x[:, 0]
    return x if axis == 1 else np.ascontiguousarray(x.T)


def write_wav_pcm16(path, data, sample_rate):
    """float [n, 2] -> PCM_16 file like `sf.write(..., 'PCM_16')`: scale 2^15, round half to even, clip."""
    pcm = np.clip(np.rint(np.asarray(data, dtype=np.float64) * 32768.0), -32768, 32767).astype('<i2')
    _write_riff(path, pcm, sample_rate)


def _write_riff(path, pcm_i16, sample_rate):
    with wave.open(path, 'wb') as w:
        w.setnchannels(pcm_i16.shape[1])
        w.setsampwidth(2)
        w.setframerate(sample_rate)
        w.writeframes(pcm_i16.astype('<i2', copy=False).tobytes())


class Song_Dataset_Inference:
    """The songs under `target_dir` and where their stems live (data_loader/data_loader.py:545-603).  `raw(idx)` reads the
    files (host work, done by the loader thread), `to_device(raw)` decodes them: {role: float32 [n_stems, 2, T]}."""

    def __init__(self, args, normalizer=None):
        self.args = args
        self.instruments = args.instruments
        self.data_dir_paths = sorted(glob(f"{args.target_dir}*/"))
        # with --do_not_separate the reference drops the separation-model component (data_loader.py:555-556)
        self.stem_level_directory_name = args.stem_level_directory_name if args.do_not_separate \
            else os.path.join(args.stem_level_directory_name, args.separation_model)
        self.roles = [("input", args.input_file_name), ("reference", args.reference_file_name)]
        if args.interpolation:
            self.roles.append(("reference_B", args.reference_file_name_2interpolate))
        self.normalizer = normalizer
        self.device_io = bool(getattr(args, "device_io", True))

    def __len__(self):
        return len(self.data_dir_paths)

    def stem_path(self, idx, name, inst):
        return os.path.join(self.data_dir_paths[idx], self.stem_level_directory_name, name, inst + '.wav')

    def raw(self, idx):
        out = {"dir_name": os.path.dirname(self.data_dir_paths[idx])}
        for role, name in self.roles:
            stems = []
            for inst in self.instruments:
                pcm = wav_io.read_wav_pcm(self.stem_path(idx, name, inst), sample_rate=self.args.sample_rate)
                stems.append(wav_io.pin_pcm(pcm) if self.device_io else pcm)
            if len({s.shape[0] for s in stems}) != 1:
                # the reference's torch.stack(input_stems) (data_loader.py:598-600) raises here as well
                raise RuntimeError(f"stems of {out['dir_name']}/{name} differ in length: {[s.shape[0] for s in stems]}")
            out[role] = stems
        return out

    def to_device(self, raw, device):
        songs = {"dir_name": raw["dir_name"]}
        for role, _ in self.roles:
            T = raw[role][0].shape[0]
            stems = torch.empty(len(self.instruments), 2, T, dtype=torch.float32, device=device)
            for i, pcm in enumerate(raw[role]):
                if self.device_io:
                    wav_io.decode_pcm(pcm, device, out=stems[i])
                else:   # host numpy path, same values: x / 2^15, de-interleave, clamp (data_loader.py:589-590)
                    x = pcm / float(2 ** (8 * pcm.dtype.itemsize - 1))
                    x = np.repeat(x, 2, axis=1) if x.shape[1] == 1 else x
                    stems[i].copy_(torch.from_numpy(np.clip(x.T, -1.0, 1.0)).float())
            if role == "input" and self.normalizer is not None:
                # FX normalisation of the INPUT stems only (data_loader.py:586-587), then the same clamp
                stems = torch.stack([self.normalizer.normalize_audio(stems[i], src=inst)
                                     for i, inst in enumerate(self.instruments)], dim=0).clamp_(-1.0, 1.0)
            songs[role] = stems
        return songs


class _Prefetcher:
    """Reads song idx + 1 from disk into pinned memory while song idx is on the GPU."""

    def __init__(self, dataset, device):
        self.dataset, self.device, self.q = dataset, device, queue.Queue(maxsize=1)
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        torch.cuda.set_device(self.device)      # pinned allocations of this thread belong to this rank's context
        for idx in range(len(self.dataset)):
            try:
                self.q.put(("ok", self.dataset.raw(idx)))
            except BaseException as exc:  # surfaced in the consumer, in order
                self.q.put(("error", exc))
                return
        self.q.put(("done", None))

    def __iter__(self):
        while True:
            kind, item = self.q.get()
            if kind == "done":
                return
            if kind == "error":
                raise item
            yield item


class _Writer:
    """Writes finished PCM to disk off the critical path: (path, pinned int16 tensor, CUDA event of its D2H copy)."""

    def __init__(self, sample_rate):
        self.sample_rate, self.q, self.error = sample_rate, queue.Queue(), None
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        while True:
            job = self.q.get()
            if job is None:
                return
            path, host, event = job
            try:
                event.synchronize()
                _write_riff(path, host.numpy(), self.sample_rate)
            except BaseException as exc:
                self.error = exc

    def submit(self, path, pcm_dev):
        host = torch.empty(pcm_dev.shape, dtype=torch.int16, pin_memory=pcm_dev.numel() > 0)
        host.copy_(pcm_dev, non_blocking=True)
        event = torch.cuda.Event()
        event.record()
        self.q.put((path, host, event))

    def close(self):
        self.q.put(None)
        self.thread.join()
        if self.error is not None:
            raise self.error


def cut_rows(stems, seg_len, n_seg):
    """[n_stems, 2, T] -> rows [n_stems * n_seg, 2, seg_len] (stem-major), zero-padded to n_seg * seg_len; one copy."""
    n_stems, _, T = stems.shape
    if n_seg * seg_len != T:
        stems = torch.cat((stems, stems.new_zeros(n_stems, 2, n_seg * seg_len - T)), dim=-1)
    return stems.view(n_stems, 2, n_seg, seg_len).permute(0, 2, 1, 3).reshape(n_stems * n_seg, 2, seg_len)


def join_rows(rows, n_stems, T):
    """inverse of cut_rows: rows [n_stems * n_seg, 2, seg] -> [n_stems, 2, T] (concatenate on time, crop; :165-169)."""
    n_seg, seg = rows.shape[0] // n_stems, rows.shape[-1]
    return rows.view(n_stems, n_seg, 2, seg).permute(0, 2, 1, 3).reshape(n_stems, 2, n_seg * seg)[..., :T]


class Mixing_Style_Transfer_Inference:
    def __init__(self, args, trained_w_ddp=True):
        if not torch.cuda.is_available() or args.inference_device == 'cpu':
            raise RuntimeError("Mixing_Style_Transfer_Inference (B200 engine) needs a CUDA device: there is no CPU path "
                               "(the reference's own CPU forward is only used as the parity oracle)")
        if not args.do_not_separate:
            raise NotImplementedError("source separation runs the external `demucs` CLI (style_transfer.py:77-90), which is "
                                      "outside this engine: separate the stems first and pass --do_not_separate True")
        self.rank, self.world_size = shard.world()
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.args = args
        self.device_io = bool(getattr(args, "device_io", True))
        self.segment_length = args.segment_length
        self.batch_size = args.batch_size
        self.sample_rate = 44100    # sampling rate should be 44100
        self.time_in_seconds = int(args.segment_length // self.sample_rate)
        self.output_dir = args.target_dir if args.output_dir is None else args.output_dir
        self.target_dir = args.target_dir

        conv = args.cfg_converter
        self.models = {
            'effects_encoder': FXencoder(args.cfg_encoder).to(self.device),
            'mixing_converter': TCNModel(nparams=conv["condition_dimension"], ninputs=2, noutputs=2,
                                         nblocks=conv["nblocks"], dilation_growth=conv["dilation_growth"],
                                         kernel_size=conv["kernel_size"], channel_width=conv["channel_width"],
                                         stack_size=conv["stack_size"], cond_dim=conv["condition_dimension"],
                                         causal=conv["causal"]).to(self.device)}
        self.reload_weights({'effects_encoder': args.ckpt_path_enc, 'mixing_converter': args.ckpt_path_conv},
                            ddp=trained_w_ddp)

        normalizer = None
        if args.normalize_input:
            from music_mixing_style_transfer_b200.mixing_manipulator.data_normalization import Audio_Effects_Normalizer
            normalizer = Audio_Effects_Normalizer(precomputed_feature_path=args.precomputed_normalization_feature,
                                                  STEMS=args.instruments, EFFECTS=args.normalization_order)
        self.data_loader = Song_Dataset_Inference(args, normalizer)
        self.stats = {"audio_seconds": 0.0, "wall_seconds": 0.0, "songs": 0}
        if self.rank == 0:
            self.save_args(args)

    def reload_weights(self, ckpt_paths, ddp=True):
        """`{"model": state_dict}` checkpoints, `module.` prefix of DDP training stripped (:94-108)."""
        for name, model in self.models.items():
            model.load_state_dict(checkpoint_state_dict(ckpt_paths[name], self.device, ddp=ddp))
            print(f"---reloaded checkpoint weights : {name} ---")

    # ---- planning: how a stem of T samples is cut (pure host logic, tests/test_host_logic.py) ----
    def plan_cut(self, T, segment_length, cut_above, song_name="<song>"):
        """(segment length, segment count).  Stems longer than `cut_above` samples are cut into `segment_length` pieces with
        the zero-padded tail of batchwise_segmentization (:281-301, which asserts T >= args.segment_length); shorter ones
        go through as one segment of their own length (:131-132, :139-140)."""
        if T <= cut_above:
            return T, 1
        assert T >= self.args.segment_length, (
            "Error : Insufficient duration!\n\t Target song's length is shorter than segment length.\n\t "
            f"Song name : {song_name}\n\t Consider changing the 'segment_length' or song with sufficient duration")
        return segment_length, T // segment_length + 1          # pad = seg - T % seg in (0, seg]  ->  floor(T / seg) + 1

    # ---- device stages ----
    def _local(self, n_rows):
        return shard.shard_bounds(n_rows, self.world_size, self.rank)

    def embed_stems(self, ref_sets):
        """ref_sets: list of (stems [n_stems, 2, T], seg_len, n_seg) -> [len(ref_sets), n_stems, 2048]: per stem the mean
        FXencoder embedding over all its segments (:144-153), rows sharded over the ranks, one all-reduce of the sums."""
        encoder = self.models["effects_encoder"].eval()
        n_stems = ref_sets[0][0].shape[0]
        sums = torch.zeros(len(ref_sets), n_stems, encoder.config["channels"][-1], device=self.device)
        for k, (stems, seg_len, n_seg) in enumerate(ref_sets):
            lo, hi = self._local(n_stems * n_seg)
            if hi == lo:
                continue
            rows = cut_rows(stems, seg_len, n_seg)[lo:hi]
            step = max(1, MAX_ROWS_PER_LAUNCH // seg_len)
            for s in range(0, hi - lo, step):
                emb = encoder(rows[s:s + step].contiguous())
                first = lo + s                      # global row of emb[0]; rows are stem-major
                for st in range(first // n_seg, (first + emb.shape[0] - 1) // n_seg + 1):
                    a, b = max(st * n_seg, first) - first, min((st + 1) * n_seg, first + emb.shape[0]) - first
                    sums[k, st] += emb[a:b].sum(dim=0)      # per-stem partial sums in a fixed order: bit-reproducible
        if self.world_size > 1:
            torch.distributed.all_reduce(sums)
        counts = torch.tensor([n for _, _, n in ref_sets], dtype=torch.float32, device=self.device)
        return sums / counts.view(-1, 1, 1)

    def convert_rows(self, rows, cond):
        """rows [R, 2, seg] (all stems of the song), cond [R, 2048] -> converted rows [R, 2, seg] on every rank."""
        converter = self.models["mixing_converter"].eval()
        R, _, seg = rows.shape
        lo, hi = self._local(R)
        local = torch.empty(hi - lo, 2, seg, dtype=torch.float32, device=self.device)
        step = max(1, MAX_ROWS_PER_LAUNCH // seg)
        auto = converter.precision == "auto"
        launches = []
        if auto:
            converter.precision = "f16f8"
        try:
            for s in range(lo, hi, step):
                e = min(hi, s + step)
                flag = torch.zeros(1, dtype=torch.int32, device=self.device) if auto else None
                x, c = rows[s:e].contiguous(), cond[s:e].contiguous()
                converter(x, c, out=local[s - lo:e - lo], range_flag=flag)
                launches.append((x, c, s, e, flag))
            if auto and launches:
                # one read-back for all launches: did any activation leave the f16f8 operand range?  (TCNModel.forward)
                excess = torch.cat([l[4] for l in launches]).view(torch.float32).cpu()
                for (x, c, s, e, _), v in zip(launches, excess.tolist()):
                    if v > 0.0:
                        converter.rerun_bf16x3(x, c, local[s - lo:e - lo])
        finally:
            if auto:
                converter.precision = "auto"
        if self.world_size > 1:
            return shard.allgather_segments(local, shard.shard_counts(R, self.world_size))
        return local

    def _emit(self, writer, cur_out_dir, stems_out, tag):
        """Per-instrument files (--save_each_inst) and the remix `sum(inst_outputs)` as PCM_16 (:170-177); rank 0 only."""
        if self.rank != 0:
            return
        jobs = [(f"mixture_{tag}.wav", stems_out)]
        if self.args.save_each_inst:
            jobs = [(f"{name}_{tag}.wav", stems_out[i]) for i, name in enumerate(self.args.instruments)] + jobs
        for fname, y in jobs:
            path = os.path.join(cur_out_dir, fname)
            if self.device_io:
                writer.submit(path, wav_io.encode_mix_pcm16(y))
            else:
                y = y.cpu().numpy()
                write_wav_pcm16(path, (y if y.ndim == 2 else y.sum(axis=0, dtype=np.float32)).T, self.args.sample_rate)

    def _run(self, tag, plan_song):
        """Shared song loop.  plan_song(songs) -> (input cut, [reference cut, ...], cond_fn(embs, stem_of_row, seg_of_row))."""
        writer = _Writer(self.args.sample_rate)
        t0 = time.perf_counter()
        try:
            with torch.no_grad():
                for raw in _Prefetcher(self.data_loader, self.device):
                    songs = self.data_loader.to_device(raw, self.device)
                    dir_name = songs["dir_name"]
                    print(f"---inference file name : {dir_name}---")
                    cur_out_dir = dir_name.replace(self.target_dir, self.output_dir)
                    if self.rank == 0:
                        os.makedirs(cur_out_dir, exist_ok=True)
                    (seg, n_seg), ref_sets, cond_fn = plan_song(songs)
                    inp = songs["input"]
                    n_stems, _, T = inp.shape
                    embs = self.embed_stems(ref_sets)
                    row = torch.arange(n_stems * n_seg, device=self.device)
                    cond = cond_fn(embs, row // n_seg, row % n_seg)
                    out_rows = self.convert_rows(cut_rows(inp, seg, n_seg), cond)
                    self._emit(writer, cur_out_dir, join_rows(out_rows, n_stems, T).contiguous(), tag)
                    self.stats["audio_seconds"] += T / self.sample_rate
                    self.stats["songs"] += 1
        finally:
            writer.close()
        torch.cuda.synchronize()
        self.stats["wall_seconds"] += time.perf_counter() - t0
        return self.stats

    # Inference whole song
    def inference(self):
        print("\n======= Start to inference music mixing style transfer =======")
        a = self.args
        tag = 'output' if a.normalize_input else 'output_notnormed'

        def plan_song(songs):
            name = songs["dir_name"]
            cut_in = self.plan_cut(songs["input"].shape[-1], a.segment_length, a.segment_length, name)
            # the reference cuts the style reference at `segment_length_ref` but only when it is longer than TWICE
            # `segment_length` (:133-136, quirk q5)
            Tr = songs["reference"].shape[-1]
            ref_sets = [(songs["reference"], *self.plan_cut(Tr, a.segment_length_ref, 2 * a.segment_length, name))]
            return cut_in, ref_sets, lambda embs, stem, seg_idx: embs[0][stem]

        return self._run(tag, plan_song)

    # Inference whole song, interpolating between two reference styles
    def inference_interpolation(self):
        print("\n======= Start to inference interpolation examples =======")
        a = self.args
        tag = 'output_interpolation' if a.normalize_input else 'output_notnormed_interpolation'
        S = a.interpolate_segments

        def plan_song(songs):
            name = songs["dir_name"]
            T = songs["input"].shape[-1]
            # the input is always cut, into pieces of T // S + 1 samples (:196-200)
            cut_in = self.plan_cut(T, T // S + 1, -1, name)
            # A is cut at `segment_length_ref`, B at `segment_length`; both only when longer than `segment_length_ref`
            # (:203-216, quirk q5)
            ref_sets = [(songs["reference"], *self.plan_cut(songs["reference"].shape[-1], a.segment_length_ref,
                                                            a.segment_length_ref, name)),
                        (songs["reference_B"], *self.plan_cut(songs["reference_B"].shape[-1], a.segment_length,
                                                              a.segment_length_ref, name))]

            def cond_fn(embs, stem, seg_idx):
                # the weight follows the reference's BATCH index (:247-251, quirk q4): segments of one batch share it
                batch_idx = (seg_idx // self.batch_size).to(torch.float32)
                w = ((S - 1 - batch_idx) / (S - 1)).unsqueeze(1)
                return w * embs[0][stem] + (1 - w) * embs[1][stem]

            return cut_in, ref_sets, cond_fn

        return self._run(tag, plan_song)

    # function that segmentize an entire song into batch (kept for callers of the reference's method, :274-301)
    def batchwise_segmentization(self, target_song, song_name, segment_length, discard_last=False):
        return segment_into_batches(target_song, segment_length, self.args.batch_size, min_length=self.args.segment_length,
                                    name=song_name, discard_last=discard_last)

    # save current inference arguments
    def save_args(self, params):
        dump_arguments(params, f"{self.output_dir}style_transfer_inference_configurations.txt")


def build_parser():
    import argparse

    def str2bool(v):
        if v.lower() in ('yes', 'true', 't', 'y', '1'):
            return True
        elif v.lower() in ('no', 'false', 'f', 'n', '0'):
            return False
        else:
            raise argparse.ArgumentTypeError('Boolean value expected.')

    ''' Configurations for music mixing style transfer '''
    repo_root = os.path.dirname(os.path.dirname(currentdir))
    default_ckpt_path_enc = os.path.join(repo_root, 'weights', 'FXencoder_ps.pt')
    default_ckpt_path_conv = os.path.join(repo_root, 'weights', 'MixFXcloner_ps.pt')
    default_norm_feature_path = os.path.join(repo_root, 'weights', 'musdb18_fxfeatures_eqcompimagegain.npy')

    parser = argparse.ArgumentParser()

    directory_args = parser.add_argument_group('Directory args')
    # directory paths
    directory_args.add_argument('--target_dir', type=str, default='./samples/style_transfer/')
    directory_args.add_argument('--output_dir', type=str, default=None, help='if no output_dir is specified (None), the results will be saved inside the target_dir')
    directory_args.add_argument('--input_file_name', type=str, default='input')
    directory_args.add_argument('--reference_file_name', type=str, default='reference')
    directory_args.add_argument('--reference_file_name_2interpolate', type=str, default='reference_B')
    # saved weights
    directory_args.add_argument('--ckpt_path_enc', type=str, default=default_ckpt_path_enc)
    directory_args.add_argument('--ckpt_path_conv', type=str, default=default_ckpt_path_conv)
    directory_args.add_argument('--precomputed_normalization_feature', type=str, default=default_norm_feature_path)

    inference_args = parser.add_argument_group('Inference args')
    inference_args.add_argument('--sample_rate', type=int, default=44100)
    inference_args.add_argument('--segment_length', type=int, default=2**19)        # segmentize input according to this duration
    inference_args.add_argument('--segment_length_ref', type=int, default=2**19)    # segmentize reference according to this duration
    # stem-level instruments & separation
    # (declared with type=str2bool in the reference, :367: only the default is usable there; here the flag also takes the names)
    inference_args.add_argument('--instruments', type=str, nargs='+', default=["drums", "bass", "other", "vocals"], help='instrumental tracks to perform style transfer')
    inference_args.add_argument('--stem_level_directory_name', type=str, default='separated')
    inference_args.add_argument('--save_each_inst', type=str2bool, default=False)
    inference_args.add_argument('--do_not_separate', type=str2bool, default=False)
    inference_args.add_argument('--separation_model', type=str, default='mdx_extra')
    # FX normalization
    inference_args.add_argument('--normalize_input', type=str2bool, default=True)
    # Effects to be normalized, order matters.  The reference declares this flag with type=str2bool (:372), so there only the default
    # list is usable; here the flag also takes the effect names: --normalization_order loudness eq imager loudness
    inference_args.add_argument('--normalization_order', type=str, nargs='+',
                                default=['loudness', 'eq', 'compression', 'imager', 'loudness'])
    # interpolation
    inference_args.add_argument('--interpolation', type=str2bool, default=False)
    inference_args.add_argument('--interpolate_segments', type=int, default=30)

    device_args = parser.add_argument_group('Device args')
    device_args.add_argument('--workers', type=int, default=1)
    device_args.add_argument('--inference_device', type=str, default='gpu', help="the B200 engine only runs on CUDA devices")
    device_args.add_argument('--batch_size', type=int, default=1)   # for processing long audio
    device_args.add_argument('--separation_device', type=str, default='cpu', help="device for performing source separation using Demucs")
    device_args.add_argument('--device_io', type=str2bool, default=True, help="(B200 engine) decode PCM / remix / quantise on the GPU; False = host numpy path, same bits")
    return parser


def main(argv=None):
    import yaml

    parser = build_parser()
    args = parser.parse_args(argv)
    args._parser = parser

    # one process per GPU under torchrun; single process otherwise
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        if not dist.is_initialized():
            dist.init_process_group("nccl")

    # load network configurations
    with open(os.path.join(currentdir, 'configs.yaml'), 'r') as f:
        configs = yaml.full_load(f)
    args.cfg_encoder = configs['Effects_Encoder']['default']
    args.cfg_converter = configs['TCN']['default']

    # Perform music mixing style transfer
    engine = Mixing_Style_Transfer_Inference(args)
    return engine.inference_interpolation() if args.interpolation else engine.inference()


if __name__ == '__main__':
    main()
