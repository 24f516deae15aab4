"""
    Inference code of music mixing style transfer on the B200 engine -- the entry that mirrors the reference's
    `inference/style_transfer.py` (class Mixing_Style_Transfer_Inference :27, inference :112, inference_interpolation
    :181, batchwise_segmentization :274, flags :346-381).

    Process : converts the mixing style of the input music recording to that of the refernce music.
                files inside the target directory should be organized as follow
                    "path_to_data_directory"/"song_name_#1"/input.wav
                    "path_to_data_directory"/"song_name_#1"/reference.wav
                    ...
                where the 'input' and 'reference' should share the same names.

    What is the same: flags, directory layout, segmentation (incl. the extra zero segment when the length is an exact
    multiple, :287-288), per-stem encoder -> mean embedding -> TCN -> concat -> crop -> 4-stem sum -> PCM_16 files,
    interpolation weights indexed by batch (:250).
    What differs, on purpose: (a) the networks run on libmst_b200.so (CUDA only, no CPU mode); (b) the embedding is the
    mean over ALL reference segments also when the last batch is short, where the reference's torch.stack would throw
    (:152); (c) under torchrun (WORLD_SIZE > 1) input batches are sharded over the ranks (shard.py); (d) the CPU
    "FX normalisation" pre-step (--normalize_input, data_loader.py:586-587) and the demucs subprocess (:77-90) are outside
    this engine: pass --normalize_input False --do_not_separate True with stems already separated; (e) the host I/O either
    side of the forward runs on the device (wav_io.py / csrc/pcm.cu): raw int16 PCM is copied to the GPU and
    converted / de-interleaved / clamped there, segments are cut from the device-resident stem, the converted stems stay
    on the device, and remix + PCM_16 quantisation produce ONE int16 mixture that is copied back and written -- bit-identical
    to the host path (`--device_io False` keeps that path).
"""
import os
import sys
import wave
from glob import glob

import numpy as np
import torch

currentdir = os.path.dirname(os.path.realpath(__file__))
sys.path.append(os.path.dirname(os.path.dirname(currentdir)))
from music_mixing_style_transfer_b200.networks import FXencoder, TCNModel  # noqa: E402
from music_mixing_style_transfer_b200 import shard, wav_io  # noqa: E402


# ---- WAV I/O with the reference loader's semantics (mixing_style_transfer/data_loader/loader_utils.py:47-70) ----
def load_wav_segment(audio_path, start_point=None, duration=None, axis=1, sample_rate=44100):
    start_point = 0 if start_point is None else start_point
    pt_wav = wave.open(audio_path, 'r')
    duration = pt_wav.getnframes() if duration is None else duration
    if pt_wav.getframerate() != sample_rate:
        raise ValueError(f"ValueError: input audio's sample rate should be {sample_rate}")
    pt_wav.setpos(start_point)
    x = pt_wav.readframes(duration)
    if pt_wav.getsampwidth() == 2:
        x = np.frombuffer(x, dtype=np.int16)
        X = x / float(2 ** 15)    # needs to be 16 bit format
    elif pt_wav.getsampwidth() == 4:
        x = np.frombuffer(x, dtype=np.int32)
        X = x / float(2 ** 31)    # needs to be 32 bit format
    else:
        raise ValueError("ValueError: input audio's bit depth should be 16 or 32-bit")
    # exception for stereo channels
    if pt_wav.getnchannels() == 2:
        X_l = np.expand_dims(X[::2], axis=axis)
        X_r = np.expand_dims(X[1::2], axis=axis)
        X = np.concatenate((X_l, X_r), axis=axis)
    return X


def write_wav_pcm16(path, data, sample_rate):
    """data: float [n, 2]; PCM_16 like `sf.write(..., 'PCM_16')` (scale 2^15, round, clip)."""
    pcm = np.clip(np.rint(np.asarray(data, dtype=np.float64) * 32768.0), -32768, 32767).astype('<i2')
    with wave.open(path, 'wb') as w:
        w.setnchannels(pcm.shape[1])
        w.setsampwidth(2)
        w.setframerate(sample_rate)
        w.writeframes(pcm.tobytes())


class Song_Dataset_Inference:
    """Stems of every song directory (mixing_style_transfer/data_loader/data_loader.py:545-603), without the CPU FX
    normaliser.  Items: (input_stems [4,2,T], reference_stems [4,2,T'][, reference_B], dir_name)."""

    def __init__(self, args):
        self.args = args
        self.data_dir = args.target_dir
        self.interpolate = args.interpolation
        self.instruments = args.instruments
        self.data_dir_paths = sorted(glob(f"{self.data_dir}*/"))
        self.input_name = args.input_file_name
        self.reference_name = args.reference_file_name
        self.stem_level_directory_name = args.stem_level_directory_name
        if args.normalize_input:
            raise NotImplementedError(
                "--normalize_input True: the CPU FX-normalisation pre-step (mixing_manipulator/data_normalization.py) is "
                "outside this engine's hot path (SURVEY.md 8f-2); run with --normalize_input False")

    def __len__(self):
        return len(self.data_dir_paths)

    def load_stems(self, dir_path, name):
        stems = []
        for inst in self.instruments:
            p = os.path.join(dir_path, self.stem_level_directory_name, self.args.separation_model, name, inst + '.wav')
            if getattr(self.args, "device_io", True):
                # raw PCM -> GPU, int -> float / de-interleave / clamp there (csrc/pcm.cu); same values as the host path
                stems.append(wav_io.load_wav_to_device(p, sample_rate=self.args.sample_rate))
            else:
                x = load_wav_segment(p, axis=0, sample_rate=self.args.sample_rate)       # [2, T]
                stems.append(torch.from_numpy(np.clip(x, -1.0, 1.0)).float())             # data_loader.py:589-590
        return torch.stack(stems, dim=0)

    def __getitem__(self, idx):
        d = self.data_dir_paths[idx]
        items = [self.load_stems(d, self.input_name), self.load_stems(d, self.reference_name)]
        if self.interpolate:
            items.append(self.load_stems(d, self.args.reference_file_name_2interpolate))
        return (*items, d)

    def __iter__(self):
        for i in range(len(self)):
            item = self[i]
            # batch_size=1 DataLoader collation of the reference: leading batch dim, dir name in a list
            yield tuple(t.unsqueeze(0) for t in item[:-1]) + ([item[-1]],)


class Mixing_Style_Transfer_Inference:
    def __init__(self, args, trained_w_ddp=True):
        if not torch.cuda.is_available() or args.inference_device == 'cpu':
            raise RuntimeError("Mixing_Style_Transfer_Inference (B200 engine) needs a CUDA device: there is no CPU path "
                               "(the reference's own CPU forward is only used as the parity oracle)")
        self.rank, self.world_size = shard.world()
        self.device = torch.device("cuda", torch.cuda.current_device())

        # inference computational hyperparameters
        self.args = args
        self.device_io = bool(getattr(args, "device_io", True))
        self.segment_length = args.segment_length
        self.batch_size = args.batch_size
        self.sample_rate = 44100    # sampling rate should be 44100
        self.time_in_seconds = int(args.segment_length // self.sample_rate)

        # directory configuration
        self.output_dir = args.target_dir if args.output_dir is None else args.output_dir
        self.target_dir = args.target_dir

        # load model and its checkpoint weights
        self.models = {}
        self.models['effects_encoder'] = FXencoder(args.cfg_encoder).to(self.device)
        self.models['mixing_converter'] = TCNModel(nparams=args.cfg_converter["condition_dimension"],
                                                   ninputs=2,
                                                   noutputs=2,
                                                   nblocks=args.cfg_converter["nblocks"],
                                                   dilation_growth=args.cfg_converter["dilation_growth"],
                                                   kernel_size=args.cfg_converter["kernel_size"],
                                                   channel_width=args.cfg_converter["channel_width"],
                                                   stack_size=args.cfg_converter["stack_size"],
                                                   cond_dim=args.cfg_converter["condition_dimension"],
                                                   causal=args.cfg_converter["causal"]).to(self.device)

        ckpt_paths = {'effects_encoder': args.ckpt_path_enc,
                      'mixing_converter': args.ckpt_path_conv}
        # reload saved model weights
        self.reload_weights(ckpt_paths, ddp=trained_w_ddp)

        # load data loader for the inference procedure
        self.data_loader = Song_Dataset_Inference(args)

        # save current arguments
        if self.rank == 0:
            self.save_args(args)
        if not self.args.do_not_separate:
            raise NotImplementedError("source separation runs the external `demucs` CLI (style_transfer.py:77-90), which is "
                                      "outside this engine: separate the stems first and pass --do_not_separate True")

    # reload model weights from the target checkpoint path
    def reload_weights(self, ckpt_paths, ddp=True):
        for cur_model_name in self.models.keys():
            checkpoint = torch.load(ckpt_paths[cur_model_name], map_location=self.device)
            from collections import OrderedDict
            new_state_dict = OrderedDict()
            for k, v in checkpoint["model"].items():
                # remove `module.` if the model was trained with DDP
                name = k[7:] if ddp else k
                new_state_dict[name] = v
            # load params
            self.models[cur_model_name].load_state_dict(new_state_dict)
            print(f"---reloaded checkpoint weights : {cur_model_name} ---")

    # ---- device-side pieces ----
    def encode_reference(self, ref_batches):
        """Mean embedding over every reference segment (style_transfer.py:144-153)."""
        feats = []
        with torch.no_grad():
            for cur_ref_data in ref_batches:
                cur_ref_data = cur_ref_data.to(self.device, non_blocking=True)
                feats.append(self.models["effects_encoder"].eval()(cur_ref_data))
        return torch.cat(feats, dim=0).mean(dim=0)

    def convert(self, in_batches, cond_of_batch):
        """TCN over the input batches; under torchrun each rank takes a contiguous slice of every batch and the
        results are all-gathered (shard.py).  cond_of_batch(idx) -> [2048] embedding for batch idx."""
        outs = []
        with torch.no_grad():
            for idx, cur_data in enumerate(in_batches):
                cond = cond_of_batch(idx).unsqueeze(0)
                n = cur_data.shape[0]
                lo, hi = shard.shard_bounds(n, self.world_size, self.rank)
                local = cur_data[lo:hi].to(self.device, non_blocking=True)
                if hi > lo:
                    y = self.models["mixing_converter"].eval()(local, cond)
                else:
                    y = torch.empty(0, 2, cur_data.shape[-1], device=self.device)
                if self.world_size > 1:
                    y = shard.allgather_segments(y, shard.shard_counts(n, self.world_size))
                outs.append(y.detach() if self.device_io else y.cpu().detach())
        return outs

    def combine(self, infered_data_list, length):
        # combine back to whole song (:165-169); a device tensor under --device_io, a numpy array otherwise
        seq = [torch.cat(torch.unbind(b, dim=0), dim=-1) for b in infered_data_list]
        whole = torch.cat(seq, dim=-1)[:, :length]
        return whole if self.device_io else whole.numpy()

    def write_outputs(self, cur_out_dir, inst_outputs, output_name_tag):
        """Per-instrument files (--save_each_inst) and the remix `sum(inst_outputs)` as PCM_16 (:170-177)."""
        if self.rank != 0:
            return
        if self.device_io:
            if self.args.save_each_inst:
                for name, y in zip(self.args.instruments, inst_outputs):
                    wav_io.write_wav_pcm16_from_device(os.path.join(cur_out_dir, f"{name}_{output_name_tag}.wav"), y,
                                                       self.args.sample_rate)
            wav_io.write_wav_pcm16_from_device(os.path.join(cur_out_dir, f"mixture_{output_name_tag}.wav"),
                                               torch.stack(inst_outputs, dim=0), self.args.sample_rate)
        else:
            if self.args.save_each_inst:
                for name, y in zip(self.args.instruments, inst_outputs):
                    write_wav_pcm16(os.path.join(cur_out_dir, f"{name}_{output_name_tag}.wav"), y.transpose(-1, -2),
                                    self.args.sample_rate)
            write_wav_pcm16(os.path.join(cur_out_dir, f"mixture_{output_name_tag}.wav"),
                            sum(inst_outputs).transpose(-1, -2), self.args.sample_rate)

    # Inference whole song
    def inference(self, ):
        print("\n======= Start to inference music mixing style transfer =======")
        # normalized input
        output_name_tag = 'output' if self.args.normalize_input else 'output_notnormed'

        for step, (input_stems, reference_stems, dir_name) in enumerate(self.data_loader):
            print(f"---inference file name : {dir_name[0]}---")
            cur_out_dir = dir_name[0].replace(self.target_dir, self.output_dir)
            os.makedirs(cur_out_dir, exist_ok=True)
            ''' stem-level inference '''
            inst_outputs = []
            for cur_inst_idx, cur_inst_name in enumerate(self.args.instruments):
                print(f'\t{cur_inst_name}...')
                ''' segmentize whole songs into batch '''
                if len(input_stems[0][cur_inst_idx][0]) > self.args.segment_length:
                    cur_inst_input_stem = self.batchwise_segmentization(input_stems[0][cur_inst_idx],
                                                                        dir_name[0],
                                                                        segment_length=self.args.segment_length,
                                                                        discard_last=False)
                else:
                    cur_inst_input_stem = [input_stems[:, cur_inst_idx]]
                if len(reference_stems[0][cur_inst_idx][0]) > self.args.segment_length * 2:
                    cur_inst_reference_stem = self.batchwise_segmentization(reference_stems[0][cur_inst_idx],
                                                                            dir_name[0],
                                                                            segment_length=self.args.segment_length_ref,
                                                                            discard_last=False)
                else:
                    cur_inst_reference_stem = [reference_stems[:, cur_inst_idx]]

                ''' inference '''
                # first extract reference style embedding (every rank computes it: 1 % of the work, no broadcast needed
                # for file-level inference; the benchmark path uses shard.sharded_style_transfer with the broadcast)
                infered_ref_data_avg = self.encode_reference(cur_inst_reference_stem)
                # mixing style converter
                infered_data_list = self.convert(cur_inst_input_stem, lambda idx: infered_ref_data_avg)
                # final output of current instrument
                fin_data_out_inst = self.combine(infered_data_list, input_stems[0][cur_inst_idx].shape[-1])

                inst_outputs.append(fin_data_out_inst)
            # per-instrument outputs (--save_each_inst) and the remix
            self.write_outputs(cur_out_dir, inst_outputs, output_name_tag)

    # Inference whole song
    def inference_interpolation(self, ):
        print("\n======= Start to inference interpolation examples =======")
        # normalized input
        output_name_tag = 'output_interpolation' if self.args.normalize_input else 'output_notnormed_interpolation'

        for step, (input_stems, reference_stems_A, reference_stems_B, dir_name) in enumerate(self.data_loader):
            print(f"---inference file name : {dir_name[0]}---")
            cur_out_dir = dir_name[0].replace(self.target_dir, self.output_dir)
            os.makedirs(cur_out_dir, exist_ok=True)
            ''' stem-level inference '''
            inst_outputs = []
            for cur_inst_idx, cur_inst_name in enumerate(self.args.instruments):
                print(f'\t{cur_inst_name}...')
                ''' segmentize whole song '''
                # segmentize input according to number of interpolating segments
                interpolate_segment_length = input_stems[0][cur_inst_idx].shape[1] // self.args.interpolate_segments + 1
                cur_inst_input_stem = self.batchwise_segmentization(input_stems[0][cur_inst_idx],
                                                                    dir_name[0],
                                                                    segment_length=interpolate_segment_length,
                                                                    discard_last=False)
                # batchwise segmentize 2 reference tracks
                if len(reference_stems_A[0][cur_inst_idx][0]) > self.args.segment_length_ref:
                    cur_inst_reference_stem_A = self.batchwise_segmentization(reference_stems_A[0][cur_inst_idx],
                                                                              dir_name[0],
                                                                              segment_length=self.args.segment_length_ref,
                                                                              discard_last=False)
                else:
                    cur_inst_reference_stem_A = [reference_stems_A[:, cur_inst_idx]]
                if len(reference_stems_B[0][cur_inst_idx][0]) > self.args.segment_length_ref:
                    # the reference cuts B with `segment_length`, A with `segment_length_ref` (:205 vs :212)
                    cur_inst_reference_stem_B = self.batchwise_segmentization(reference_stems_B[0][cur_inst_idx],
                                                                              dir_name[0],
                                                                              segment_length=self.args.segment_length,
                                                                              discard_last=False)
                else:
                    cur_inst_reference_stem_B = [reference_stems_B[:, cur_inst_idx]]

                ''' inference '''
                infered_ref_data_avg_A = self.encode_reference(cur_inst_reference_stem_A)
                infered_ref_data_avg_B = self.encode_reference(cur_inst_reference_stem_B)

                # perform linear interpolation on embedding space; the weight is indexed by BATCH (:247-251)
                def cond_of_batch(cur_idx):
                    cur_weight = (self.args.interpolate_segments - 1 - cur_idx) / (self.args.interpolate_segments - 1)
                    return cur_weight * infered_ref_data_avg_A + (1 - cur_weight) * infered_ref_data_avg_B

                infered_data_list = self.convert(cur_inst_input_stem, cond_of_batch)
                fin_data_out_inst = self.combine(infered_data_list, input_stems[0][cur_inst_idx].shape[-1])
                inst_outputs.append(fin_data_out_inst)
            # per-instrument outputs (--save_each_inst) and the remix
            self.write_outputs(cur_out_dir, inst_outputs, output_name_tag)

    # function that segmentize an entire song into batch
    def batchwise_segmentization(self, target_song, song_name, segment_length, discard_last=False):
        assert target_song.shape[-1] >= self.args.segment_length, \
            f"Error : Insufficient duration!\n\t \
                Target song's length is shorter than segment length.\n\t \
                Song name : {song_name}\n\t \
                Consider changing the 'segment_length' or song with sufficient duration"

        # discard restovers (last segment)
        if discard_last:
            target_length = target_song.shape[-1] - target_song.shape[-1] % segment_length
            target_song = target_song[:, :target_length]
        # pad last segment
        else:
            pad_length = segment_length - target_song.shape[-1] % segment_length
            target_song = torch.cat((target_song, torch.zeros(2, pad_length, device=target_song.device)), axis=-1)

        # segmentize according to the given segment_length
        whole_batch_data = []
        batch_wise_data = []
        for cur_segment_idx in range(target_song.shape[-1] // segment_length):
            batch_wise_data.append(target_song[..., cur_segment_idx * segment_length:(cur_segment_idx + 1) * segment_length])
            if len(batch_wise_data) == self.args.batch_size:
                whole_batch_data.append(torch.stack(batch_wise_data, dim=0))
                batch_wise_data = []
        if batch_wise_data:
            whole_batch_data.append(torch.stack(batch_wise_data, dim=0))

        return whole_batch_data

    # save current inference arguments
    def save_args(self, params):
        info = '\n[args]\n'
        parser = getattr(params, "_parser", None)
        groups = parser._action_groups if parser is not None else []
        for sub_args in groups:
            if sub_args.title in ['positional arguments', 'optional arguments', 'options']:
                continue
            size_sub = len(sub_args._group_actions)
            info += f'  {sub_args.title} ({size_sub})\n'
            for i, arg in enumerate(sub_args._group_actions):
                prefix = '-'
                info += f'      {prefix} {arg.dest:20s}: {getattr(params, arg.dest)}\n'
        info += '\n'

        os.makedirs(self.output_dir, exist_ok=True)
        record_path = f"{self.output_dir}style_transfer_inference_configurations.txt"
        with open(record_path, 'w') as f:
            np.savetxt(f, [info], delimiter=" ", fmt="%s")


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
    inference_args.add_argument('--instruments', type=str2bool, default=["drums", "bass", "other", "vocals"], help='instrumental tracks to perform style transfer')
    inference_args.add_argument('--stem_level_directory_name', type=str, default='separated')
    inference_args.add_argument('--save_each_inst', type=str2bool, default=False)
    inference_args.add_argument('--do_not_separate', type=str2bool, default=False)
    inference_args.add_argument('--separation_model', type=str, default='mdx_extra')
    # FX normalization
    inference_args.add_argument('--normalize_input', type=str2bool, default=True)
    inference_args.add_argument('--normalization_order', type=str2bool, default=['loudness', 'eq', 'compression', 'imager', 'loudness'])  # Effects to be normalized, order matters
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
        dist.init_process_group("nccl")

    # load network configurations
    with open(os.path.join(currentdir, 'configs.yaml'), 'r') as f:
        configs = yaml.full_load(f)
    args.cfg_encoder = configs['Effects_Encoder']['default']
    args.cfg_converter = configs['TCN']['default']

    # Perform music mixing style transfer
    inference_style_transfer = Mixing_Style_Transfer_Inference(args)
    if args.interpolation:
        inference_style_transfer.inference_interpolation()
    else:
        inference_style_transfer.inference()


if __name__ == '__main__':
    main()
