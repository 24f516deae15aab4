"""
    Inference code of extracting embeddings from music recordings using FXencoder on the B200 engine -- the entry that
    mirrors the reference's `inference/feature_extraction.py` (class FXencoder_Inference :20, save_averaged_embeddings :69,
    batchwise_segmentization :114, flags :168-182; SURVEY.md 8f-3, BASELINE config 1).

    Process : extracts FX embeddings of each song inside the target directory.

    What is the same: flags, recursive `**/*.wav` walk, mono -> stereo duplication (:87-89), segmentation with zero padding of
    the last segment (a full extra zero segment when the length is an exact multiple, :127-128), batches of `batch_size`
    segments through the encoder, mean over all segments, `<name>_fx_embedding.npy` next to the file (or under output_dir)
    and the `...configurations.txt` dump.
    What differs, on purpose: the encoder runs on libmst_b200.so (CUDA only; `--inference_device cpu` raises instead of
    silently running a different code path); WAVs are read with the stdlib `wave` module (the reference goes through its
    `data_loader` package, which imports librosa / soundfile) and converted on the GPU (wav_io.py).
"""
import os
import sys
from glob import glob

import numpy as np
import torch

currentdir = os.path.dirname(os.path.realpath(__file__))
sys.path.append(os.path.dirname(os.path.dirname(currentdir)))
from music_mixing_style_transfer_b200.networks import FXencoder  # noqa: E402
from music_mixing_style_transfer_b200 import wav_io  # noqa: E402


class FXencoder_Inference:
    def __init__(self, args, trained_w_ddp=True):
        if args.inference_device == 'cpu' or not torch.cuda.is_available():
            raise RuntimeError("FXencoder_Inference (B200 engine) needs a CUDA device: there is no CPU path "
                               "(pass --inference_device gpu; the reference's CPU forward is only the parity oracle)")
        self.device = torch.device("cuda", torch.cuda.current_device())

        # inference computational hyperparameters
        self.args = args
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
        ckpt_paths = {'effects_encoder': args.ckpt_path_enc}
        # reload saved model weights
        self.reload_weights(ckpt_paths, ddp=trained_w_ddp)

        # save current arguments
        self.save_args(args)

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

    def embed_song(self, target_song_whole, target_file_path="<tensor>"):
        """[2, T] float32 (device) -> averaged FX embedding [2048] (device): segment, encode every batch, mean (:92-107)."""
        whole_batch_data = self.batchwise_segmentization(target_song_whole, target_file_path)
        infered_c_list = []
        with torch.no_grad():
            self.models["effects_encoder"].eval()
            for cur_data in whole_batch_data:
                infered_c_list.append(self.models["effects_encoder"](cur_data.to(self.device)))
        return torch.mean(torch.cat(infered_c_list, dim=0), dim=0).squeeze()

    # save averaged embedding from whole songs
    def save_averaged_embeddings(self, ):
        print(f'\n\n=====Inference seconds : {self.time_in_seconds}=====')
        target_file_paths = sorted(glob(os.path.join(self.target_dir, '**', '*.wav'), recursive=True))
        for step, target_file_path in enumerate(target_file_paths):
            print(f"\nInference step : {step+1}/{len(target_file_paths)}")
            print(f"---current file path : {target_file_path}---")
            ''' load waveform signal: raw PCM -> GPU -> float32 [2, T] (mono duplicated), csrc/pcm.cu '''
            target_song_whole = wav_io.load_wav_to_device(target_file_path, self.device, sample_rate=self.sample_rate)
            avg_c_feat = self.embed_song(target_song_whole, target_file_path).cpu().detach().numpy()
            # save outputs
            cur_output_path = target_file_path.replace(self.target_dir, self.output_dir).replace('.wav', '_fx_embedding.npy')
            os.makedirs(os.path.dirname(cur_output_path), exist_ok=True)
            np.save(cur_output_path, avg_c_feat)

    # function that segmentize an entire song into batch
    def batchwise_segmentization(self, target_song, target_file_path, discard_last=False):
        assert target_song.shape[-1] >= self.segment_length, \
            f"Error : Insufficient duration!\n\t \
                Target song's length is shorter than segment length.\n\t \
                Song name : {target_file_path}\n\t \
                Consider changing the 'segment_length' or song with sufficient duration"

        # discard restovers (last segment)
        if discard_last:
            target_length = target_song.shape[-1] - target_song.shape[-1] % self.segment_length
            target_song = target_song[:, :target_length]
        # pad last segment
        else:
            pad_length = self.segment_length - target_song.shape[-1] % self.segment_length
            target_song = torch.cat((target_song, torch.zeros(2, pad_length, device=target_song.device)), axis=-1)

        whole_batch_data = []
        batch_wise_data = []
        for cur_segment_idx in range(target_song.shape[-1] // self.segment_length):
            batch_wise_data.append(target_song[..., cur_segment_idx * self.segment_length:(cur_segment_idx + 1) * self.segment_length])
            if len(batch_wise_data) == self.batch_size:
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
        record_path = f"{self.output_dir}feature_extraction_inference_configurations.txt"
        with open(record_path, 'w') as f:
            np.savetxt(f, [info], delimiter=" ", fmt="%s")


def build_parser():
    ''' Configurations for inferencing music effects encoder '''
    import argparse
    repo_root = os.path.dirname(os.path.dirname(currentdir))
    default_ckpt_path = os.path.join(repo_root, 'weights', 'FXencoder_ps.pt')

    parser = argparse.ArgumentParser()

    directory_args = parser.add_argument_group('Directory args')
    directory_args.add_argument('--target_dir', type=str, default='./samples/')
    directory_args.add_argument('--output_dir', type=str, default=None, help='if no output_dir is specified (None), the results will be saved inside the target_dir')
    directory_args.add_argument('--ckpt_path_enc', type=str, default=default_ckpt_path)

    inference_args = parser.add_argument_group('Inference args')
    inference_args.add_argument('--segment_length', type=int, default=44100 * 10)  # segmentize input according to this duration
    inference_args.add_argument('--batch_size', type=int, default=1)              # for processing long audio
    inference_args.add_argument('--inference_device', type=str, default='gpu', help="the B200 engine only runs on CUDA devices")
    return parser


def main(argv=None):
    import yaml
    parser = build_parser()
    args = parser.parse_args(argv)
    args._parser = parser

    # load network configurations
    with open(os.path.join(currentdir, 'configs.yaml'), 'r') as f:
        configs = yaml.full_load(f)
    args.cfg_encoder = configs['Effects_Encoder']['default']

    # Extract features using pre-trained FXencoder
    inference_encoder = FXencoder_Inference(args)
    inference_encoder.save_averaged_embeddings()


if __name__ == '__main__':
    main()
