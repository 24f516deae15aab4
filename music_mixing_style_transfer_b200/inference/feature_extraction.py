"""FX-embedding extraction on the B200 engine -- the entry that mirrors the reference's `inference/feature_extraction.py`
(class FXencoder_Inference :20, save_averaged_embeddings :69, batchwise_segmentization :114, flags :168-182; SURVEY.md
8f-3, BASELINE config 1): one averaged FXencoder embedding per WAV file under a directory.

Same as the reference: the flags, the recursive `**/*.wav` walk, mono -> stereo duplication (:87-89), segments of
`segment_length` with the zero-padded tail (a full extra zero segment when the length is an exact multiple, :127-128),
batches of `batch_size` segments through the encoder, the mean over all segments saved as `<name>_fx_embedding.npy` next to
the file (or under `--output_dir`), and the `...configurations.txt` dump.
Different on purpose: the encoder runs on libmst_b200.so (CUDA only; `--inference_device cpu` raises instead of silently
running another code path); WAVs are read with the stdlib `wave` module and converted on the GPU (wav_io.py / csrc/pcm.cu)
-- the reference goes through its `data_loader` package, which needs librosa / soundfile; segments are views cut on the
device.
"""
import os
import sys
from glob import glob

import numpy as np
import torch

currentdir = os.path.dirname(os.path.realpath(__file__))
sys.path.append(os.path.dirname(os.path.dirname(currentdir)))
from music_mixing_style_transfer_b200 import wav_io  # noqa: E402
from music_mixing_style_transfer_b200.inference._common import (checkpoint_state_dict, dump_arguments,  # noqa: E402
                                                                 segment_into_batches)
from music_mixing_style_transfer_b200.networks import FXencoder  # noqa: E402

SAMPLE_RATE = 44100     # the checkpoints were trained at 44.1 kHz; other rates are rejected by the WAV reader


class FXencoder_Inference:
    def __init__(self, args, trained_w_ddp=True):
        if args.inference_device == 'cpu' or not torch.cuda.is_available():
            raise RuntimeError("FXencoder_Inference (B200 engine) needs a CUDA device: there is no CPU path "
                               "(pass --inference_device gpu; the reference's CPU forward is only the parity oracle)")
        self.args = args
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.segment_length, self.batch_size = args.segment_length, args.batch_size
        self.sample_rate = SAMPLE_RATE
        self.time_in_seconds = int(args.segment_length // SAMPLE_RATE)
        self.target_dir = args.target_dir
        self.output_dir = args.target_dir if args.output_dir is None else args.output_dir
        self.models = {'effects_encoder': FXencoder(args.cfg_encoder).to(self.device)}
        self.reload_weights({'effects_encoder': args.ckpt_path_enc}, ddp=trained_w_ddp)
        self.save_args(args)

    def reload_weights(self, ckpt_paths, ddp=True):
        for name, model in self.models.items():
            model.load_state_dict(checkpoint_state_dict(ckpt_paths[name], self.device, ddp=ddp))
            print(f"---reloaded checkpoint weights : {name} ---")

    def batchwise_segmentization(self, target_song, target_file_path, discard_last=False):
        return segment_into_batches(target_song, self.segment_length, self.batch_size, name=target_file_path,
                                    discard_last=discard_last)

    def embed_song(self, song, name="<tensor>"):
        """[2, T] float32 -> averaged FX embedding [2048] on the device: every batch of segments through the encoder, mean
        over all segments (reference :92-107)."""
        encoder = self.models["effects_encoder"].eval()
        with torch.no_grad():
            feats = [encoder(batch.to(self.device)) for batch in self.batchwise_segmentization(song, name)]
        return torch.cat(feats, dim=0).mean(dim=0).squeeze()

    def save_averaged_embeddings(self):
        print(f'\n\n=====Inference seconds : {self.time_in_seconds}=====')
        paths = sorted(glob(os.path.join(self.target_dir, '**', '*.wav'), recursive=True))
        for step, path in enumerate(paths, start=1):
            print(f"\nInference step : {step}/{len(paths)}\n---current file path : {path}---")
            # raw PCM -> GPU -> float32 [2, T], mono duplicated (csrc/pcm.cu)
            song = wav_io.load_wav_to_device(path, self.device, sample_rate=self.sample_rate)
            out_path = path.replace(self.target_dir, self.output_dir).replace('.wav', '_fx_embedding.npy')
            os.makedirs(os.path.dirname(out_path), exist_ok=True)
            np.save(out_path, self.embed_song(song, path).cpu().numpy())

    def save_args(self, params):
        dump_arguments(params, f"{self.output_dir}feature_extraction_inference_configurations.txt")


def build_parser():
    ''' Configurations for inferencing music effects encoder '''
    import argparse
    weights_dir = os.path.join(os.path.dirname(os.path.dirname(currentdir)), 'weights')
    parser = argparse.ArgumentParser()
    grp = parser.add_argument_group('Directory args')
    grp.add_argument('--target_dir', type=str, default='./samples/')
    grp.add_argument('--output_dir', type=str, default=None,
                     help='if no output_dir is specified (None), the results will be saved inside the target_dir')
    grp.add_argument('--ckpt_path_enc', type=str, default=os.path.join(weights_dir, 'FXencoder_ps.pt'))
    grp = parser.add_argument_group('Inference args')
    grp.add_argument('--segment_length', type=int, default=44100 * 10, help='segmentize input according to this duration')
    grp.add_argument('--batch_size', type=int, default=1, help='for processing long audio')
    grp.add_argument('--inference_device', type=str, default='gpu', help="the B200 engine only runs on CUDA devices")
    return parser


def main(argv=None):
    import yaml
    parser = build_parser()
    args = parser.parse_args(argv)
    args._parser = parser
    with open(os.path.join(currentdir, 'configs.yaml'), 'r') as f:
        args.cfg_encoder = yaml.full_load(f)['Effects_Encoder']['default']
    FXencoder_Inference(args).save_averaged_embeddings()


if __name__ == '__main__':
    main()
