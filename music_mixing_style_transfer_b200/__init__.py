"""B200-native (sm_100a) engine for the segment-batched style-transfer forward path of
jhtonyKoo/music_mixing_style_transfer: FXencoder, MixFXcloner TCN and the EQ/compressor/imager/gain FX chain,
behind the reference's own module surfaces.  Compute lives in libmst_b200.so (include/mst_b200.h); there is no CPU
or PyTorch fallback."""
__all__ = ["networks", "mixing_manipulator", "inference", "shard"]
