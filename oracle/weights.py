"""TEST INFRASTRUCTURE -- the seeded weight / audio generators live in the package (bench.py feeds the CUDA path with
them too); re-exported here so the oracle-side code reads `oracle.weights`."""
from music_mixing_style_transfer_b200.synthetic import *  # noqa: F401,F403
from music_mixing_style_transfer_b200.synthetic import (ENC_CHANNELS, ENC_KERNELS, ENC_STRIDES, TCN_OUTPUT_GAIN,  # noqa: F401
                                                         make_encoder_state_dict, make_tcn_state_dict, synthetic_audio)
