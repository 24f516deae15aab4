"""`mixing_manipulator` surface (EQ / compressor / imager / gain + chain) backed by csrc/fx.cu."""
from .audio_effects_chain import create_effects_augmentation_chain  # noqa: F401
from .common_audioeffects import (FX_ALL, FX_COMP, FX_EQ, FX_GAIN, FX_IMAGER, FX_RMSNORM, AugmentationChain,  # noqa: F401
                                  Compressor, Equaliser, Gain, MidSideImager, Parameter, ParameterList, Processor,
                                  fx_chain_forward)
