"""`mixing_manipulator` surface (EQ / compressor / imager / gain / panner / Haas + chain, input FX normaliser) backed by
csrc/fx2.cu and csrc/fxnorm.cu."""
from .audio_effects_chain import create_effects_augmentation_chain  # noqa: F401
from .common_audioeffects import (FX_ALL, FX_COMP, FX_EQ, FX_GAIN, FX_IMAGER, FX_RMSNORM, AugmentationChain,  # noqa: F401
                                  Compressor, Equaliser, Gain, Haas, MidSideImager, Panner, Parameter, ParameterList,
                                  Processor,
                                  fx_chain_forward)
from .data_normalization import Audio_Effects_Normalizer  # noqa: F401,E402
