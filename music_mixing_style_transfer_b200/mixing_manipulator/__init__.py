"""`mixing_manipulator` surface (EQ / compressor / imager / gain / panner / Haas / reverbs + chain factories, input FX
normaliser) backed by csrc/fx2.cu, csrc/fxnorm.cu, csrc/spectral.cu and csrc/reverb.cu."""
from .audio_effects_chain import (create_effects_augmentation_chain, create_inst_effects_augmentation_chain,  # noqa: F401
                                  load_impulse_responses)
from .common_audioeffects import (FX_ALL, FX_COMP, FX_EQ, FX_GAIN, FX_IMAGER, FX_RMSNORM, AlgorithmicReverb,  # noqa: F401
                                  AugmentationChain, Compressor, ConvolutionalReverb, Equaliser, Gain, Haas, MidSideImager,
                                  Panner, Parameter, ParameterList, Processor, fx_chain_forward)
from .data_normalization import Audio_Effects_Normalizer  # noqa: F401,E402
