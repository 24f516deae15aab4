"""Effects-chain factory -- same surface as the reference's `mixing_manipulator/audio_effects_chain.py:17-95`: the four
effects of BASELINE's FX chain (eq, comp, imager, gain) plus the panner.  The expander and the reverbs
(algorithmic: pymixconsole comb / all-pass components; convolutional: impulse-response data sets) raise NotImplementedError."""
from .common_audioeffects import (AugmentationChain, Compressor, Equaliser, Gain, MidSideImager, Panner, Processor)


# create augmentation effects chain according to targeted effects with their applying probability
def create_effects_augmentation_chain(effects,
                                      ir_dir_path=None,
                                      sample_rate=44100,
                                      shuffle=False,
                                      parallel=False,
                                      parallel_weight_factor=None):
    '''
        Args:
            effects (list of tuples or string) : First tuple element is string denoting the target effects.
                                                    Second tuple element is probability of applying current effects.
            ir_dir_path (string) : unused here (convolutional reverb is outside the B200 FX chain)
            sample_rate (int) : using sampling rate
            shuffle (boolean) : shuffle FXs inside current FX chain
            parallel (boolean) : compute parallel FX computation (alpha * input + (1-alpha) * manipulated output)
            parallel_weight_factor : the value of alpha for parallel FX computation. default=None : random value in between (0.0, 0.5)
    '''
    fx_list = []
    apply_prob = []
    for cur_fx in effects:
        # store probability to apply current effects. default is to set as 100%
        if isinstance(cur_fx, tuple):
            apply_prob.append(cur_fx[1])
            cur_fx = cur_fx[0]
        else:
            apply_prob.append(1)

        # processors of each audio effects
        if isinstance(cur_fx, AugmentationChain) or isinstance(cur_fx, Processor):
            fx_list.append(cur_fx)
        elif cur_fx.lower() == 'gain':
            fx_list.append(Gain())
        elif 'eq' in cur_fx.lower():
            fx_list.append(Equaliser(n_channels=2, sample_rate=sample_rate))
        elif 'comp' in cur_fx.lower():
            fx_list.append(Compressor(sample_rate=sample_rate))
        elif 'expand' in cur_fx.lower():
            raise NotImplementedError(f"effect {cur_fx!r}: the expander has no B200 kernel yet (SURVEY.md 8f-4)")
        elif 'pan' in cur_fx.lower():
            fx_list.append(Panner())
        elif 'image' in cur_fx.lower():
            fx_list.append(MidSideImager())
        elif any(k in cur_fx.lower() for k in ('algorithmic', 'reverb')):
            raise NotImplementedError(f"effect {cur_fx!r}: the reverbs have no B200 kernel yet (SURVEY.md 8f-4)")
        else:
            raise ValueError(f"make sure the target effects are in the Augment FX chain : received fx called {cur_fx}")

    aug_chain_in = []
    for cur_i, cur_fx in enumerate(fx_list):
        normalize = False if isinstance(cur_fx, AugmentationChain) or cur_fx.name == 'Gain' else True
        aug_chain_in.append((cur_fx, apply_prob[cur_i], normalize))

    return AugmentationChain(fxs=aug_chain_in, shuffle=shuffle, parallel=parallel,
                             parallel_weight_factor=parallel_weight_factor)
