"""TEST INFRASTRUCTURE ONLY -- see pyloudnorm/__init__.py."""
from oracle.norm_oracle import loudness_gain_apply as loudness  # noqa: F401


def peak(data, target):
    """pyloudnorm.normalize.peak: gain = 10^(target / 20) / max|data| (restated; used by get_comp_matching,
    utils_data_normalization.py:374)."""
    import numpy as np
    return (np.power(10.0, target / 20.0) / np.max(np.abs(data))) * data
