"""TEST INFRASTRUCTURE ONLY -- see pyloudnorm/__init__.py."""
from oracle.norm_oracle import loudness_gain_apply as loudness  # noqa: F401
