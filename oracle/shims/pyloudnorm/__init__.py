"""TEST INFRASTRUCTURE ONLY -- restatement of the parts of pyloudnorm==0.1.0 (requirements.txt:9; PyPI, NOT vendored in the
reference and not installed here) that the reference calls: `Meter(rate).integrated_loudness(data)` and
`normalize.loudness(data, input_loudness, target_loudness)` (fx_utils.py:223-229).  PARITY UNPINNED: this follows the published
ITU-R BS.1770-4 algorithm as pyloudnorm implements it (K-weighting = RBJ high shelf +4 dB / 1500 Hz / Q 1/sqrt(2) followed by
an RBJ high pass 38 Hz / Q 0.5, each applied with scipy.signal.lfilter and stored back in the data's dtype; 400 ms blocks with
75 % overlap; absolute gate -70 LUFS, relative gate -10 LU), not the package's source.  The arithmetic itself lives in
oracle/norm_oracle.py so that the oracle does not depend on import order."""
from oracle.norm_oracle import Meter  # noqa: F401
from . import normalize  # noqa: F401
