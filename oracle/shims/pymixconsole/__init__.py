"""TEST INFRASTRUCTURE ONLY -- stub of the un-vendored PyPI package `pymixconsole==0.0.1`
(reference requirements.txt:12) so that mixing_manipulator/common_audioeffects.py imports.
Only the class shells the reference touches are provided; the IIR arithmetic is restated in
oracle/fx_oracle.py (RBJ cookbook + scipy.signal.lfilter float64) -- PARITY UNPINNED for the EQ."""
from . import components, parameter, parameter_list, processor  # noqa: F401
