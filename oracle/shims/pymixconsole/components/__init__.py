from . import iirfilter, allpass, comb  # noqa: F401
