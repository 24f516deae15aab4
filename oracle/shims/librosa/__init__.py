"""TEST INFRASTRUCTURE ONLY -- empty stand-in so that the reference's mixing_manipulator modules import in this image (librosa
is not installed).  Nothing the pinned code paths call lives here."""
from . import display  # noqa: F401
