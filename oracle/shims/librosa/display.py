"""TEST INFRASTRUCTURE ONLY -- empty stand-in (see librosa/__init__.py)."""
