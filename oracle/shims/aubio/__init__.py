"""TEST INFRASTRUCTURE ONLY -- empty stand-in: aubio (onset detection of the compressor matching) is not installed."""
