"""TEST INFRASTRUCTURE ONLY -- empty stand-in for soxbindings (absent; outside the hot path)."""


class Transformer:
    pass
