class Processor:
    """The reference monkey-patches __init__/__repr__/update (common_audioeffects.py:86-88)."""

    def __init__(self, *a, **k):
        pass
