class Parameter:
    def __init__(self, name, value, kind, processor=None, units="", minimum=None, maximum=None, options=None, **kw):
        self.name, self.value, self.kind = name, value, kind
        self.processor, self.units = processor, units
        self.min, self.max, self.options = minimum, maximum, options

    def __repr__(self):
        return f"Parameter({self.name}={self.value})"
