class ParameterList:
    def __init__(self):
        self._names = []

    def add(self, parameter):
        setattr(self, parameter.name, parameter)
        self._names.append(parameter.name)

    def __iter__(self):
        return iter(getattr(self, n) for n in self._names)
