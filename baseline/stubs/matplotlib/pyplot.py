def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)

    def _unavailable(*a, **k):
        raise RuntimeError("matplotlib stub: pyplot.%s is not available in this environment" % name)
    return _unavailable
