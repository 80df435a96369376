"""Import-time stand-in for `lpips` (absent from this image; never called on the HairFast.swap(align=False) inference
path -- SURVEY Appendix D).  Any use raises."""


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)

    class _Missing:
        def __init__(self, *a, **k):
            raise RuntimeError("lpips stub: '%s' is not available in this environment" % name)

    _Missing.__name__ = name
    return _Missing
