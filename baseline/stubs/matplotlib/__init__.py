"""Import-time stand-in for matplotlib (models/encoder4editing/models/psp.py:1-3 calls matplotlib.use('Agg'); plotting
is never reached during inference)."""


def use(*args, **kwargs):
    return None
