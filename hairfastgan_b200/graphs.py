"""CUDA-graph replay of the encoder-family forwards at swap()'s batch sizes.

One `swap()` calls e4e at B = 3 and 2, the FS encoder at B = 3, FeatureEncoderMult at B = 1 twice, FeatureiResnet once
and BiSeNet five times at B = 1: 100-350 kernel launches each, every one a ctypes call from Python.
At these sizes the kernels finish faster than the host can enqueue them, so the hot path of a single triple was
launch-bound (round-2 measurement: 26 ms of hot path in a swap whose kernels need well under half of that).  Each
network forward is a fixed launch sequence for a given input shape, so after the second call with the same signature it
is captured once (`torch.cuda.CUDAGraph`, the library only enqueues on the current stream and allocates nothing) and
replayed: input copied into the graph's static buffer, one `cudaGraphLaunch`, outputs cloned out.

* keyed on (network, input shape / dtype / device, parameter version key): a weight change recaptures
* batch <= HAIRFAST_GRAPH_MAX_BATCH (default 8): the graph's private pool holds one set of activations per signature,
  which is only wanted at latency-bound sizes (the throughput path at B = 48 stays eager)
* the generator forward (`randomize_noise=True`, no explicit noise list) is captured too, its 17 `normal_()` draws
  included: torch's graph-safe Philox state makes a replay consume exactly the (seed, offset) range the eager call
  would, so `seed_setter` reproducibility and the reference's draw order hold (tested bit for bit)
* HAIRFAST_CUDA_GRAPHS=0 disables it; a failed capture disables it for that signature (eager CUDA path, never a CPU one)
"""
from __future__ import annotations

import os

import torch

__all__ = ["run", "enabled", "stats"]

_STATS = {"captures": 0, "replays": 0, "eager": 0, "failed": 0, "replayed_kernels": 0}


def enabled() -> bool:
    return os.environ.get("HAIRFAST_CUDA_GRAPHS", "1") not in ("0", "false", "off")


def _max_batch() -> int:
    return int(os.environ.get("HAIRFAST_GRAPH_MAX_BATCH", "8"))


def stats() -> dict:
    return dict(_STATS)


def _flatten(out, acc):
    if torch.is_tensor(out):
        acc.append(out)
        return ("t", len(acc) - 1)
    if isinstance(out, (list, tuple)):
        return ("l" if isinstance(out, list) else "u", [_flatten(o, acc) for o in out])
    return ("c", out)


def _rebuild(spec, tensors):
    kind, val = spec
    if kind == "t":
        return tensors[val].clone()
    if kind == "c":
        return val
    seq = [_rebuild(s, tensors) for s in val]
    return seq if kind == "l" else tuple(seq)


class _Entry:
    __slots__ = ("graph", "static_in", "outs", "spec", "pack_key", "kernels")


def _capture(fn, xs, pack_key, restore_rng: bool):
    ent = _Entry()
    ent.pack_key = pack_key
    ent.static_in = [None if x is None else x.clone() for x in xs]
    dev = next(x for x in xs if x is not None).device
    cur = torch.cuda.current_stream(dev)
    side = torch.cuda.Stream(device=dev)
    rng = torch.cuda.get_rng_state(dev) if restore_rng else None
    side.wait_stream(cur)
    with torch.cuda.stream(side):                       # warm-up on a side stream (PyTorch's capture recipe)
        fn(*ent.static_in)
    cur.wait_stream(side)
    torch.cuda.synchronize(dev)
    if rng is not None:                                 # the warm-up's random draws must not count: the replay below is
        torch.cuda.set_rng_state(rng, dev)              # THE forward of this call and consumes the stream like eager
    ent.graph = torch.cuda.CUDAGraph()
    from . import _lib
    n0 = _lib.lib().hf_total_launch_count()
    with torch.cuda.graph(ent.graph):
        out = fn(*ent.static_in)
    ent.kernels = int(_lib.lib().hf_total_launch_count() - n0)    # this library's kernels inside the graph
    ent.outs = []
    ent.spec = _flatten(out, ent.outs)
    return ent


def run_multi(owner, tag, pack_key, fn, xs, batch: int, uses_rng: bool = False):
    """`fn(*xs)` (xs: CUDA tensors or None) -- eagerly the first time a signature is seen, captured on the second call,
    replayed afterwards.  With `uses_rng` the function draws from torch's CUDA generator inside the graph (graph-safe
    Philox: a replay consumes the same (seed, offset) range as the eager call would, so seeded runs are unchanged)."""
    first = next((x for x in xs if x is not None), None)
    if (not enabled() or first is None or not first.is_cuda or batch > _max_batch()
            or torch.cuda.is_current_stream_capturing()):
        _STATS["eager"] += 1
        return fn(*xs)
    cache = owner.__dict__.setdefault("_hf_graphs", {})
    sig = (tag, tuple(None if x is None else (tuple(x.shape), x.dtype) for x in xs), first.device.index)
    ent = cache.get(sig)
    if isinstance(ent, _Entry) and ent.pack_key != pack_key:
        ent = None                                      # weights changed: drop the stale graph
        cache.pop(sig, None)
    if ent is None:                                     # first sighting: eager (packs weights, sets kernel attributes)
        cache[sig] = 1
        _STATS["eager"] += 1
        return fn(*xs)
    if ent is False:
        _STATS["eager"] += 1
        return fn(*xs)
    if ent == 1:
        try:
            ent = cache[sig] = _capture(fn, [None if x is None else x.contiguous() for x in xs], pack_key, uses_rng)
            _STATS["captures"] += 1
        except Exception as exc:                        # noqa: BLE001 -- not capturable here: stay eager for this signature
            cache[sig] = False
            _STATS["failed"] += 1
            import warnings
            warnings.warn(f"hairfastgan_b200.graphs: CUDA-graph capture of {tag!r} failed ({exc!r}); this signature "
                          "stays on the eager CUDA path", RuntimeWarning, stacklevel=2)
            torch.cuda.synchronize(first.device)
            return fn(*xs)
    for st, x in zip(ent.static_in, xs):
        if st is not None:
            st.copy_(x)
    ent.graph.replay()
    _STATS["replays"] += 1
    _STATS["replayed_kernels"] += ent.kernels                    # hf_total_launch_count() cannot see a replay
    return _rebuild(ent.spec, ent.outs)


def run(owner, tag: str, pack_key, fn, x: torch.Tensor):
    """Single-input form (the encoder-family forwards)."""
    return run_multi(owner, tag, pack_key, fn, (x,), int(x.shape[0]) if x.dim() else 1)
