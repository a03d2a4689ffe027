"""Launch accounting for bench.py: counts every C-ABI kernel launch and, when enabled, brackets
each launch with CUDA events on the current stream together with its algorithmic FLOPs/bytes."""
import torch

_count = 0
_events_on = False
_records = []


def reset():
    global _count, _records
    _count = 0
    _records = []


def uncount():
    """Take back the last launch() whose C-ABI call reported that it launched nothing."""
    global _count
    _count -= 1
    if _events_on and _records:
        _records.pop()


def launch_count():
    return _count


def enable_events(on):
    global _events_on, _records
    _events_on = bool(on)
    if on:
        _records = []


def io_bytes(io):
    """Algorithmic bytes of a memory-bound launch: every listed tensor/view read or written once at its storage dtype."""
    return float(sum(t.numel() * t.element_size() for t in io if t is not None))


def launch(kind, fn, flops=0.0, nbytes=0.0, tag="", io=None):
    global _count
    _count += 1
    if not _events_on:
        return fn()
    if io is not None:
        nbytes = io_bytes(io)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = fn()
    e1.record()
    _records.append((kind, e0, e1, float(flops), float(nbytes), tag))
    return rc


def records():
    torch.cuda.synchronize()
    return [(k, e0.elapsed_time(e1), f, b, t) for (k, e0, e1, f, b, t) in _records]


def summary():
    out = {}
    for k, ms, f, b, _ in records():
        d = out.setdefault(k, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
        d["ms"] += ms
        d["flops"] += f
        d["bytes"] += b
        d["launches"] += 1
    return out
