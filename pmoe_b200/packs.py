"""Packed weight operands that follow the live parameters.

The conv / linear kernels read weights as packed [cout_pad][K] bf16 (fp32 in parity mode) operands; the fp32
`nn.Parameter`s in the reference's layout stay the source of truth (reference: the ATen convolution reads `conv.weight`
itself on every call, model/blocks/basics.py:51,54). Each packed operand is described ONCE by an int32 index map (packed
element i = parameter element idx[i], -1 = zero padding), obtained by running the layer's packing recipe on a stand-in
tensor that holds its own element numbers. After that the operand is refreshed by a gather kernel (`pmoe_pack_gather`)
into a STATIC buffer whenever the parameter changed — `(data_ptr, _version)`; the fused optimizers bump `_version` — and on
every pass while a CUDA graph is being captured, so graph replays re-pack from the live parameters as well. All stale
operands of a model are refreshed by one multi-tensor launch at the start of a pass (`refresh_all`).
"""
import ctypes as C
import weakref

import numpy as np
import torch

from . import _lib, profiler
from ._lib import check, lib, stream_ptr

_JOB = np.dtype([("w", "<u8"), ("idx", "<u8"), ("out", "<u8"), ("n", "<i8"), ("dtype", "<i4"), ("chunk0", "<i4")])
assert _JOB.itemsize == 40

_PASS = 0                     # bumped by begin_pass(): one model call = one pass
_REGISTRY = weakref.WeakKeyDictionary()   # top-level module -> _Registry
_current = None               # registry of the module whose pass is running


class Pack:
    """One packed operand: `out` (static buffer) = gather(owner, idx)."""
    __slots__ = ("owner", "idx", "out", "ver", "stamp", "code", "key", "_inv")

    def __init__(self, owner, idx, out, key):
        self.owner, self.idx, self.out, self.key = owner, idx, out, key
        self._inv = None
        self.ver, self.stamp = None, -1
        self.code = _lib.BF16 if out.dtype == torch.bfloat16 else _lib.F32

    def inverse(self):
        """int32 map parameter element -> packed element (every parameter element appears exactly once in a forward packing)."""
        if self._inv is None:
            flat = self.idx.reshape(-1).long()
            pos = torch.nonzero(flat >= 0).reshape(-1)
            n = self.owner.numel()
            inv = torch.full((n,), -1, dtype=torch.int64, device=flat.device)
            inv[flat[pos]] = pos
            if pos.numel() != n or bool((inv < 0).any()):
                raise RuntimeError("pmoe_b200 packs: this packing is not a bijection onto the parameter (no inverse map)")
            self._inv = inv.to(torch.int32).contiguous()
        return self._inv

    def version(self):
        return (self.owner.data_ptr(), self.owner._version)

    def stale(self, capturing):
        return self.ver != self.version() or (capturing and self.stamp != _PASS)

    def mark(self):
        self.ver, self.stamp = self.version(), _PASS

    def refresh(self):
        w = self.owner.detach()
        if w.dtype != torch.float32 or not w.is_contiguous():
            raise RuntimeError("pmoe_b200: parameters must be contiguous fp32 tensors (got %s)" % w.dtype)
        check(profiler.launch("pack_gather", lambda: lib().pmoe_pack_gather(
            w.data_ptr(), self.idx.data_ptr(), self.out.data_ptr(), self.code, self.idx.numel(), stream_ptr())), "pack_gather")
        self.mark()


class _Registry:
    def __init__(self):
        self.packs = []
        self.prev = None
        self.table = None      # (device table of PmoePackJob for ALL packs, n_jobs, total_chunks, key)

    def build_table(self):
        if not self.packs:
            self.table = None
            return
        key = tuple((p.owner.data_ptr(), p.idx.data_ptr(), p.out.data_ptr()) for p in self.packs)
        if self.table is not None and self.table[3] == key:
            return
        chunk = int(lib().pmoe_pack_chunk_elems())
        recs, c0 = [], 0
        for p in self.packs:
            n = p.idx.numel()
            recs.append((p.owner.data_ptr(), p.idx.data_ptr(), p.out.data_ptr(), n, p.code, c0))
            c0 += (n + chunk - 1) // chunk
        host = torch.from_numpy(np.array(recs, dtype=_JOB).view(np.uint8).copy())
        self.table = (host.to(self.packs[0].out.device), len(recs), c0, key)


def _capturing():
    return torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()


def begin_pass(module):
    """Start of one model call: refresh every registered operand whose parameter changed (all of them under capture)."""
    global _PASS, _current
    _PASS += 1
    reg = _REGISTRY.get(module)
    if reg is None:
        reg = _REGISTRY[module] = _Registry()
    reg.prev = _current
    _current = reg
    cap = _capturing()
    stale = [p for p in reg.packs if p.stale(cap)]
    if not stale:
        return
    if len(stale) == len(reg.packs) and len(stale) > 1:
        if not cap:
            reg.build_table()
        key = tuple((p.owner.data_ptr(), p.idx.data_ptr(), p.out.data_ptr()) for p in reg.packs)
        if reg.table is not None and reg.table[3] == key:
            tdev, n, chunks, _ = reg.table
            check(profiler.launch("pack_gather", lambda: lib().pmoe_pack_gather_mt(tdev.data_ptr(), n, chunks, stream_ptr())),
                  "pack_gather_mt")
            for p in stale:
                p.mark()
            return
    for p in stale:
        p.refresh()


def end_pass():
    global _current
    reg = _current
    if reg is None:
        return
    _current, reg.prev = reg.prev, None
    if not _capturing():
        reg.build_table()


def current():
    """The registry of the pass that is running (None outside a pass)."""
    return _current


class resumed:
    """Backward of a pass: operands first built there (data-gradient packings) join the registry of the module whose forward
    recorded the tape, so the next pass refreshes them with everything else in the one multi-tensor launch instead of one
    gather launch each (one per layer and expert, inside a captured graph on every replay)."""

    def __init__(self, reg):
        self.reg = reg

    def __enter__(self):
        global _current
        self.prev = _current
        if self.reg is not None:
            _current = self.reg
        return self

    def __exit__(self, *exc):
        global _current
        _current = self.prev
        if self.reg is not None and self.reg.table is None and not _capturing():
            self.reg.build_table()
        return False


def stand_in(shape, device):
    n = 1
    for s in shape:
        n *= int(s)
    if n >= (1 << 24):
        raise RuntimeError("pmoe_b200: weight with %d elements exceeds the index stand-in's exact range" % n)
    return (torch.arange(n, device=device, dtype=torch.float32) + 1.0).view(tuple(shape))


def packed(owner, key, shape, build, dtype):
    """The packed operand of parameter `owner` for recipe `build` (float tensor shaped `shape` -> packed float tensor),
    cached under `key` on the parameter. Returns (tensor, Pack)."""
    cache = owner.__dict__.setdefault("_pmoe_pack", {})
    ent = cache.get(key)
    if ent is None or ent.idx.device != owner.device or ent.out.dtype != dtype:
        with torch.no_grad():
            idx = (build(stand_in(shape, owner.device)).round().to(torch.int32) - 1).contiguous()
        ent = Pack(owner, idx, torch.empty(idx.shape, dtype=dtype, device=owner.device), key)
        if len(cache) > 12:
            cache.clear()
        cache[key] = ent
        if _current is not None:  # drop operands that were evicted from their parameter's cache, add the new one
            _current.packs = [p for p in _current.packs if p.owner.__dict__.get("_pmoe_pack", {}).get(p.key) is p]
            _current.packs.append(ent)
            _current.table = None
    if ent.stale(_capturing()):
        ent.refresh()
    return ent.out, ent


def packed_group(owners, key, shape, build, dtype):
    """K equally shaped parameters (the experts' Linears) packed by the same recipe into ONE stacked (K, rows, cols) operand
    (grouped GEMM: expert = image axis). The index map is shared; each slice is refreshed from its own parameter.
    Returns (stacked tensor, [Pack per expert])."""
    w0 = owners[0]
    gcache = w0.__dict__.setdefault("_pmoe_gpack", {})
    ids = tuple(id(o) for o in owners)
    hit = gcache.get(key)
    if hit is None or hit[0] != ids or hit[1].device != w0.device or hit[1].dtype != dtype:
        with torch.no_grad():
            idx = (build(stand_in(shape, w0.device)).round().to(torch.int32) - 1).contiguous()
        stack = torch.empty((len(owners),) + tuple(idx.shape), dtype=dtype, device=w0.device)
        ents = []
        for e, o in enumerate(owners):
            ent = Pack(o, idx, stack[e], ("g", key))
            o.__dict__.setdefault("_pmoe_pack", {})[("g", key)] = ent
            ents.append(ent)
        if len(gcache) > 12:
            gcache.clear()
        hit = gcache[key] = (ids, stack, ents)
        if _current is not None:
            _current.packs = [p for p in _current.packs if p.owner.__dict__.get("_pmoe_pack", {}).get(p.key) is p]
            _current.packs += ents
            _current.table = None
    cap = _capturing()
    for ent in hit[2]:
        if ent.stale(cap):
            ent.refresh()
    return hit[1], hit[2]


def scatter_grad(packed_grad, idx, dst_flat, accumulate, alpha=1.0):
    """dst_flat[idx[i]] (+)= alpha * packed_grad[i]: packed fp32 weight gradient -> the parameter's own layout."""
    assert packed_grad.dtype == torch.float32 and packed_grad.is_contiguous() and dst_flat.dtype == torch.float32
    assert packed_grad.numel() == idx.numel(), (packed_grad.shape, idx.shape)
    check(profiler.launch("unpack_scatter", lambda: lib().pmoe_unpack_scatter(
        packed_grad.data_ptr(), idx.data_ptr(), dst_flat.data_ptr(), idx.numel(), float(alpha), int(accumulate), stream_ptr())),
        "unpack_scatter")


def unpack_grads(packed_grad, pack, dst_flats, accumulates, alpha=1.0):
    """Packed fp32 weight gradient(s) -> gradient slot(s) in the parameter layout through the pack's inverse map (coalesced
    writes). packed_grad: one packed gradient (shape of pack.idx) with one slot, or (K, ...) stacked gradients of K equally shaped
    parameters with K slots (a None slot is skipped)."""
    inv = pack.inverse()
    n_packed = pack.idx.numel()
    K = len(dst_flats)
    assert packed_grad.dtype == torch.float32 and packed_grad.is_contiguous() and packed_grad.numel() == K * n_packed
    for g0 in range(0, K, 16):
        g1 = min(K, g0 + 16)
        ptrs = (C.c_void_p * (g1 - g0))(*[None if d is None else d.data_ptr() for d in dst_flats[g0:g1]])
        accs = (C.c_int32 * (g1 - g0))(*[int(bool(a)) for a in accumulates[g0:g1]])
        base = packed_grad.data_ptr() + g0 * n_packed * 4
        check(profiler.launch("unpack_scatter", lambda: lib().pmoe_unpack_gather_group(
            base, inv.data_ptr(), ptrs, accs, g1 - g0, inv.numel(), n_packed, float(alpha), stream_ptr())), "unpack_gather_group")


def scatter_grad_group(packed_grad, idx, dst_flats, accumulates, alpha=1.0):
    """packed_grad (K, ...) stacked packed gradients of K equally shaped parameters -> their K gradient slots, one launch per 16."""
    K = packed_grad.shape[0]
    n = idx.numel()
    assert packed_grad.dtype == torch.float32 and packed_grad.is_contiguous() and packed_grad.numel() == K * n
    for g0 in range(0, K, 16):
        g1 = min(K, g0 + 16)
        ptrs = (C.c_void_p * (g1 - g0))(*[None if d is None else d.data_ptr() for d in dst_flats[g0:g1]])
        accs = (C.c_int32 * (g1 - g0))(*[int(bool(a)) for a in accumulates[g0:g1]])
        base = packed_grad.data_ptr() + g0 * n * 4
        check(profiler.launch("unpack_scatter", lambda: lib().pmoe_unpack_scatter_group(
            base, idx.data_ptr(), ptrs, accs, g1 - g0, n, float(alpha), stream_ptr())), "unpack_scatter_group")


def bump(tensors):
    """Tell torch that kernels wrote these tensors through raw pointers (derived caches key on `_version`)."""
    ts = [t for t in tensors if t is not None]
    if ts:
        torch.autograd.graph.increment_version(ts)
