"""General (training-capable) execution of the reference modules on the B200 kernels.

A forward pass records a static tape of backward closures; `TapeFunction` exposes the whole model
call as ONE torch.autograd.Function so the reference trainers' `loss.backward()`, `clip_grad_norm_`
and optimizers work unchanged on the real nn.Parameters. Inside the tape everything is our own
kernels: gradient accumulation at fan-out points is fused into the producing kernel (residual /
accumulate arguments), BatchNorm follows each layer's own `.training` flag (so a "frozen" U-Net that
the trainer flipped back to train mode keeps updating its running statistics, exactly like the
reference — punet.py:55 vs train_1.py:123), and frozen up-stream sub-networks do no backward work.
"""
import ctypes as C

import torch

from . import _lib, config, dp, nhwc, ops, packs, profiler
from ._lib import ACT, check, lib, stream_ptr, view4
from .nhwc import Act, dtype_code
from .ops import TAPS3, pad_ch


# ------------------------------------------------------------------------------------------------ tape
import contextlib
import os as _os0

MULTI_STREAM = _os0.environ.get("PMOE_MULTI_STREAM", "1") == "1"   # independent sub-networks (expert encoders) on side streams
MULTI_STREAM_MAX_PIXELS = int(_os0.environ.get("PMOE_MULTI_STREAM_MAX_BATCH", "160")) * 224 * 224   # ... when one sub-network's batch is small enough to leave SMs idle (B <= 160 at 224^2)
_SIDE_STREAMS = {}


def _side_stream(device, index):
    key = (device.index if device.index is not None else torch.cuda.current_device(), index)
    st = _SIDE_STREAMS.get(key)
    if st is None:
        st = _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return st


ACCUMULATE_IN_PLACE = True  # gradients of parameters that already have a .grad are added into it by the kernels themselves


class BnParamGrads(C.Structure):  # PmoeBnParamGrads
    _fields_ = [("dgamma", C.c_void_p), ("dbeta", C.c_void_p), ("n", C.c_int32), ("accumulate", C.c_int32)]


def _bn_pgrads(tape, bn, c):
    """Gradient slots of a BatchNorm's weight / bias for the apply kernel's side output -> (struct or None, params to mark done)."""
    pg, done, acc = BnParamGrads(), [], None
    for prm, field in ((bn.weight, "dgamma"), (bn.bias, "dbeta")):
        if prm is None or not prm.requires_grad:
            continue
        slot, existed = tape.pgrad_slot(prm)
        if acc is None:
            acc = existed
        elif acc != existed:      # one slot fresh, one not: make both accumulate
            if not existed:
                slot.zero_()
            acc = True
        setattr(pg, field, slot.data_ptr())
        done.append(prm)
    if not done:
        return None, done
    pg.n, pg.accumulate = int(c), int(bool(acc))
    return pg, done


class Tape:
    def __init__(self, dtype, save):
        self.dtype = dtype
        self.save = save            # False: forward only (no_grad), nothing is kept for backward
        self.ops = []
        self.grads = {}             # id(Act) -> NHWC gradient tensor
        self.pgrads = {}            # id(param) -> fp32 gradient in the parameter's own layout
        self.params = {}            # id(param) -> param
        self.alive = []             # keeps Acts (and therefore their ids) alive until backward
        self.pending = {}           # id(param) -> gradient contributions still to come (data-parallel overlap)
        self.reported = set()
        self.bucketer = None        # pmoe_b200.dp.GradBucketer during a data-parallel backward
        self.presums = {}           # id(Act) -> (sum dy*[y>0], sum dy*y) reduced by the kernel that wrote the Act's gradient
        self.lazy = {}              # id(Act) -> (dy, gate, dmean): gradient dy*gate[n,c] + dmean[n,c] that is never stored (eca_op)
        self.pending_dgrad = {}     # id(Act) -> [(dy, packed dgrad weights, Act, ...)]: skinny grouped data gradients of one source, merged into ONE GEMM when the source's gradient is asked for
        self.raw_grads = {}         # id(Act) -> gradient of the RAW conv output behind the Act's BatchNorm + ReLU, already formed by the consumer (bn_relu_maxpool_op)
        self.arena = {}             # dtype -> [zeroed chunk, elements handed out]
        self.direct = set()         # id(param) whose gradient was accumulated straight into param.grad
        self.branch_stream = None   # side stream of the branch being recorded / replayed (None: the caller's stream)
        self.side_streams = []      # side streams this pass has forked work onto
        self.nbt = []               # BatchNorm num_batches_tracked buffers to bump at the end of the pass
        self.touched = []           # buffers written through raw pointers (running statistics): versions bumped at the end

    # one memset per chunk instead of one per accumulator: the statistics / weight-gradient accumulators of a pass are
    # slices of a few zero-filled chunks (every slice is handed out once, so it is still zero when its kernel runs)
    _ARENA_ELEMS = {torch.float64: 1 << 16, torch.float32: 1 << 24}

    def zeros(self, shape, dtype, device):
        n = 1
        for d in (shape if isinstance(shape, (tuple, list)) else (shape,)):
            n *= int(d)
        n_al = (n + 63) // 64 * 64  # 256-byte aligned slices
        key = (dtype, device, id(self.branch_stream))   # one arena per stream: a chunk is zero-filled on the stream that uses it
        ent = self.arena.get(key)
        if ent is None or ent[1] + n_al > ent[0].numel():
            ent = [torch.zeros(max(self._ARENA_ELEMS.get(dtype, 1 << 16), n_al), dtype=dtype, device=device), 0]
            self.arena[key] = ent
        out = ent[0][ent[1]:ent[1] + n].view(shape)
        ent[1] += n_al
        return out

    def expect(self, *params):
        """Forward-time announcement that a recorded backward closure will add_pgrad() to these parameters."""
        if not self.save:
            return
        for p in params:
            if p is not None and p.requires_grad:
                self.pending[id(p)] = self.pending.get(id(p), 0) + 1

    def record(self, fn):
        if self.save:
            self.ops.append((fn, self.branch_stream))

    @contextlib.contextmanager
    def branch(self, index, enable=True):
        """Run (and later replay the backward of) an independent sub-network on its own CUDA stream: the K expert encoders of a
        mixture share nothing but the input, and at small per-GPU batches their late stages are too small to fill 148 SMs one
        kernel at a time. Forward: the side stream waits for the caller's stream at entry; `join()` makes the caller's stream wait
        for every side stream. Backward: a closure recorded inside a branch replays on the same side stream, after the caller's
        stream has produced the gradients it consumes. Inside CUDA-graph capture the branches become parallel graph branches."""
        if not (MULTI_STREAM and enable):
            yield
            return
        main = torch.cuda.current_stream()
        side = _side_stream(main.device, index)
        side.wait_stream(main)
        if side not in self.side_streams:
            self.side_streams.append(side)
        prev, self.branch_stream = self.branch_stream, side
        try:
            with torch.cuda.stream(side):
                yield
        finally:
            self.branch_stream = prev

    def join(self):
        """The caller's stream waits for all side streams (end of the forked section of the forward pass)."""
        main = torch.cuda.current_stream()
        for s_ in self.side_streams:
            main.wait_stream(s_)

    def track(self, act):
        if self.save:
            self.alive.append(act)
        return act

    def grad_of(self, act):
        if self.pending_dgrad and id(act) in self.pending_dgrad:
            _flush_dgrad(self, id(act))
        g = self.grads.pop(id(act), None)
        if g is not None and self.branch_stream is not None:
            g.record_stream(self.branch_stream)   # may have been produced (allocated) on another stream: keep it until this one is done
        return g

    def pgrad_slot(self, p):
        """(flat fp32 buffer of p's gradient, whether it already holds a contribution). Under data parallelism the buffer
        is p's slot of the flat all-reduce bucket, so kernels write gradients where NCCL reads them."""
        k = id(p)
        if k in self.pgrads:
            return self.pgrads[k].view(-1), True
        self.params[k] = p
        if self.bucketer is not None:
            buf = self.bucketer.slot(p)
            self.pgrads[k] = buf.view(p.shape)
            return buf, False
        g = p.grad
        if ACCUMULATE_IN_PLACE and g is not None and g.dtype == torch.float32 and g.is_contiguous() and g.device == p.device:
            # the parameter already carries a gradient (micro-batch accumulation, captured graphs with static .grad): the kernels
            # add into it directly instead of handing autograd a second tensor to add (one read-modify-write of every gradient less)
            self.pgrads[k] = g
            self.direct.add(k)
            return g.view(-1), True
        full = torch.empty(p.shape, dtype=torch.float32, device=p.device)  # owns its storage: AccumulateGrad can take it without a copy
        self.pgrads[k] = full
        return full.view(-1), False

    def pgrad_done(self, p):
        """One announced contribution to p's gradient has been written."""
        k = id(p)
        left = self.pending.get(k, 0) - 1
        self.pending[k] = left
        if left == 0 and self.bucketer is not None:
            # last contribution: the bucket may leave for its all-reduce while the tape keeps running
            self.reported.add(k)
            self.bucketer.ready(p, None)

    def add_pgrad(self, p, g):
        if not p.requires_grad:
            return
        slot, existed = self.pgrad_slot(p)
        g = g.reshape(-1)
        if g.dtype == torch.float64 and g.is_contiguous() and g.is_cuda:
            check(profiler.launch("cvt_f64_f32", lambda: lib().pmoe_cvt_f64_f32(g.data_ptr(), slot.data_ptr(), g.numel(), int(existed),
                                                                               stream_ptr())), "cvt_f64_f32")
        elif existed:
            slot.add_(g)
        else:
            slot.copy_(g)
        self.pgrad_done(p)

    def backward(self):
        main = torch.cuda.current_stream()
        forked = []
        for fn, side in reversed(self.ops):
            if side is None:
                if forked:       # back on the caller's stream: everything the side streams produced is needed (or about to be freed)
                    for s_ in forked:
                        main.wait_stream(s_)
                    forked = []
                self.branch_stream = None
                fn()
            else:
                if side not in forked:
                    side.wait_stream(main)   # the gradients this branch consumes were produced on the caller's stream
                    forked.append(side)
                self.branch_stream = side
                with torch.cuda.stream(side):
                    fn()
        self.branch_stream = None
        for s_ in forked:
            main.wait_stream(s_)
        for key in list(self.pending_dgrad):   # (a source nobody asked for: nothing depends on it, the queue is just emptied)
            self.pending_dgrad.pop(key, None)
        self.ops = []
        self.alive = []
        self.grads = {}
        self.lazy = {}
        self.raw_grads = {}
        self.pending_dgrad = {}
        if self.bucketer is not None:  # gradients whose announced contributions did not all arrive
            for k in self.pgrads:
                if k not in self.reported:
                    self.reported.add(k)
                    self.bucketer.ready(self.params[k], None)

    def finish_forward(self):
        """End of the forward pass: the bookkeeping the kernels' raw-pointer writes owe to torch."""
        if self.nbt:
            torch._foreach_add_(self.nbt, 1)
            self.nbt = []
        if self.touched:
            packs.bump(self.touched)
            self.touched = []


def _new_act(tape, t, c, rg):
    a = Act(t, c)
    a.rg = bool(rg) and tape.save
    return tape.track(a)


def _rg(a):
    return getattr(a, "rg", False)


def _any_rg(params):
    return any(p is not None and p.requires_grad for p in params)


# ------------------------------------------------------------------------------------------------ weight packing
# ------------------------------------------------------------------------------------------------ raw launch helpers
def _mask_from_x(dz, x, act, fwd):
    """The ReLU mask can be recomputed from the pre-BN tensor with the forward's own scale/shift (one tensor read less per
    pass) when the fast contiguous-bf16 kernels apply."""
    return (fwd is not None and act == "relu" and x is not None and dz.dtype == torch.bfloat16 and dz.is_contiguous()
            and x.is_contiguous() and x.shape == dz.shape)


PRE_ACTS = ("hswish",)                       # activations whose derivative is not a function of their output
_PIECEWISE = ("relu6", "hswish", "hsigmoid")  # ... and those whose derivative the kernels can take at the recomputed pre-activation


def _pre_from_x(x, act, fwd):
    """MobileNet activations: derivative at the pre-activation fwd_scale*x + fwd_shift recomputed from the layer input (generic
    kernels, any dtype / strides); the saved output is then not read."""
    return fwd is not None and x is not None and act in _PIECEWISE


def _bn_bwd_reduce(tape, dz, z, x, act, mean, rstd, cpad, fwd=None):
    s1 = tape.zeros(cpad, torch.float64, dz.device)
    s2 = tape.zeros(cpad, torch.float64, dz.device) if x is not None else None
    mx = _mask_from_x(dz, x, act, fwd) or _pre_from_x(x, act, fwd)
    vdz = view4(dz)
    vz = view4(z) if (z is not None and not mx) else _lib.null_view()
    vx = view4(x) if x is not None else _lib.null_view()
    check(profiler.launch("bn_bwd_reduce", lambda: lib().pmoe_bn_bwd_reduce(
        C.byref(vdz), C.byref(vz), C.byref(vx), dtype_code(dz), ACT[act], _lib.ptr(mean), _lib.ptr(rstd), s1.data_ptr(),
        _lib.ptr(s2), _lib.ptr(fwd[0] if mx else None), _lib.ptr(fwd[1] if mx else None), stream_ptr()),
        io=(dz, None if (z is None or mx) else z, x)), "bn_bwd_reduce")
    return s1, s2


FUSE_BN_CHAIN_SUMS = True  # tests switch it off to compare against the separate reduce pass
FUSE_ECA_BN_BWD = _os0.environ.get("PMOE_FUSE_ECA_BN_BWD", "1") != "0"   # ECA gate behind conv+BN+ReLU: its input gradient is never stored
GATE_GRAD_FROM_WGRAD = True  # the stem's first ECA gate takes its gradient from the conv's per-image weight gradient (see eca_op)


def _bn_bwd_apply_sums(tape, dz, z, x, act, mean, rstd, gamma, s1, s2, inv_n, dx, fwd, pgrads=None):
    """bn_bwd_apply (batch statistics, ReLU) that also reduces sum dx*[x>0] and sum dx*x for the upstream BatchNorm whose ReLU
    output x is. Returns the two fp64 sums, or None (nothing launched) when the tensors do not qualify."""
    cp = dz.shape[3]
    if not (FUSE_BN_CHAIN_SUMS and dz.dtype == torch.bfloat16 and act == "relu" and dz.is_contiguous() and x.is_contiguous()
            and dx.is_contiguous() and (z is None or z.is_contiguous()) and 256 % (cp // 8) == 0):
        return None
    mx = _mask_from_x(dz, x, act, fwd)
    if z is None and not mx:
        return None
    n1 = tape.zeros(cp, torch.float64, dz.device)
    n2 = tape.zeros(cp, torch.float64, dz.device)
    vdz, vx, vdx = view4(dz), view4(x), view4(dx)
    vz = view4(z) if (z is not None and not mx) else _lib.null_view()
    check(profiler.launch("bn_bwd_apply", lambda: lib().pmoe_bn_bwd_apply_sums(
        C.byref(vdz), C.byref(vz), C.byref(vx), dtype_code(dz), ACT[act], _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(gamma),
        _lib.ptr(s1), _lib.ptr(s2), float(inv_n), C.byref(vdx), _lib.ptr(fwd[0] if mx else None), _lib.ptr(fwd[1] if mx else None),
        n1.data_ptr(), n2.data_ptr(), None if pgrads is None else C.byref(pgrads), stream_ptr()), io=(dz, None if mx else z, x, dx)),
        "bn_bwd_apply_sums")
    return n1, n2


def _bn_bwd_apply(dz, z, x, act, mean, rstd, gamma, s1, s2, inv_n, batch_stats, dx, dres, acc_dres, fwd=None, pgrads=None):
    mx = (_mask_from_x(dz, x, act, fwd) and (dx is None or dx.is_contiguous()) and (dres is None or dres.is_contiguous())) \
        or _pre_from_x(x, act, fwd)
    vdz = view4(dz)
    vz = view4(z) if (z is not None and not mx) else _lib.null_view()
    vx = view4(x) if x is not None else _lib.null_view()
    vdx = view4(dx) if dx is not None else _lib.null_view()
    vdr = view4(dres) if dres is not None else _lib.null_view()
    check(profiler.launch("bn_bwd_apply", lambda: lib().pmoe_bn_bwd_apply(
        C.byref(vdz), C.byref(vz), C.byref(vx), dtype_code(dz), ACT[act], _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(gamma),
        _lib.ptr(s1), _lib.ptr(s2), float(inv_n), int(batch_stats), C.byref(vdx), C.byref(vdr), int(acc_dres),
        _lib.ptr(fwd[0] if mx else None), _lib.ptr(fwd[1] if mx else None), None if pgrads is None else C.byref(pgrads), stream_ptr()),
        io=(dz, None if (z is None or mx) else z, x, dx, dres, dres if (acc_dres and dres is not None) else None)), "bn_bwd_apply")


def _axpy(src, dst, alpha=1.0, bcast=None, accumulate=False):
    vs = view4(src) if src is not None else _lib.null_view()
    vd = view4(dst)
    check(profiler.launch("axpy", lambda: lib().pmoe_axpy(
        C.byref(vs), C.byref(vd), dtype_code(dst), float(alpha), _lib.ptr(bcast), 0 if bcast is None else bcast.stride(0),
        int(accumulate), stream_ptr()), io=(src, dst, dst if accumulate else None)), "axpy")


def _accumulate(tape, act, g):
    """grads[act] += g (g already shaped like act.t)."""
    k = id(act)
    tape.presums.pop(k, None)  # sums reduced from an earlier, now incomplete gradient
    if k in tape.grads:
        _axpy(g, tape.grads[k], 1.0, None, True)
    else:
        tape.grads[k] = g


def _grad_buffer(tape, act):
    """Existing gradient buffer of `act` (to be accumulated into) or a fresh one; returns (tensor, existed)."""
    k = id(act)
    if k in tape.grads:
        tape.presums.pop(k, None)
        return tape.grads[k], True
    g = torch.empty(act.t.shape, dtype=act.t.dtype, device=act.t.device)
    tape.grads[k] = g
    return g, False


# ------------------------------------------------------------------------------------------------ conv (+BN +act +residual)
class Src:
    """One physical K-source of a conv: `t` (NHWC tensor/view) belongs to Act `act`; it carries the weight's input
    channels [cin0, cin0 + sum(logical)) laid out as `lay` = [(logical, padded)...]; `sel(g)` maps a gradient buffer
    of `act` to the view that corresponds to `t` (identity unless `t` is a spatial parity view)."""
    __slots__ = ("act", "t", "cin0", "lay", "sel")

    def __init__(self, act, t=None, cin0=0, lay=None, sel=None):
        self.act = act
        self.t = act.t if t is None else t
        self.cin0 = cin0
        self.lay = [(act.c, self.t.shape[3])] if lay is None else list(lay)
        self.sel = sel if sel is not None else (lambda g: g)

    @property
    def cpad(self):
        return self.t.shape[3]

    @property
    def nlog(self):
        return sum(g[0] for g in self.lay)


def concat_sources(acts, layouts=None):
    """Virtual channel concat: source i carries the next block of the weight's input channels."""
    out, c0 = [], 0
    for i, a in enumerate(acts):
        s = Src(a, None, c0, None if layouts is None else layouts[i])
        out.append(s)
        c0 += s.nlog
    return out


def default_segdefs(nsrc, ksize):
    pad = ksize // 2
    return [(i, r - pad, s - pad, r, s) for r in range(ksize) for s in range(ksize) for i in range(nsrc)]


def stride2_sources(x, ksize):
    """Sources + segment definitions of a stride-2 conv (k=3,p=1 or k=1,p=0) as stride-1 taps over the four
    spatial parity views of x (ih = 2*oh - pad + r)."""
    views = {}
    srcs, segdefs = [], []

    def view(p, q):
        if (p, q) not in views:
            views[(p, q)] = len(srcs)
            srcs.append(Src(x, x.t[:, p::2, q::2, :], 0, None, (lambda g, p=p, q=q: g[:, p::2, q::2, :])))
        return views[(p, q)]

    if ksize == 1:
        segdefs.append((view(0, 0), 0, 0, 0, 0))
    else:
        par = {0: (1, -1), 1: (0, 0), 2: (1, 0)}  # tap r -> (row parity, row offset in that view)
        for r in range(3):
            for s in range(3):
                (p, dh), (q, dw) = par[r], par[s]
                segdefs.append((view(p, q), dh, dw, r, s))
    return srcs, segdefs


def _pack_cols(wf, src, r, s):
    """(cout, padded channels of src) block of the weight for tap (r, s)."""
    cols, ci = [], src.cin0
    for (gl, gp) in src.lay:
        blk = wf[:, ci:ci + gl, r, s]
        if gp > gl:
            blk = torch.nn.functional.pad(blk, (0, gp - gl))
        cols.append(blk)
        ci += gl
    return cols


def _pack_fwd(weight, srcs, segdefs, cop, dtype):
    """-> (packed [cop][K] operand, its packs.Pack): K enumerates (segment, padded channel of the segment's source)."""
    def build(wf):
        cols = []
        for (i, _, _, r, s) in segdefs:
            cols += _pack_cols(wf, srcs[i], r, s)
        wp = torch.cat(cols, dim=1)
        if cop > wp.shape[0]:
            wp = torch.nn.functional.pad(wp, (0, 0, 0, cop - wp.shape[0]))
        return wp
    key = "f|%s|%s|%d|%s" % (";".join("%d:%s" % (x.cin0, x.lay) for x in srcs), segdefs, cop, dtype)
    return packs.packed(_owner(weight), key, weight.shape, build, dtype)


def _owner(weight):
    return getattr(weight, "owner", weight)


def _wgrad_to_param(tape, dwp, pack, weight):
    """Packed fp32 weight gradient -> the parameter's (out, in, kh, kw) gradient slot (one scatter launch)."""
    p = _owner(weight)
    slot, existed = tape.pgrad_slot(p)
    packs.unpack_grads(dwp if dwp.is_contiguous() else dwp.contiguous(), pack, [slot], [existed])
    tape.pgrad_done(p)


def _group_to_params(tape, stacked, pack_list, params):
    """Stacked (K, ...) packed gradients of K equally shaped parameters -> their gradient slots in one scatter launch."""
    slots, accs, done = [], [], []
    for prm in params:
        if prm is None or not prm.requires_grad:
            slots.append(None)
            accs.append(False)
            continue
        slot, existed = tape.pgrad_slot(prm)
        slots.append(slot)
        accs.append(existed)
        done.append(prm)
    packs.unpack_grads(stacked.contiguous(), pack_list[0], slots, accs)
    for prm in done:
        tape.pgrad_done(prm)


def _pack_dgrad(weight, src, segs_i, co_pad, dtype):
    """rows = physical channels of `src`, K = (segment of this source, padded cout)."""
    def build(wf):
        cout = wf.shape[0]
        blocks = []
        for (_, _, _, r, s) in segs_i:
            blk = torch.cat(_pack_cols(wf, src, r, s), dim=1).t()  # (phys channels, cout)
            if co_pad > cout:
                blk = torch.nn.functional.pad(blk, (0, co_pad - cout))
            blocks.append(blk)
        wd = torch.cat(blocks, dim=1)
        rows_pad = ops.cout_padded(wd.shape[0])
        if rows_pad > wd.shape[0]:
            wd = torch.nn.functional.pad(wd, (0, 0, 0, rows_pad - wd.shape[0]))
        return wd
    key = "d|%d:%s|%s|%d|%s" % (src.cin0, src.lay, segs_i, co_pad, dtype)
    return packs.packed(_owner(weight), key, weight.shape, build, dtype)[0]


FUSE_STRIDE2_DGRAD = True  # tests switch it off to compare against the one-launch-per-parity-view form
_S2_TAP = {(0, 0): 1, (1, 0): 2, (1, 1): 0}  # (output parity, shift into dy) -> kernel tap of a k3/s2/p1 conv


def _stride2_dgrad_fused(tape, x, weight, dy, co_pad, ck_d, n, oh, ow, flops, tag):
    """Data gradient of a 3x3 / stride-2 / pad-1 conv in ONE launch (instead of one per spatial parity view of x, each a
    small-K, store-bound GEMM): dx[2i+a, 2j+b, :] = sum over the shifts (dh, dw) in {0,1}^2 of dy[i+dh, j+dw, :] @ W[tap(a,dh),
    tap(b,dw)] — a GEMM over the dy grid with K = 4 shifts x Cout and N = 4 parities x Cin (7 of the 16 blocks are zero), whose
    column blocks (2a, 2a+1) are the (n, i, j, 2*Cin) view of the output rows 2i+a (the ConvTranspose2d store path)."""
    cs = x.cpad
    shifts = ((0, 0), (0, 1), (1, 0), (1, 1))

    def build(wf):
        cout, cin = wf.shape[0], wf.shape[1]
        rows = []
        for a in range(2):
            for b in range(2):
                cols = []
                for (dh, dw) in shifts:
                    blk = torch.zeros(cs, co_pad, dtype=torch.float32, device=wf.device)
                    if (a, dh) in _S2_TAP and (b, dw) in _S2_TAP:
                        blk[:cin, :cout] = wf[:, :, _S2_TAP[(a, dh)], _S2_TAP[(b, dw)]].t()
                    cols.append(blk)
                rows.append(torch.cat(cols, 1))
        return torch.cat(rows, 0)  # (4*cs, 4*co_pad)
    wd4 = packs.packed(_owner(weight), "s2d4|%d|%d|%s" % (cs, co_pad, dy.dtype), weight.shape, build, dy.dtype)[0]
    g = torch.empty(x.t.shape, dtype=dy.dtype, device=dy.device)   # every pixel of every parity is written
    tape.grads[id(x)] = g
    tape.presums.pop(id(x), None)
    dsegs = [(0, dh, dw, 0, co_pad // ck_d) for (dh, dw) in shifts]
    rows = g.view(n, oh, 2, ow, 2 * cs)
    ops.conv([dy], wd4, dsegs, ck_d, rows[:, :, 0], out_extra=[rows[:, :, 1]], out_cols=2 * cs, flops=flops, tag="dgrad " + tag)


def _bn_finalize(tape, bn, ssum, ssq, count, cout):
    """Batch statistics -> (mean, rstd, scale, shift) and the running-statistics update of a training BatchNorm2d."""
    track = bn.track_running_stats and bn.running_mean is not None
    mom = 0.1 if bn.momentum is None else bn.momentum
    mean, rstd, scale, shift = nhwc.bn_finalize(ssum, ssq, count, cout, bn.weight.detach(), bn.bias.detach(), bn.eps, mom,
                                                bn.running_mean if track else None, bn.running_var if track else None)
    if track:
        tape.touched += [bn.running_mean, bn.running_var]
        if bn.num_batches_tracked is not None:
            tape.nbt.append(bn.num_batches_tracked)
    return mean, rstd, scale, shift


def _bn_tail_forward(tape, bn, raw, ssum, ssq, count, cout, cstore, act, residual, out, pool=None, out_stats=None):
    mean, rstd, scale, shift = _bn_finalize(tape, bn, ssum, ssq, count, cout)
    z_t = out if out is not None else torch.empty(raw.shape, dtype=raw.dtype, device=raw.device)
    fused = False
    if (pool is not None or out_stats is not None) and residual is None:
        # the statistics of the output that the next layer needs (ECA / avg-pool sums, a directly following BatchNorm) ride on
        # the pass that writes it
        fused = nhwc.affine_act_stats(raw, scale[:cstore], shift[:cstore], act, z_t, pool, 0, out_stats)
    if not fused:
        nhwc.affine_act(raw, scale[:cstore], shift[:cstore], act, None if residual is None else residual.t, out=z_t)
        if pool is not None:
            nhwc.channel_sums(z_t, out=pool)
        if out_stats is not None:
            v = view4(z_t)
            check(profiler.launch("channel_stats", lambda: lib().pmoe_channel_stats(
                C.byref(v), dtype_code(z_t), out_stats[0].data_ptr(), out_stats[1].data_ptr(), stream_ptr()), io=(z_t,)), "channel_stats")
            if len(out_stats) > 2:
                out_stats[2].add_((z_t > 0).sum(dim=(0, 1, 2)).double())
    # (scale, shift) let the backward recompute the ReLU mask from `raw` instead of reading z (not with a residual add)
    return z_t, mean, rstd, ((scale, shift) if residual is None else None)


def _padded_gamma(bn, cstore):
    """BatchNorm weight as a [cstore] fp32 vector: the parameter itself when no channel padding is needed."""
    g = bn.weight.detach()
    if g.numel() == cstore and g.dtype == torch.float32 and g.is_contiguous():
        return g
    return ops.pad_vec(g, cstore, 0.0)


def _eval_affine(bn, bias, cout, cop):
    if bn is not None:
        def build():
            scale = ops.pad_vec(bn.weight.detach().float() / torch.sqrt(bn.running_var.float() + bn.eps), cop, 0.0)
            shift = ops.pad_vec(bn.bias.detach().float() - bn.running_mean.float() * scale[:cout], cop, 0.0)
            return scale, shift
        from .infer import cached
        return cached(bn, "evalaff%d" % cop, [bn.weight, bn.bias, bn.running_mean, bn.running_var], build)
    if bias is not None:
        return None, _padded_bias(bias, cop)
    return None, None


def _padded_bias(bias, cop, repeat=1):
    """Bias as the [cop] (x repeat) fp32 shift vector of the GEMM epilogue: the parameter itself when nothing needs padding,
    otherwise a packed operand refreshed with the weights (no per-call fill + copy launches)."""
    b = bias.detach()
    if repeat == 1 and b.numel() == cop and b.dtype == torch.float32 and b.is_contiguous():
        return b
    n = b.numel()
    return packs.packed(bias, "bias|%d|%d" % (cop, repeat), bias.shape,
                        lambda v: torch.nn.functional.pad(v, (0, cop - n)).repeat(repeat), torch.float32)[0]


def conv_op(tape, srcs, weight, bias=None, bn=None, act=None, residual=None, want_pool=False, ksize=3, layouts=None,
            segdefs=None, out=None, pool_out=None, pool_stride=0, out_hw=None, tag="", want_out_stats=False, stride2_of=None):
    """conv / linear layer over `srcs` (list of Act = virtual channel concat, or list of Src), followed by BatchNorm
    (batch statistics when bn.training, folded running statistics otherwise), optional residual add and activation.
    Returns (Act z, pool_sum or None)."""
    dt = tape.dtype
    if act in PRE_ACTS and tape.save and not (bn is not None and bn.training):
        y, _ = conv_op(tape, srcs, weight, bias, bn, None, residual, False, ksize, layouts, segdefs, None, None, 0, out_hw, tag, False, stride2_of)
        z = act_op(tape, y, act, tag)
        pool = None
        if want_pool:
            pool = nhwc.channel_sums(z.t, out=pool_out) if pool_out is not None else nhwc.channel_sums(z.t)
        return z, pool
    if not isinstance(srcs[0], Src):
        srcs = concat_sources(srcs, layouts)
    dev = srcs[0].t.device
    cout = weight.shape[0]
    cop, cstore = ops.cout_padded(cout), pad_ch(cout)
    if segdefs is None:
        segdefs = default_segdefs(len(srcs), ksize)
    phys = [x.cpad for x in srcs]
    ck = ops.choose_ck(phys)
    segs = [(i, dh, dw, 0, phys[i] // ck) for (i, dh, dw, _, _) in segdefs]
    wp, wpk = _pack_fwd(weight, srcs, segdefs, cop, dt)
    assert bias is None or bn is None, "conv_op: a conv followed by BatchNorm carries no bias in the reference (basics.py:51-55)"
    n, h, w, _ = srcs[0].t.shape
    if out_hw is not None:
        h, w = out_hw
    flops = 2.0 * n * h * w * cout * sum(srcs[i].nlog for (i, _, _, _, _) in segdefs)
    bn_train = bn is not None and bn.training
    rg_in = any(_rg(x.act) or getattr(x.act, "gate_grad", None) is not None for x in srcs) or _any_rg([_owner(weight), bias]) \
        or (bn is not None and _any_rg([bn.weight, bn.bias])) or (residual is not None and _rg(residual))
    pool = None
    if want_pool:
        pool = pool_out if pool_out is not None else torch.zeros(n, cop, dtype=torch.float32, device=dev)
    raw = mean = rstd = gamma_p = scale = fwd_aff = out_stats = None
    src_ts = [x.t for x in srcs]
    if bn_train:
        raw = torch.empty(n, h, w, cstore, dtype=dt, device=dev)
        ssum = tape.zeros(cop, torch.float64, dev)
        ssq = tape.zeros(cop, torch.float64, dev)
        if cop <= 512:
            ops.conv(src_ts, wp, segs, ck, raw, stat_sum=ssum, stat_sqsum=ssq, flops=flops, tag=tag)
        else:  # the epilogue keeps its per-channel partial sums in shared memory for at most 512 channels (resnet50: 1024/2048)
            ops.conv(src_ts, wp, segs, ck, raw, flops=flops, tag=tag)
            vr = view4(raw)
            check(profiler.launch("channel_stats", lambda: lib().pmoe_channel_stats(C.byref(vr), dtype_code(raw), ssum.data_ptr(),
                                                                                    ssq.data_ptr(), stream_ptr()), io=(raw,)), "channel_stats")
        if want_out_stats:
            out_stats = (tape.zeros(cstore, torch.float64, dev), tape.zeros(cstore, torch.float64, dev))
            if act == "relu" and residual is None and not want_pool and dt == torch.bfloat16:
                out_stats += (tape.zeros(cstore, torch.float64, dev),)   # count of outputs > 0 (closed-form backward of this BatchNorm)
        z_t, mean, rstd, fwd_aff = _bn_tail_forward(tape, bn, raw, ssum, ssq, n * h * w, cout, cstore, act, residual, out,
                                                    pool=pool if want_pool else None, out_stats=out_stats)
        gamma_p = _padded_gamma(bn, cstore)
    else:
        scale, shift = _eval_affine(bn, bias, cout, cop)
        z_t = out if out is not None else torch.empty(n, h, w, cstore, dtype=dt, device=dev)
        ops.conv(src_ts, wp, segs, ck, z_t, scale=scale, shift=shift, act=act,
                 residual=None if residual is None else residual.t, pool_sum=pool, pool_stride=pool_stride, flops=flops, tag=tag)
    z = _new_act(tape, z_t, cout, rg_in)
    z.stats = out_stats
    z.bn_relu = bool(bn_train and act == "relu" and residual is None and fwd_aff is not None)  # z = relu(BN(raw)), nothing added
    if z.bn_relu:
        z.bn_ctx = (bn, raw, mean, rstd, fwd_aff, gamma_p, cout, cstore)
    if not (tape.save and rg_in):
        return z, pool

    def backward():
        lazy = tape.lazy.pop(id(z), None)
        if lazy is not None and (id(z) in tape.grads or not (bn_train and fwd_aff is not None and act == "relu" and id(z) in tape.presums)):
            # something else contributed to (or invalidated the sums of) this gradient: store the lazy part after all
            g, existed = _grad_buffer(tape, z)
            vd, vg = view4(lazy[0]), view4(g)
            check(profiler.launch("eca_bwd_apply", lambda: lib().pmoe_eca_bwd_apply(
                C.byref(vd), dtype_code(lazy[0]), lazy[1].data_ptr(), lazy[1].stride(0), lazy[2].data_ptr(), lazy[2].stride(0), C.byref(vg),
                int(existed), stream_ptr()), io=(lazy[0], g, g if existed else None)), "eca_bwd_apply")
            lazy = None
        rawg = tape.raw_grads.pop(id(z), None)
        if rawg is not None and (lazy is not None or id(z) in tape.grads):
            raise RuntimeError("pmoe_b200 conv_op: the gradient of the raw conv output was formed by the consumer, but another "
                               "gradient of the activation exists as well (" + tag + ")")
        dz = None if (lazy is not None or rawg is not None) else tape.grad_of(z)
        if dz is None and lazy is None and rawg is None:
            return
        z_saved = z_t if act not in (None, "none") else None
        dres, acc_dres = None, False
        if residual is not None and _rg(residual):
            dres, acc_dres = _grad_buffer(tape, residual)
        # gradient w.r.t. the raw conv output
        dy = rawg if rawg is not None else torch.empty(n, h, w, cstore, dtype=dt, device=dev)
        if rawg is not None:
            pass   # BatchNorm backward (and its affine gradients) done by the consumer: bn_relu_maxpool_op
        elif bn_train:
            pres = tape.presums.pop(id(z), None)
            if pres is not None and fwd_aff is not None and act == "relu":
                # the kernel that wrote dz also reduced sum dz*[z>0] and sum dz*z; with z = relu(scale*raw + shift):
                # sum dz*m*raw = (sum dz*z - shift * sum dz*m) / scale, and the kernels' second sum is rstd*(that - mean*first)
                sc, sh = fwd_aff[0][:cstore].double(), fwd_aff[1][:cstore].double()
                s1 = pres[0]
                sraw = torch.where(sc != 0, (pres[1] - sh * s1) / torch.where(sc != 0, sc, torch.ones_like(sc)), torch.zeros_like(sc))
                s2 = rstd[:cstore].double() * (sraw - mean[:cstore].double() * s1)
            else:
                s1, s2 = _bn_bwd_reduce(tape, dz, z_saved, raw, act, mean, rstd, cstore, fwd=fwd_aff)
            pg, pdone = _bn_pgrads(tape, bn, cout)   # d weight = s2, d bias = s1: written by the apply kernel
            if lazy is not None:
                vl, vr, vo = view4(lazy[0]), view4(raw), view4(dy)
                s1c, s2c = s1.contiguous(), s2.contiguous()
                check(profiler.launch("eca_bn_bwd_apply", lambda: lib().pmoe_eca_bn_bwd_apply(
                    C.byref(vl), C.byref(vr), lazy[1].data_ptr(), lazy[1].stride(0), lazy[2].data_ptr(), lazy[2].stride(0),
                    fwd_aff[0].data_ptr(), fwd_aff[1].data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma_p.data_ptr(), s1c.data_ptr(),
                    s2c.data_ptr(), 1.0 / (n * h * w), C.byref(vo), None if pg is None else C.byref(pg), stream_ptr()),
                    io=(lazy[0], raw, dy)), "eca_bn_bwd_apply")
            else:
                _bn_bwd_apply(dz, z_saved, raw, act, mean, rstd, gamma_p, s1, s2, 1.0 / (n * h * w), 1, dy, dres, acc_dres, fwd=fwd_aff,
                              pgrads=pg)
            for prm in pdone:
                tape.pgrad_done(prm)
        else:
            if bn is not None and _any_rg([bn.weight, bn.bias]):
                raise NotImplementedError("pmoe_b200: gradients of BatchNorm affine parameters in eval mode are not supported")
            if bias is not None and bias.requires_grad:
                s1, _ = _bn_bwd_reduce(tape, dz, z_saved, None, act, None, None, cstore)
                tape.add_pgrad(bias, s1[:cout])
            _bn_bwd_apply(dz, z_saved, None, act, None, None, None if scale is None else scale[:cstore].contiguous(), None, None,
                          0.0, 0, dy, dres, acc_dres)
        gate_hooks = [x for x in srcs if getattr(x.act, "gate_grad", None) is not None]
        if gate_hooks:
            # per-image weight gradient (grouped mode of the tensor-core kernel): its sum over the images is dW, and contracted
            # with W over (co, tap) it is the gradient of the ECA gate that scaled this conv's input
            if len(srcs) != 1 or dt != torch.bfloat16:
                raise RuntimeError("pmoe_b200 conv_op: a gate-gradient source must be the conv's only source on the tensor-core path")
            dwn = tape.zeros((n, cop, wp.shape[1]), torch.float32, dev)
            ops.conv_wgrad(src_ts, segs, ck, dy, dwn, flops=flops, tag="wgrad " + tag)
            cpx = srcs[0].cpad
            dg = (dwn.view(n, cop, -1, cpx) * wp.float().view(1, cop, -1, cpx)).sum(dim=(1, 2))
            gate_hooks[0].act.gate_grad(dg)
            if weight.requires_grad:
                _wgrad_to_param(tape, dwn.sum(dim=0), wpk, weight)
        elif weight.requires_grad:
            dwp = tape.zeros((cop, wp.shape[1]), torch.float32, dev)
            ops.conv_wgrad(src_ts, segs, ck, dy, dwp, flops=flops, tag="wgrad " + tag)
            _wgrad_to_param(tape, dwp, wpk, weight)
        # data gradients: one launch per physical source whose owner needs them
        ck_d = ops.choose_ck([cstore])
        pre = {id(x.act): (id(x.act) in tape.grads) for x in srcs}
        if (FUSE_STRIDE2_DGRAD and stride2_of is not None and ksize == 3 and len(srcs) == 4 and _rg(stride2_of) and id(stride2_of) not in tape.grads
                and dt == torch.bfloat16 and not config.FORCE_SIMT and stride2_of.cpad % 64 == 0 and stride2_of.c == stride2_of.cpad):
            _stride2_dgrad_fused(tape, stride2_of, weight, dy, cstore, ck_d, n, h, w, flops, tag)
            return
        for i, src in enumerate(srcs):
            if not _rg(src.act):
                continue
            # taps in REVERSE order: the data gradient reads dy at (-dh, -dw), so this enumerates the offsets in the
            # canonical (-1,-1) .. (1,1) order the resident / halo tensor-core kernels recognise
            segs_i = [sd for sd in segdefs if sd[0] == i][::-1]
            if not segs_i:
                continue
            wd = _pack_dgrad(weight, src, segs_i, cstore, dt)
            dsegs = [(0, -dh, -dw, 0, cstore // ck_d) for (_, dh, dw, _, _) in segs_i]
            k = id(src.act)
            tape.presums.pop(k, None)  # this launch changes (or creates) the gradient: sums reduced earlier no longer describe it
            if k not in tape.grads:
                g = torch.empty(src.act.t.shape, dtype=dt, device=dev)
                if src.t.shape != src.act.t.shape:
                    g.zero_()  # spatial parity views: pixels that no view of this conv covers must read as zero
                tape.grads[k] = g
            g = tape.grads[k]
            gv = src.sel(g)
            ops.conv([dy], wd, dsegs, ck_d, gv, residual=gv if pre[k] else None,
                     flops=2.0 * n * h * w * cout * src.nlog * len(segs_i), tag="dgrad " + tag)

    if bn_train:
        tape.expect(bn.weight, bn.bias)
    else:
        tape.expect(bias)
    tape.expect(_owner(weight))
    tape.record(backward)
    return z, pool


def act_op(tape, x, act, tag=""):
    """Stand-alone activation (Hardswish after a Linear / an eval-mode conv): y = act(x); backward at the pre-activation x."""
    dev = x.t.device
    cp = x.cpad
    ones = torch.ones(cp, dtype=torch.float32, device=dev)
    zeros = torch.zeros(cp, dtype=torch.float32, device=dev)
    y = nhwc.affine_act(x.t, ones, zeros, act)
    ya = _new_act(tape, y, x.c, _rg(x))
    if tape.save and _rg(x):
        def backward():
            dz = tape.grad_of(ya)
            if dz is None:
                return
            g, existed = _grad_buffer(tape, x)
            tmp = g if not existed else torch.empty_like(g)
            _bn_bwd_apply(dz, None if act in _PIECEWISE else y, x.t, act, None, None, None, None, None, 0.0, 0, tmp, None, False,
                          fwd=(ones, zeros))
            if existed:
                _axpy(tmp, g, 1.0, None, True)
        tape.record(backward)
    return ya


def dwconv_op(tape, x, conv, bn, act, tag=""):
    """Depthwise nn.Conv2d(c, c, k, stride, (k-1)//2, groups=c, bias=False) + BatchNorm2d + activation: the middle layer of
    torchvision's MobileNetV2 / V3 InvertedResidual (backbone.py:75-104). Returns the Act."""
    dt, dev = tape.dtype, x.t.device
    k, stride, pad = int(conv.kernel_size[0]), int(conv.stride[0]), int(conv.padding[0])
    if conv.groups != conv.in_channels or conv.in_channels != conv.out_channels or conv.bias is not None or conv.dilation[0] != 1:
        raise RuntimeError("pmoe_b200 dwconv_op: depthwise, bias-free, undilated convolutions only (got %r)" % (conv,))
    c, cp = x.c, x.cpad
    n, h, w, _ = x.t.shape
    oh, ow = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    weight = conv.weight

    def build(wf):   # (c, 1, k, k) -> [k*k][cp], tap-major
        return torch.nn.functional.pad(wf[:, 0].permute(1, 2, 0).reshape(k * k, c), (0, cp - c))
    wpk, went = packs.packed(weight, "dw|%d|%d" % (k, cp), weight.shape, build, torch.float32)
    raw = torch.empty(n, oh, ow, cp, dtype=dt, device=dev)
    vx, vr = view4(x.t), view4(raw)
    io_b = (x.t, raw)
    check(profiler.launch("dwconv", lambda: lib().pmoe_dwconv_fwd(C.byref(vx), wpk.data_ptr(), cp, C.byref(vr), dtype_code(raw), k, stride, pad,
                                                                   stream_ptr()), io=io_b), "dwconv_fwd")
    bn_train = bn is not None and bn.training
    rg = _rg(x) or weight.requires_grad or (bn is not None and _any_rg([bn.weight, bn.bias]))
    mean = rstd = gamma_p = fwd_aff = scale = shift = None
    if bn_train:
        ssum, ssq = tape.zeros(cp, torch.float64, dev), tape.zeros(cp, torch.float64, dev)
        check(profiler.launch("channel_stats", lambda: lib().pmoe_channel_stats(C.byref(vr), dtype_code(raw), ssum.data_ptr(), ssq.data_ptr(),
                                                                                stream_ptr()), io=(raw,)), "channel_stats")
        z_t, mean, rstd, fwd_aff = _bn_tail_forward(tape, bn, raw, ssum, ssq, n * oh * ow, c, cp, act, None, None)
        gamma_p = _padded_gamma(bn, cp)
    else:
        scale, shift = _eval_affine(bn, None, c, cp)
        if scale is None:
            scale = torch.ones(cp, dtype=torch.float32, device=dev)
            shift = torch.zeros(cp, dtype=torch.float32, device=dev)
        z_t = nhwc.affine_act(raw, scale, shift, act)
    z = _new_act(tape, z_t, c, rg)
    if not (tape.save and rg):
        return z

    def backward():
        dz = tape.grad_of(z)
        if dz is None:
            return
        z_saved = z_t if act not in (None, "none") else None
        dy = torch.empty(n, oh, ow, cp, dtype=dt, device=dev)
        if bn_train:
            s1, s2 = _bn_bwd_reduce(tape, dz, z_saved, raw, act, mean, rstd, cp, fwd=fwd_aff)
            pg, pdone = _bn_pgrads(tape, bn, c)
            _bn_bwd_apply(dz, z_saved, raw, act, mean, rstd, gamma_p, s1, s2, 1.0 / (n * oh * ow), 1, dy, None, False, fwd=fwd_aff, pgrads=pg)
            for prm in pdone:
                tape.pgrad_done(prm)
        else:
            if bn is not None and _any_rg([bn.weight, bn.bias]):
                raise NotImplementedError("pmoe_b200: gradients of BatchNorm affine parameters in eval mode are not supported")
            pre = act in _PIECEWISE
            _bn_bwd_apply(dz, None if pre else z_saved, raw if pre else None, act, None, None, scale[:cp].contiguous(), None, None, 0.0, 0,
                          dy, None, False, fwd=(scale, shift) if pre else None)
        vdy = view4(dy)
        if weight.requires_grad:
            dwp = tape.zeros((k * k, cp), torch.float32, dev)
            check(profiler.launch("dwconv_wgrad", lambda: lib().pmoe_dwconv_wgrad(C.byref(vx), C.byref(vdy), dwp.data_ptr(), cp, dtype_code(dy), k,
                                                                                 stride, pad, stream_ptr()), io=(x.t, dy)), "dwconv_wgrad")
            _wgrad_to_param(tape, dwp, went, weight)
        if _rg(x):
            g, existed = _grad_buffer(tape, x)
            vg = view4(g)
            check(profiler.launch("dwconv_dgrad", lambda: lib().pmoe_dwconv_dgrad(C.byref(vdy), wpk.data_ptr(), cp, C.byref(vg), dtype_code(g), k,
                                                                                 stride, pad, int(existed), stream_ptr()),
                                  io=(dy, g, g if existed else None)), "dwconv_dgrad")

    if bn_train:
        tape.expect(bn.weight, bn.bias)
    tape.expect(weight)
    tape.record(backward)
    return z


def se_op(tape, se, x, tag=""):
    """torchvision SqueezeExcitation (MobileNetV3): x * hardsigmoid(fc2(relu(fc1(avgpool(x))))), fc1 / fc2 = 1x1 Conv2d with bias.
    The two small linears run on the tape as 1x1 convs over the pooled (1,1,N,C) vector; the gate scales x per (n, c)."""
    n, h, w, cp = x.t.shape
    sums = nhwc.channel_sums(x.t)
    inter = InterRepr(tape, x, sums)
    a = feature_act(tape, inter)
    hid, _ = conv_op(tape, [a], se.fc1.weight, se.fc1.bias, None, "relu", ksize=1, tag=tag + ".fc1")
    gact, _ = conv_op(tape, [hid], se.fc2.weight, se.fc2.bias, None, "hsigmoid", ksize=1, tag=tag + ".fc2")
    gate = gact.t.view(n, gact.cpad).float()
    if gact.cpad != cp:
        raise RuntimeError("pmoe_b200 se_op: gate / activation channel padding mismatch (%d vs %d)" % (gact.cpad, cp))
    y = nhwc.scale_channels(x.t, gate)
    rg = _rg(x) or _rg(gact)
    ya = _new_act(tape, y, x.c, rg)
    if tape.save and rg:
        def backward():
            dy = tape.grad_of(ya)
            if dy is None:
                return
            if _rg(gact):   # d gate[n, c] = sum over pixels of dy * x
                dgate = tape.zeros((n, cp), torch.float64, dy.device)
                va, vb = view4(dy), view4(x.t)
                check(profiler.launch("prod_channel_sums", lambda: lib().pmoe_prod_channel_sums(
                    C.byref(va), C.byref(vb), dtype_code(dy), dgate.data_ptr(), dgate.stride(0), stream_ptr()), io=(dy, x.t)), "prod_channel_sums")
                seed_vec(tape, gact, dgate[:, :gact.c].float())
            if _rg(x):
                g, existed = _grad_buffer(tape, x)
                if existed:
                    tmp = nhwc.scale_channels(dy, gate)
                    _axpy(tmp, g, 1.0, None, True)
                else:
                    nhwc.scale_channels(dy, gate, out=g)
        tape.record(backward)
    return ya


def maxpool_op(tape, x, k, stride, pad):
    need_bwd = tape.save and _rg(x)
    if need_bwd:
        y, idx = nhwc.maxpool(x, k, stride, pad, want_idx=True)
    else:
        y = nhwc.maxpool(x, k, stride, pad)
    ya = _new_act(tape, y.t, x.c, _rg(x))
    if need_bwd:
        def backward():
            dy = tape.grad_of(ya)
            if dy is None:
                return
            g, existed = _grad_buffer(tape, x)
            vdy, vdx = view4(dy), view4(g)
            check(profiler.launch("maxpool_bwd", lambda: lib().pmoe_maxpool_bwd_idx(
                C.byref(vdy), idx.data_ptr(), C.byref(vdx), dtype_code(g), k, stride, pad, int(existed), stream_ptr()),
                io=(dy, idx, g, g if existed else None)), "maxpool_bwd")
        tape.record(backward)
    return ya


def conv_transpose_op(tape, up, x, tag=""):
    """nn.ConvTranspose2d(k=2, s=2): four 1x1 GEMMs into the pixel-shuffle views of the output."""
    dt, dev = tape.dtype, x.t.device
    weight, bias = up.weight, up.bias
    cin, cout = weight.shape[0], weight.shape[1]
    cop, cstore = ops.cout_padded(cout), pad_ch(cout)
    n, h, w, cp = x.t.shape
    ck = ops.choose_ck([cp])
    segs = ops.conv_segments([(0, 0)], [cp], ck)
    out = torch.empty(n, 2 * h, 2 * w, cstore, dtype=dt, device=dev)
    wf = weight.detach().float()
    flops = 2.0 * n * h * w * cin * cout
    views = [out[:, a::2, b::2, :] for a in range(2) for b in range(2)]
    if cstore % 64 == 0 and dt == torch.bfloat16 and not config.FORCE_SIMT:
        # one GEMM with N = 4*Cout; column blocks (2a, 2a+1) land in the (n, h, w, 2*C) view of output rows 2h+a
        def build4(wf):
            w4 = torch.zeros(4 * cstore, cp, dtype=torch.float32, device=wf.device)
            for q, (a, b) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
                w4[q * cstore:q * cstore + cout, :cin] = wf[:, :, a, b].t()
            return w4
        wp4 = packs.packed(weight, "T4|%d|%d|%s" % (cp, cstore, dt), weight.shape, build4, dt)[0]
        shift4 = _padded_bias(bias, cstore, repeat=4)
        rows = out.view(n, h, 2, w, 2 * cstore)
        ops.conv([x.t], wp4, segs, ck, rows[:, :, 0], shift=shift4, out_extra=[rows[:, :, 1]], out_cols=2 * cstore, flops=4 * flops,
                 tag="convT " + tag)
    else:
        shift = ops.pad_vec(bias.detach(), cop, 0.0)
        for q, (a, b) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
            wp = ops.pack_conv_weight(wf[:, :, a, b].t().reshape(cout, cin, 1, 1), [x.c], [cp], [(0, 0)], cop, dt)
            ops.conv([x.t], wp, segs, ck, views[q], shift=shift, flops=flops, tag="convT " + tag)
    rg = _rg(x) or _any_rg([weight, bias])
    ya = _new_act(tape, out, cout, rg)
    if tape.save and rg:
        def backward():
            dy = tape.grad_of(ya)
            if dy is None:
                return
            if bias.requires_grad:
                s1, _ = _bn_bwd_reduce(tape, dy, None, None, None, None, None, cstore)
                tape.add_pgrad(bias, s1[:cout])
            if weight.requires_grad:
                gw = torch.empty(cin, cout, 2, 2, dtype=torch.float32, device=dev)
                for a in range(2):
                    for b in range(2):
                        dwp = tape.zeros((cop, cp), torch.float32, dev)
                        ops.conv_wgrad([x.t], segs, ck, dy[:, a::2, b::2, :], dwp, flops=flops, tag="wgrad convT " + tag)
                        gw[:, :, a, b] = dwp[:cout, :cin].t()
                tape.add_pgrad(weight, gw)
            if _rg(x):
                # dx[n,h,w,ci] = sum_{a,b,co} dy[n,2h+a,2w+b,co] * W[ci,co,a,b]: 4 parity views of dy as 4 K segments
                views = [dy[:, a::2, b::2, :] for a in range(2) for b in range(2)]
                ck_d = ops.choose_ck([cstore])
                dsegs = [(q, 0, 0, 0, cstore // ck_d) for q in range(4)]
                blocks = []
                for a in range(2):
                    for b in range(2):
                        blk = wf[:, :, a, b]  # (cin, cout)
                        blocks.append(torch.nn.functional.pad(blk, (0, cstore - cout)))
                wd = torch.cat(blocks, 1)  # (cin, 4*cstore)
                rows = ops.cout_padded(cp)
                wd = torch.nn.functional.pad(wd, (0, 0, 0, rows - cin)).to(dt).contiguous()
                g, existed = _grad_buffer(tape, x)
                ops.conv(views, wd, dsegs, ck_d, g, residual=g if existed else None, flops=4 * flops, tag="dgrad convT " + tag)
        tape.expect(bias, weight)
        tape.record(backward)
    return ya


def eca_op(tape, eca_mod, x, layout=None, pool_in=None):
    """EfficientBlock: out = x * sigmoid(conv1d(mean_hw(x))). layout = (groups, logical, slot)."""
    w = eca_mod.conv.weight
    n, h, wd, cp = x.t.shape
    groups, gl, gs = layout if layout is not None else (1, x.c, cp)
    sums = pool_in if pool_in is not None else nhwc.channel_sums(x.t)
    gate = nhwc.eca_gate(sums, h * wd, w.detach(), groups, gl, gs)
    y = nhwc.scale_channels(x.t, gate)
    rg = _rg(x) or w.requires_grad
    if (GATE_GRAD_FROM_WGRAD and tape.save and not _rg(x) and w.requires_grad and tape.dtype == torch.bfloat16 and not config.FORCE_SIMT
            and groups == 1):
        # ECA on a network INPUT (the stem's first gate): x needs no gradient, and the gate's gradient
        #   d gate[n, c] = sum_p d(x*gate)[n, p, c] * x[n, p, c]
        # equals sum over (co, tap) of W[co, c, tap] * dW_n[co, c, tap] / gate[n, c], with dW_n the consuming conv's PER-IMAGE weight
        # gradient (it sees x * gate as its input). The conv computes dW_n anyway (grouped mode of the weight-gradient kernel),
        # so the 64 -> 16-channel data-gradient launch and two passes over the input-sized tensors are not needed at all.
        ya = _new_act(tape, y, x.c, False)

        def gate_grad(dg_gated):   # (N, Cpad) fp32: sum over (co, tap) of W * dW_n
            dgate = (dg_gated.double() / gate.double().clamp_min(1e-30)).contiguous()
            dmean = torch.empty(n, cp, dtype=torch.float32, device=dgate.device)
            dw = tape.zeros(w.numel(), torch.float64, dgate.device)
            wf = w.detach().reshape(-1)
            check(profiler.launch("eca_gate_bwd", lambda: lib().pmoe_eca_gate_bwd(
                dgate.data_ptr(), dgate.stride(0), gate.data_ptr(), gate.stride(0), sums.data_ptr(), sums.stride(0), n,
                1.0 / float(h * wd), wf.data_ptr(), wf.numel(), groups, gl, gs, dmean.data_ptr(), dmean.stride(0),
                dw.data_ptr(), stream_ptr())), "eca_gate_bwd")
            tape.add_pgrad(w, dw)
        ya.gate_grad = gate_grad
        tape.expect(w)
        return ya
    ya = _new_act(tape, y, x.c, rg)
    if tape.save and rg:
        def backward():
            dy = tape.grad_of(ya)
            if dy is None:
                return
            dgate = tape.zeros((n, cp), torch.float64, dy.device)  # cancelling sums: kept in fp64
            va, vb = view4(dy), view4(x.t)
            lazy_ok = False
            if (FUSE_ECA_BN_BWD and FUSE_BN_CHAIN_SUMS and _rg(x) and getattr(x, "bn_relu", False) and id(x) not in tape.grads
                    and dy.dtype == torch.bfloat16 and dy.is_contiguous() and x.t.is_contiguous() and x.t.dtype == torch.bfloat16
                    and 128 % (cp // 8) == 0 and groups == 1 and gate.stride(0) % 4 == 0):
                p1 = tape.zeros((n, cp), torch.float64, dy.device)
                m0 = tape.zeros((n, cp), torch.float64, dy.device)
                rc = profiler.launch("eca_bn_bwd_sums", lambda: lib().pmoe_eca_bn_bwd_sums(
                    C.byref(va), C.byref(vb), p1.data_ptr(), dgate.data_ptr(), m0.data_ptr(), dgate.stride(0), stream_ptr()), io=(dy, x.t))
                if rc == 0:
                    lazy_ok = True
                elif rc != -2:
                    check(rc, "eca_bn_bwd_sums")
                else:
                    profiler.uncount()
            if not lazy_ok:
                check(profiler.launch("prod_channel_sums", lambda: lib().pmoe_prod_channel_sums(
                    C.byref(va), C.byref(vb), dtype_code(dy), dgate.data_ptr(), dgate.stride(0), stream_ptr()), io=(dy, x.t)), "prod_channel_sums")
            dmean = torch.empty(n, cp, dtype=torch.float32, device=dy.device)
            dw = tape.zeros(w.numel(), torch.float64, dy.device)
            wf = w.detach().reshape(-1)
            check(profiler.launch("eca_gate_bwd", lambda: lib().pmoe_eca_gate_bwd(
                dgate.data_ptr(), dgate.stride(0), gate.data_ptr(), gate.stride(0), sums.data_ptr(), sums.stride(0), n,
                1.0 / float(h * wd), wf.data_ptr(), wf.numel(), groups, gl, gs, dmean.data_ptr(), dmean.stride(0),
                dw.data_ptr(), stream_ptr())), "eca_gate_bwd")
            tape.add_pgrad(w, dw)
            if _rg(x):
                if lazy_ok:
                    # x = relu(BN(raw)) of the conv upstream, this is its whole gradient so far and it is affine in dy per (image,
                    # channel): hand (dy, gate, dmean) and the BatchNorm's two backward sums to that layer instead of storing it
                    sm = sums[:, :cp].double()
                    gt = gate[:, :cp].double()
                    dmd = dmean.double()                         # the per-pixel constant of the gradient (d mean / HW)
                    n1 = (gt * p1 + dmd * m0).sum(dim=0)          # sum dc1 * [c1 > 0]
                    n2 = (gt * dgate + dmd * sm).sum(dim=0)       # sum dc1 * c1
                    tape.presums[id(x)] = (n1, n2)
                    tape.lazy[id(x)] = (dy, gate, dmean)
                    return
                g, existed = _grad_buffer(tape, x)
                vd, vg = view4(dy), view4(g)
                if (FUSE_BN_CHAIN_SUMS and getattr(x, "bn_relu", False) and not existed and dy.dtype == torch.bfloat16
                        and dy.is_contiguous() and x.t.is_contiguous() and 256 % (cp // 8) == 0):
                    # x = relu(BN(raw)) of the conv upstream and this is its whole gradient: that layer's backward sums ride along
                    n1 = tape.zeros(cp, torch.float64, dy.device)
                    n2 = tape.zeros(cp, torch.float64, dy.device)
                    vx = view4(x.t)
                    check(profiler.launch("eca_bwd_apply", lambda: lib().pmoe_eca_bwd_apply_sums(
                        C.byref(vd), dtype_code(dy), gate.data_ptr(), gate.stride(0), dmean.data_ptr(), dmean.stride(0), C.byref(vg),
                        C.byref(vx), n1.data_ptr(), n2.data_ptr(), stream_ptr()), io=(dy, g, x.t)), "eca_bwd_apply_sums")
                    tape.presums[id(x)] = (n1, n2)
                    return
                check(profiler.launch("eca_bwd_apply", lambda: lib().pmoe_eca_bwd_apply(
                    C.byref(vd), dtype_code(dy), gate.data_ptr(), gate.stride(0), dmean.data_ptr(), dmean.stride(0), C.byref(vg),
                    int(existed), stream_ptr()), io=(dy, g, g if existed else None)), "eca_bwd_apply")
        tape.expect(w)
        tape.record(backward)
    return ya


# ------------------------------------------------------------------------------------------------ module runners
def conv3_block(tape, seq, srcs, want_pool=False, tag=""):
    y, _ = conv_op(tape, srcs, seq[0].weight, None, seq[1], "relu", tag=tag + ".0")
    return conv_op(tape, [y], seq[3].weight, None, seq[4], "relu", want_pool=want_pool, tag=tag + ".3")


def unet(tape, net, x, out=None, out_pool=None, pool_stride=0, want_inter=False, tag="unet"):
    """UNet.forward (unet.py:50-95). Returns (Act logits, pooled bottleneck (N,512) fp32 or None)."""
    n, h, w, _ = x.t.shape
    if h % 16 or w % 16:
        raise RuntimeError("pmoe_b200 UNet needs H and W divisible by 16 (got %dx%d)" % (h, w))
    x1, _ = conv3_block(tape, net.dwn_1, [x], tag=tag + ".dwn_1")
    x2, _ = conv3_block(tape, net.dwn_2, [maxpool_op(tape, x1, 2, 2, 0)], tag=tag + ".dwn_2")
    x3, _ = conv3_block(tape, net.dwn_3, [maxpool_op(tape, x2, 2, 2, 0)], tag=tag + ".dwn_3")
    x4, _ = conv3_block(tape, net.dwn_4, [maxpool_op(tape, x3, 2, 2, 0)], tag=tag + ".dwn_4")
    x5, pool5 = conv3_block(tape, net.dwn_5, [maxpool_op(tape, x4, 2, 2, 0)], want_pool=want_inter, tag=tag + ".dwn_5")
    y = x5
    for i, (up, fwd, skip) in enumerate(((net.up_1, net.up_forw_1, x4), (net.up_2, net.up_forw_2, x3),
                                         (net.up_3, net.up_forw_3, x2), (net.up_4, net.up_forw_4, x1)), start=1):
        u = conv_transpose_op(tape, up, y, tag=tag + ".up_%d" % i)
        y, _ = conv3_block(tape, fwd, [skip, u], tag=tag + ".up_forw_%d" % i)
    logits, _ = conv_op(tape, [y], net.out.weight, net.out.bias, None, None, ksize=1, out=out, want_pool=out_pool is not None,
                        pool_out=out_pool, pool_stride=pool_stride, tag=tag + ".out")
    inter = None
    if want_inter:
        inter = InterRepr(tape, x5, pool5)
    return logits, inter


def concat_op(tape, acts):
    """torch.cat(acts, 1) as ONE physical NHWC tensor (needed where an op — the ECA gates of UNetECA, unet.py:165-178 —
    acts across the channel boundary of the concatenation). Every part must have an unpadded channel count (multiple of 16)."""
    for a in acts:
        if a.c != a.cpad:
            raise RuntimeError("pmoe_b200 concat_op: every part needs a multiple of 16 channels (got %d)" % a.c)
    n, h, w, _ = acts[0].t.shape
    ctot = sum(a.c for a in acts)
    t = torch.empty(n, h, w, ctot, dtype=acts[0].t.dtype, device=acts[0].t.device)
    c0 = 0
    for a in acts:
        _axpy(a.t, t[..., c0:c0 + a.c], 1.0, None, False)
        c0 += a.c
    ya = _new_act(tape, t, ctot, any(_rg(a) for a in acts))
    if tape.save and _rg(ya):
        def backward():
            g = tape.grad_of(ya)
            if g is None:
                return
            c0 = 0
            for a in acts:
                if _rg(a):
                    _accumulate_copy(tape, a, g[..., c0:c0 + a.c])
                c0 += a.c
        tape.record(backward)
    return ya


def unet_eca(tape, net, x, want_inter=False, tag="unet_eca"):
    """UNetECA.forward (unet.py:140-185). Returns (Act logits, InterRepr of the dwn_5 output or None)."""
    n, h, w, _ = x.t.shape
    if h % 16 or w % 16:
        raise RuntimeError("pmoe_b200 UNetECA needs H and W divisible by 16 (got %dx%d)" % (h, w))
    x1, _ = conv3_block(tape, net.dwn_1, [x], tag=tag + ".dwn_1")
    x2, _ = conv3_block(tape, net.dwn_2, [maxpool_op(tape, x1, 2, 2, 0)], tag=tag + ".dwn_2")
    x3, _ = conv3_block(tape, net.dwn_3, [maxpool_op(tape, x2, 2, 2, 0)], tag=tag + ".dwn_3")
    x4, _ = conv3_block(tape, net.dwn_4, [maxpool_op(tape, x3, 2, 2, 0)], tag=tag + ".dwn_4")
    p4 = eca_op(tape, net.eca_0, maxpool_op(tape, x4, 2, 2, 0))
    x5, pool5 = conv3_block(tape, net.dwn_5, [p4], want_pool=want_inter, tag=tag + ".dwn_5")
    y = x5
    for i, skip in enumerate((x4, x3, x2, x1), start=1):
        u = conv_transpose_op(tape, getattr(net, "up_%d" % i), y, tag=tag + ".up_%d" % i)
        cat = eca_op(tape, getattr(net, "eca_%d" % i), concat_op(tape, [skip, u]))  # skip first (unet.py:164)
        y, _ = conv3_block(tape, getattr(net, "up_forw_%d" % i), [cat], tag=tag + ".up_forw_%d" % i)
    logits, _ = conv_op(tape, [y], net.out.weight, net.out.bias, None, None, ksize=1, tag=tag + ".out")
    return logits, (InterRepr(tape, x5, pool5) if want_inter else None)


class InterRepr:
    """adaptive_avg_pool2d(x_5, 1).flatten(1) (unet.py:89-92) with its backward into x_5."""

    def __init__(self, tape, x5, pool):
        self.tape, self.x5 = tape, x5
        self.hw = x5.t.shape[1] * x5.t.shape[2]
        self.value = pool[:, :x5.c] / float(self.hw)

    def backward(self, g):
        """g: (N, C) fp32 gradient of the pooled vector."""
        if not _rg(self.x5):
            return
        gb = torch.zeros(self.x5.t.shape[0], self.x5.cpad, dtype=torch.float32, device=g.device)
        gb[:, :self.x5.c] = g.float() / float(self.hw)
        buf, existed = _grad_buffer(self.tape, self.x5)
        _axpy(None, buf, 1.0, gb, existed)


def eca_conv_block(tape, blk, x, layout=None, pool_in=None, tag="eca_block", want_out_stats=False):
    """EfficientConvBlock (basics.py:80-135). layout = (groups, logical, slot) of x's channel axis."""
    xs = eca_op(tape, blk.layer1.eca1, x, layout, pool_in)
    lay = None if layout is None else [[(layout[1], layout[2])] * layout[0]]
    c1, pool64 = conv_op(tape, [xs], blk.layer1.conv1[0].weight, None, blk.layer1.conv1[1], "relu", want_pool=True,
                         layouts=lay, tag=tag + ".conv1")
    c1s = eca_op(tape, blk.layer2.eca2, c1, None, pool64)
    y, _ = conv_op(tape, [c1s], blk.layer2.conv2[0].weight, None, blk.layer2.conv2[1], "relu", tag=tag + ".conv2",
                   want_out_stats=want_out_stats)
    return y


# ------------------------------------------------------------------------------------------------ ResNet-18 with ECA stem
def bn_act_op(tape, bn, x, act="relu", tag=""):
    """Stand-alone BatchNorm2d (+activation) on an activation that is not a fresh conv output (ResNet bn1)."""
    dt, dev = tape.dtype, x.t.device
    n, h, w, cp = x.t.shape
    c = x.c
    rg = _rg(x) or _any_rg([bn.weight, bn.bias])
    bn_train = bn.training
    mean = rstd = gamma_p = scale = fwd_aff = None
    if bn_train:
        if x.stats is not None and x.stats[0].numel() == cp:
            ssum, ssq = x.stats[0], x.stats[1]    # reduced by the kernel that wrote x (conv_op(want_out_stats=True))
        else:
            ssum = tape.zeros(cp, torch.float64, dev)
            ssq = tape.zeros(cp, torch.float64, dev)
            v = view4(x.t)
            check(profiler.launch("channel_stats", lambda: lib().pmoe_channel_stats(C.byref(v), dtype_code(x.t), ssum.data_ptr(),
                                                                                    ssq.data_ptr(), stream_ptr()), io=(x.t,)), "channel_stats")
        z_t, mean, rstd, fwd_aff = _bn_tail_forward(tape, bn, x.t, ssum, ssq, n * h * w, c, cp, act, None, None)
        gamma_p = _padded_gamma(bn, cp)
    else:
        scale, shift = _eval_affine(bn, None, c, cp)
        z_t = nhwc.affine_act(x.t, scale, shift, act)
    z = _new_act(tape, z_t, c, rg)
    if tape.save and rg:
        def backward():
            dz = tape.grad_of(z)
            if dz is None:
                return
            zs = z_t if act not in (None, "none") else None
            g, existed = _grad_buffer(tape, x)
            tmp = g if not existed else torch.empty_like(g)
            if bn_train:
                s1, s2 = _bn_bwd_reduce(tape, dz, zs, x.t, act, mean, rstd, cp, fwd=fwd_aff)
                pg, pdone = _bn_pgrads(tape, bn, c)
                nxt = None
                if getattr(x, "bn_relu", False) and not existed:
                    # x = relu(BN(raw)) of the conv just upstream and this is its only gradient so far: reduce that layer's
                    # backward sums while dx is in registers (invalidated if anything is accumulated into dx later)
                    nxt = _bn_bwd_apply_sums(tape, dz, zs, x.t, act, mean, rstd, gamma_p, s1, s2, 1.0 / (n * h * w), tmp, fwd_aff, pgrads=pg)
                if nxt is not None:
                    tape.presums[id(x)] = nxt
                else:
                    _bn_bwd_apply(dz, zs, x.t, act, mean, rstd, gamma_p, s1, s2, 1.0 / (n * h * w), 1, tmp, None, False, fwd=fwd_aff,
                                  pgrads=pg)
                for prm in pdone:
                    tape.pgrad_done(prm)
            else:
                if _any_rg([bn.weight, bn.bias]):
                    raise NotImplementedError("pmoe_b200: gradients of BatchNorm affine parameters in eval mode are not supported")
                _bn_bwd_apply(dz, zs, None, act, None, None, scale, None, None, 0.0, 0, tmp, None, False)
            if existed:
                _axpy(tmp, g, 1.0, None, True)
        if bn_train:
            tape.expect(bn.weight, bn.bias)
        tape.record(backward)
    return z


def basic_block(tape, blk, x, stride, want_pool=False, tag=""):
    """torchvision BasicBlock: relu(bn2(conv2(relu(bn1(conv1(x))))) + identity) with optional 1x1/s2 downsample. The
    downsample is recorded BEFORE conv1, so in the backward conv1's data gradient comes first and can write the whole of
    dx in one launch (`_stride2_dgrad_fused`); the downsample then accumulates into its parity view."""
    idt = x
    if blk.downsample is not None:
        if stride == 1:
            idt, _ = conv_op(tape, [x], blk.downsample[0].weight, None, blk.downsample[1], None, ksize=1, tag=tag + ".down")
        else:
            srcs, segdefs = stride2_sources(x, 1)
            idt, _ = conv_op(tape, srcs, blk.downsample[0].weight, None, blk.downsample[1], None, segdefs=segdefs,
                             out_hw=(srcs[0].t.shape[1], srcs[0].t.shape[2]), tag=tag + ".down")
    if stride == 1:
        y, _ = conv_op(tape, [x], blk.conv1.weight, None, blk.bn1, "relu", tag=tag + ".conv1")
    else:
        srcs, segdefs = stride2_sources(x, 3)
        oh, ow = srcs[0].t.shape[1], srcs[0].t.shape[2]
        y, _ = conv_op(tape, srcs, blk.conv1.weight, None, blk.bn1, "relu", segdefs=segdefs, out_hw=(oh, ow), tag=tag + ".conv1",
                       stride2_of=x)
    return conv_op(tape, [y], blk.conv2.weight, None, blk.bn2, "relu", residual=idt, want_pool=want_pool, tag=tag + ".conv2")


def bottleneck_block(tape, blk, x, stride, want_pool=False, tag=""):
    """torchvision Bottleneck (resnet50, v1.5: the stride sits on the 3x3): relu(bn3(conv3(relu(bn2(conv2(relu(bn1(conv1 x)))))))
    + identity)."""
    y, _ = conv_op(tape, [x], blk.conv1.weight, None, blk.bn1, "relu", ksize=1, tag=tag + ".conv1")
    if stride == 1:
        y, _ = conv_op(tape, [y], blk.conv2.weight, None, blk.bn2, "relu", tag=tag + ".conv2")
    else:
        srcs, segdefs = stride2_sources(y, 3)
        y, _ = conv_op(tape, srcs, blk.conv2.weight, None, blk.bn2, "relu", segdefs=segdefs,
                       out_hw=(srcs[0].t.shape[1], srcs[0].t.shape[2]), tag=tag + ".conv2", stride2_of=y)
    idt = x
    if blk.downsample is not None:
        if stride == 1:
            idt, _ = conv_op(tape, [x], blk.downsample[0].weight, None, blk.downsample[1], None, ksize=1, tag=tag + ".down")
        else:
            srcs, segdefs = stride2_sources(x, 1)
            idt, _ = conv_op(tape, srcs, blk.downsample[0].weight, None, blk.downsample[1], None, segdefs=segdefs,
                             out_hw=(srcs[0].t.shape[1], srcs[0].t.shape[2]), tag=tag + ".down")
    return conv_op(tape, [y], blk.conv3.weight, None, blk.bn3, "relu", residual=idt, want_pool=want_pool, ksize=1, tag=tag + ".conv3")


def resnet18_eca(tape, net, x, tag="backbone"):
    """ResNet._forward_impl with conv1 := EfficientConvBlock up to the global average pool (backbone.py:48-72; resnet18/34
    BasicBlock or resnet50 Bottleneck stages) -> InterRepr (B, 512 * expansion). `backbone_features` adds the fc."""
    if x.t.shape[1] % 32 or x.t.shape[2] % 32:
        raise RuntimeError("pmoe_b200 ResNet backbone needs H and W divisible by 32 (got %dx%d)" % (x.t.shape[1], x.t.shape[2]))
    stem = eca_conv_block(tape, net.conv1, x, tag=tag + ".conv1", want_out_stats=net.bn1.training)  # bn1 follows directly
    return resnet18_after_stem(tape, net, stem, tag)


FUSE_STEM_TAIL = _os0.environ.get("PMOE_FUSE_STEM_TAIL", "1") != "0"   # tests switch it off to compare with the separate launches
FUSE_STEM_UP = _os0.environ.get("PMOE_FUSE_STEM_UP", "1") != "0"       # ... and its backward continued through the conv + BN + ReLU in front


def bn_relu_maxpool_op(tape, bn, x, tag=""):
    """torchvision's bn1 -> relu -> MaxPool2d(3, 2, 1) behind the ResNet stem (backbone.py:57-61) as three launches of
    csrc/stem_tail.cu instead of five: the normalised full-resolution tensor and its gradient are never stored. Falls back to
    bn_act_op + maxpool_op for eval-mode statistics, fp32 tapes and geometries the kernels do not take."""
    n, h, w, cp = x.t.shape
    if not (FUSE_STEM_TAIL and bn.training and tape.dtype == torch.bfloat16 and x.t.dtype == torch.bfloat16 and x.t.is_contiguous()
            and h % 2 == 0 and w % 2 == 0 and cp % 8 == 0 and 256 % (cp // 8) == 0):
        return maxpool_op(tape, bn_act_op(tape, bn, x, "relu", tag=tag), 3, 2, 1)
    dev, c = x.t.device, x.c
    rg = _rg(x) or _any_rg([bn.weight, bn.bias])
    npos = None
    if x.stats is not None and x.stats[0].numel() == cp:
        ssum, ssq = x.stats[0], x.stats[1]
        npos = x.stats[2] if len(x.stats) > 2 else None
    else:
        ssum = tape.zeros(cp, torch.float64, dev)
        ssq = tape.zeros(cp, torch.float64, dev)
        v = view4(x.t)
        check(profiler.launch("channel_stats", lambda: lib().pmoe_channel_stats(C.byref(v), dtype_code(x.t), ssum.data_ptr(),
                                                                                ssq.data_ptr(), stream_ptr()), io=(x.t,)), "channel_stats")
    mean, rstd, scale, shift = _bn_finalize(tape, bn, ssum, ssq, n * h * w, c)
    scale, shift = scale[:cp], shift[:cp]
    p_t = torch.empty(n, h // 2, w // 2, cp, dtype=x.t.dtype, device=dev)
    x_at_max = torch.empty_like(p_t)
    idx = torch.empty(p_t.shape, dtype=torch.uint8, device=dev)
    vx, vp = view4(x.t), view4(p_t)
    check(profiler.launch("bn_relu_maxpool", lambda: lib().pmoe_bn_relu_maxpool_fwd(
        C.byref(vx), scale.data_ptr(), shift.data_ptr(), C.byref(vp), idx.data_ptr(), x_at_max.data_ptr(), stream_ptr()),
        io=(x.t, p_t, idx, x_at_max)), "bn_relu_maxpool_fwd")
    pa = _new_act(tape, p_t, c, rg)
    if tape.save and rg:
        gamma_p = _padded_gamma(bn, cp)

        def _stem_tail_backward_through_upstream(dp, ctx):
            """bn1 -> relu -> maxpool backward continued through the ReLU and BatchNorm of the conv that produced x: neither the
            gradient of x nor x itself is touched; the upstream conv_op receives the gradient of its RAW output (tape.raw_grads)."""
            ubn, uraw, umean, urstd, uaff, ugamma, ucout, _ = ctx
            N = float(n * h * w)
            s1 = tape.zeros(cp, torch.float64, dev)
            s2 = tape.zeros(cp, torch.float64, dev)
            e1 = tape.zeros(cp, torch.float64, dev)
            vdp = view4(dp)
            check(profiler.launch("bn_relu_maxpool_bwd_reduce", lambda: lib().pmoe_bn_relu_maxpool_bwd_reduce(
                C.byref(vdp), x_at_max.data_ptr(), scale.data_ptr(), shift.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                s1.data_ptr(), s2.data_ptr(), e1.data_ptr(), stream_ptr()), io=(dp, x_at_max)), "bn_relu_maxpool_bwd_reduce")
            # dx = A*d + B*x + C per channel (d: the pooled gradient routed to its argmax and masked by this layer's ReLU), so the two
            # backward sums of the upstream BatchNorm are linear in sums that exist already:
            #   sum dx*[x>0] = A*sum d*[x>0] + B*sum x + C*count(x>0)        (x >= 0: x*[x>0] = x)
            #   sum dx*x     = A*sum d*x     + B*sum x^2 + C*sum x
            g64, r64, m64 = gamma_p[:cp].double(), rstd[:cp].double(), mean[:cp].double()
            c1, c2 = s1 / N, s2 / N
            A = g64 * r64
            Bq = -g64 * r64 * r64 * c2
            Cq = g64 * r64 * (r64 * c2 * m64 - c1)
            sdx = s2 / r64 + m64 * s1                                    # sum d*x back from sum d*xhat
            n1 = A * e1 + Bq * ssum[:cp] + Cq * npos[:cp]
            n2 = A * sdx + Bq * ssq[:cp] + Cq * ssum[:cp]
            # ... and through the upstream forward affine as conv_op does for sums handed down by a consumer (x = scale*raw + shift
            # wherever x > 0): sum dx*m*raw = (sum dx*x - shift*sum dx*m) / scale
            usc, ush = uaff[0][:cp].double(), uaff[1][:cp].double()
            sraw = torch.where(usc != 0, (n2 - ush * n1) / torch.where(usc != 0, usc, torch.ones_like(usc)), torch.zeros_like(usc))
            u1 = n1.contiguous()
            u2 = (urstd[:cp].double() * (sraw - umean[:cp].double() * n1)).contiguous()
            pg, pdone = _bn_pgrads(tape, bn, c)
            upg, updone = _bn_pgrads(tape, ubn, ucout)
            draw = torch.empty(uraw.shape, dtype=uraw.dtype, device=dev)
            vr, vd = view4(uraw), view4(draw)
            check(profiler.launch("bn2_relu_maxpool_bwd_apply", lambda: lib().pmoe_bn2_relu_maxpool_bwd_apply(
                C.byref(vdp), idx.data_ptr(), C.byref(vr), scale.data_ptr(), shift.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                gamma_p.data_ptr(), s1.data_ptr(), s2.data_ptr(), 1.0 / N, None if pg is None else C.byref(pg),
                uaff[0].data_ptr(), uaff[1].data_ptr(), umean.data_ptr(), urstd.data_ptr(), ugamma.data_ptr(), u1.data_ptr(), u2.data_ptr(),
                None if upg is None else C.byref(upg), C.byref(vd), stream_ptr()), io=(dp, idx, uraw, draw)), "bn2_relu_maxpool_bwd_apply")
            tape.raw_grads[id(x)] = draw
            for prm in pdone + updone:
                tape.pgrad_done(prm)

        def backward():
            dp = tape.grad_of(pa)
            if dp is None:
                return
            if not dp.is_contiguous():
                dp = dp.contiguous()
            ctx = getattr(x, "bn_ctx", None)
            if (FUSE_STEM_UP and FUSE_BN_CHAIN_SUMS and ctx is not None and npos is not None and _rg(x) and id(x) not in tape.grads
                    and id(x) not in tape.lazy and 128 % (cp // 8) == 0 and ctx[1] is not None and ctx[1].is_contiguous()
                    and ctx[1].shape == x.t.shape and ctx[7] == cp):
                _stem_tail_backward_through_upstream(dp, ctx)
                return
            g, existed = _grad_buffer(tape, x)
            tmp = g if (not existed and g.is_contiguous()) else torch.empty(x.t.shape, dtype=x.t.dtype, device=dev)
            s1 = tape.zeros(cp, torch.float64, dev)
            s2 = tape.zeros(cp, torch.float64, dev)
            vdp, vxx, vdx = view4(dp), view4(x.t), view4(tmp)
            check(profiler.launch("bn_relu_maxpool_bwd_reduce", lambda: lib().pmoe_bn_relu_maxpool_bwd_reduce(
                C.byref(vdp), x_at_max.data_ptr(), scale.data_ptr(), shift.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                s1.data_ptr(), s2.data_ptr(), None, stream_ptr()), io=(dp, x_at_max)), "bn_relu_maxpool_bwd_reduce")
            pg, pdone = _bn_pgrads(tape, bn, c)
            chain = getattr(x, "bn_relu", False) and not existed and tmp is g and FUSE_BN_CHAIN_SUMS
            n1 = tape.zeros(cp, torch.float64, dev) if chain else None
            n2 = tape.zeros(cp, torch.float64, dev) if chain else None
            check(profiler.launch("bn_relu_maxpool_bwd_apply", lambda: lib().pmoe_bn_relu_maxpool_bwd_apply(
                C.byref(vdp), idx.data_ptr(), C.byref(vxx), scale.data_ptr(), shift.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                gamma_p.data_ptr(), s1.data_ptr(), s2.data_ptr(), 1.0 / (n * h * w), C.byref(vdx), _lib.ptr(n1), _lib.ptr(n2),
                None if pg is None else C.byref(pg), stream_ptr()), io=(dp, idx, x.t, tmp)), "bn_relu_maxpool_bwd_apply")
            if chain:
                tape.presums[id(x)] = (n1, n2)
            for prm in pdone:
                tape.pgrad_done(prm)
            if tmp is not g:
                _axpy(tmp, g, 1.0, None, existed)
        tape.expect(bn.weight, bn.bias)
        tape.record(backward)
    return pa


def resnet18_after_stem(tape, net, stem, tag="backbone"):
    y = bn_relu_maxpool_op(tape, net.bn1, stem, tag=tag + ".bn1")
    pool = None
    layers = (net.layer1, net.layer2, net.layer3, net.layer4)
    for li, layer in enumerate(layers, start=1):
        for bi, blk in enumerate(layer):
            last = li == len(layers) and bi == len(layer) - 1
            block = bottleneck_block if hasattr(blk, "conv3") else basic_block
            y, pool = block(tape, blk, y, blk.stride, want_pool=last, tag="%s.layer%d.%d" % (tag, li, bi))
    return InterRepr(tape, y, pool)


def backbone_head(tape, net, inter, tag="backbone"):
    """Pooled features -> the 512-vector Act the heads consume: Identity for resnet18/34, Linear(2048, 512) for resnet50
    (backbone.py:66-69)."""
    a = feature_act(tape, inter)
    fc = getattr(net, "fc", None)
    if isinstance(fc, torch.nn.Linear):
        a = linear_op(tape, [a], fc, None, tag=tag + ".fc")
    return a


def backbone_features(tape, net, x, tag="backbone"):
    if hasattr(net, "tape_features"):   # MobileNetECA walks its own (torchvision) module tree
        return net.tape_features(tape, x, tag)
    return backbone_head(tape, net, resnet18_eca(tape, net, x, tag), tag)


# ------------------------------------------------------------------------------------------------ MLP heads
def vec_act(tape, value_f32, c, rg=False):
    """(B, c) fp32 tensor -> Act of shape (1,1,B,cpad) in the tape dtype (rows = the GEMM's M axis)."""
    B = value_f32.shape[0]
    cp = pad_ch(c)
    buf = torch.zeros(B, cp, dtype=torch.float32, device=value_f32.device)
    buf[:, :c] = value_f32
    t = torch.empty(1, 1, B, cp, dtype=tape.dtype, device=value_f32.device)
    _axpy(None, t.view(B, 1, 1, cp), 1.0, buf, False)
    return _new_act(tape, t, c, rg)


def feature_act(tape, inter):
    """InterRepr (pooled backbone features) -> Act usable as a linear-layer source, wired for backward."""
    a = vec_act(tape, inter.value, inter.x5.c, _rg(inter.x5))
    if tape.save and _rg(inter.x5):
        def backward():
            g = tape.grad_of(a)
            if g is not None:
                inter.backward(g.view(g.shape[2], g.shape[3])[:, :inter.x5.c])
        tape.record(backward)
    return a


DROPOUT_STEP = None  # optional device int64 step counter: masks are seeded from it (CUDA-graph friendly, see bench.py)


def dropout_op(tape, x, p, training):
    if not training or p <= 0.0:
        return x
    seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    y = torch.empty_like(x.t)
    n = x.t.numel()
    step_dev = DROPOUT_STEP

    def run(src, dst):
        if step_dev is not None:
            check(profiler.launch("dropout", lambda: lib().pmoe_dropout_dev(src.data_ptr(), dst.data_ptr(), dtype_code(dst), n, float(p), seed,
                                                                            step_dev.data_ptr(), stream_ptr())), "dropout")
        else:
            check(profiler.launch("dropout", lambda: lib().pmoe_dropout(src.data_ptr(), dst.data_ptr(), dtype_code(dst), n, float(p), seed,
                                                                        stream_ptr())), "dropout")

    run(x.t, y)
    ya = _new_act(tape, y, x.c, _rg(x))
    if tape.save and _rg(x):
        def backward():
            dy = tape.grad_of(ya)
            if dy is None:
                return
            g, existed = _grad_buffer(tape, x)
            tmp = g if not existed else torch.empty_like(g)
            run(dy, tmp)  # the same mask (same salt, same step counter) applied to the gradient
            if existed:
                _axpy(tmp, g, 1.0, None, True)
        tape.record(backward)
    return ya


def linear_op(tape, srcs, lin, act=None, out=None, tag=""):
    """nn.Linear over the virtual concat of `srcs` (Acts shaped (1,1,B,C)) as a 1x1 conv."""
    w4 = lin.weight.view(lin.weight.shape[0], lin.weight.shape[1], 1, 1)
    return conv_op(tape, srcs, _AsConv(lin, w4), lin.bias, None, act, ksize=1, out=out, tag=tag)[0]


class _AsConv:
    """Presents a Linear weight (out,in) as a (out,in,1,1) conv weight while gradients land on the real Parameter."""

    def __init__(self, lin, w4):
        self.lin, self.w4 = lin, w4
        self.shape = w4.shape

    @property
    def requires_grad(self):
        return self.lin.weight.requires_grad

    def detach(self):
        return self.w4.detach()

    @property
    def owner(self):  # the real Parameter: pack cache and gradients live there
        return self.lin.weight


def mlp(tape, seq, srcs, tag="mlp"):
    """make_mlp Sequential (basics.py:11-45): Linear [+BatchNorm1d] + act [+Dropout] ..., last Linear [+act]."""
    mods = list(seq)
    x_srcs = srcs
    i = 0
    y = None
    while i < len(mods):
        m = mods[i]
        if not isinstance(m, torch.nn.Linear):
            raise RuntimeError("pmoe_b200 mlp: unexpected layer order at %d: %r" % (i, m))
        j = i + 1
        bn = None
        act = None
        drop = None
        if j < len(mods) and isinstance(mods[j], torch.nn.BatchNorm1d):
            bn = mods[j]
            j += 1
        if j < len(mods) and isinstance(mods[j], (torch.nn.ReLU, torch.nn.ELU, torch.nn.Tanh, torch.nn.Sigmoid)):
            act = {torch.nn.ReLU: "relu", torch.nn.ELU: "elu", torch.nn.Tanh: "tanh", torch.nn.Sigmoid: "sigmoid"}[type(mods[j])]
            j += 1
        if j < len(mods) and isinstance(mods[j], torch.nn.Dropout):
            drop = mods[j]
            j += 1
        w4 = m.weight.view(m.weight.shape[0], m.weight.shape[1], 1, 1)
        y, _ = conv_op(tape, x_srcs, _AsConv(m, w4), m.bias, bn, act, ksize=1, tag="%s.%d" % (tag, i))
        if drop is not None:
            y = dropout_op(tape, y, drop.p, drop.training)
        x_srcs = [y]
        i = j
    return y


def vec_value(act):
    """Act (1,1,B,cpad) -> fp32 (B, c)."""
    return act.t.view(act.t.shape[2], act.t.shape[3])[:, :act.c].float()


def seed_vec(tape, act, g):
    """Install a (B, c) fp32 gradient for a vector Act."""
    if g is None or not _rg(act):
        return
    B, cp = act.t.shape[2], act.t.shape[3]
    buf = torch.zeros(B, cp, dtype=torch.float32, device=g.device)
    buf[:, :act.c] = g
    gt = torch.empty(1, 1, B, cp, dtype=act.t.dtype, device=g.device)
    _axpy(None, gt.view(B, 1, 1, cp), 1.0, buf, False)
    _accumulate(tape, act, gt)


def seed_stacked(tape, act, g_bk):
    """Install a (B, K, c) fp32 gradient for a stacked (K,1,B,cpad) Act."""
    if g_bk is None or not _rg(act):
        return
    K, _, B, cp = act.t.shape
    buf = torch.zeros(K, B, cp, dtype=torch.float32, device=g_bk.device)
    buf[:, :, :act.c] = g_bk.permute(1, 0, 2)
    gt = torch.empty(K, 1, B, cp, dtype=act.t.dtype, device=g_bk.device)
    _axpy(None, gt.view(K * B, 1, 1, cp), 1.0, buf.view(K * B, cp), False)
    _accumulate(tape, act, gt)


def expert_heads(tape, ex, img_feat, speed_a, cmd_a, alt, alpha_out, ap_out, speed_out, tag="expert"):
    """BaseExpert / BaseExpertAlt after the backbone (moe.py:88-101, 113-128). The raw head outputs are written into
    slices of shared (1,1,B,K*16) buffers so the gating kernel sees all experts at once."""
    s = mlp(tape, ex.speed_encoder, [speed_a], tag + ".speed_encoder")
    c = mlp(tape, ex.command_encoder, [cmd_a], tag + ".command_encoder")
    feats = [img_feat, s, c]
    sp = mlp_last_out(tape, ex.speed_pred, feats, speed_out, tag + ".speed_pred")
    af = mlp(tape, ex.action_features, feats, tag + ".action_features")
    ap = linear_op(tape, [af], ex.action_pred, None, out=ap_out, tag=tag + ".action_pred")
    if alt:
        a1 = linear_op(tape, feats, ex.alpha[0], "relu", tag=tag + ".alpha.0")
        al = linear_op(tape, [a1], ex.alpha[2], None, out=alpha_out, tag=tag + ".alpha.2")
    else:
        al = linear_op(tape, [af], ex.alpha, None, out=alpha_out, tag=tag + ".alpha")  # ReLU applied by the gating kernel
    return al, ap, sp


def mlp_last_out(tape, seq, srcs, out, tag):
    """mlp() whose final Linear writes into `out`."""
    mods = list(seq)
    last_lin = max(i for i, m in enumerate(mods) if isinstance(m, torch.nn.Linear))
    if last_lin == 0:
        return linear_op(tape, srcs, mods[0], None, out=out, tag=tag + ".0")
    head = torch.nn.Sequential(*mods[:last_lin])
    y = mlp(tape, head, srcs, tag)
    if last_lin + 1 < len(mods):
        raise RuntimeError("pmoe_b200: l_act on an output head is not supported")
    return linear_op(tape, [y], mods[last_lin], None, out=out, tag="%s.%d" % (tag, last_lin))


# ------------------------------------------------------------------------------------------------ grouped expert heads
# The K experts of MixtureOfExperts carry identically shaped head MLPs (moe.py:53-72). Instead of K launches per Linear,
# the experts become the "image" axis of ONE launch: activations are stacked (K,1,B,C), the library picks expert e's
# weights and bias for the tiles of image e (PmoeConvTc.wpack_img_stride / shift_img_stride). Forward and data gradient
# are grouped; the weight gradient runs per expert on the views of the stacked tensors.
def _grouped_pack(lins, srcs, cop, dt, kind):
    """Stacked (K, rows, cols) packed weights of the K experts' Linears + their packs.Pack list (shared index map)."""
    w0 = lins[0].weight

    def build(w4):
        if kind == "f":
            cols = []
            for x in srcs:
                cols += _pack_cols(w4, x, 0, 0)
            wp = torch.cat(cols, dim=1)
            if cop > wp.shape[0]:
                wp = torch.nn.functional.pad(wp, (0, 0, 0, cop - wp.shape[0]))
            return wp
        # data gradient of source `srcs[0]`: rows = its physical channels, K = padded cout
        blk = torch.cat(_pack_cols(w4, srcs[0], 0, 0), dim=1).t()
        blk = torch.nn.functional.pad(blk, (0, cop - blk.shape[1]))
        rows = ops.cout_padded(blk.shape[0])
        return torch.nn.functional.pad(blk, (0, 0, 0, rows - blk.shape[0]))
    key = (kind, tuple((x.cin0, tuple(x.lay)) for x in srcs), cop, str(dt))
    return packs.packed_group([l.weight for l in lins], key, (w0.shape[0], w0.shape[1], 1, 1), build, dt)


def grouped_linear_op(tape, srcs, lins, act=None, tag=""):
    """K Linear layers of equal shape over stacked inputs: srcs = Acts of shape (K,1,B,C) (virtual concat along C)."""
    dt = tape.dtype
    K = len(lins)
    srcs = concat_sources(srcs)
    dev = srcs[0].t.device
    cout = lins[0].weight.shape[0]
    cop, cstore = ops.cout_padded(cout), pad_ch(cout)
    phys = [x.cpad for x in srcs]
    ck = ops.choose_ck(phys)
    segs = [(i, 0, 0, 0, phys[i] // ck) for i in range(len(srcs))]
    wp, wpks = _grouped_pack(lins, srcs, cop, dt, "f")
    B = srcs[0].t.shape[2]
    has_bias = lins[0].bias is not None
    shift = None
    if has_bias:   # (K, cop) stacked biases, refreshed with the weights
        shift, bias_packs = packs.packed_group([l.bias for l in lins], ("bias", cop), lins[0].bias.shape,
                                               lambda v: torch.nn.functional.pad(v, (0, cop - cout)), torch.float32)
    z_t = torch.empty(K, 1, B, cstore, dtype=dt, device=dev)
    flops = 2.0 * K * B * cout * sum(x.nlog for x in srcs)
    ops.conv([x.t for x in srcs], wp, segs, ck, z_t, shift=shift, act=act, flops=flops, tag="grouped " + tag)
    params = [l.weight for l in lins] + [l.bias for l in lins if l.bias is not None]
    rg = any(_rg(x.act) for x in srcs) or _any_rg(params)
    z = _new_act(tape, z_t, cout, rg)
    if not (tape.save and rg):
        return z

    def backward():
        dz = tape.grad_of(z)
        if dz is None:
            return
        z_saved = z_t if act not in (None, "none") else None
        dy = torch.empty(K, 1, B, cstore, dtype=dt, device=dev)
        want_bias = has_bias and any(l.bias.requires_grad for l in lins)
        bsum = None
        if (FUSE_ACT_BWD_BIAS and want_bias and act in (None, "none", "relu", "elu") and dt == torch.bfloat16 and dz.is_contiguous()
                and dz.dtype == dt and 256 % max(cstore // 8, 1) == 0):
            # activation backward and the per-expert bias gradients (per-image channel sums of dy) in one pass
            bsum = tape.zeros((K, cstore), torch.float32, dev)
            vz, vy = view4(dz), view4(dy)
            vs = view4(z_saved) if z_saved is not None else _lib.null_view()
            rc = profiler.launch("act_bwd_bias", lambda: lib().pmoe_act_bwd_bias(
                C.byref(vz), C.byref(vs), ACT[act], C.byref(vy), bsum.data_ptr(), bsum.stride(0), stream_ptr()), io=(dz, z_saved, dy))
            if rc == -2:
                profiler.uncount()
                bsum = None
            else:
                check(rc, "act_bwd_bias")
        if bsum is None:
            _bn_bwd_apply(dz, z_saved, None, act, None, None, None, None, None, 0.0, 0, dy, None, False)
            if want_bias:
                bsum = nhwc.channel_sums(dy)                  # (K, cstore): per-expert bias gradients in one launch
        if want_bias:
            _group_to_params(tape, bsum if cstore == cop else torch.nn.functional.pad(bsum, (0, cop - cstore)), bias_packs,
                             [l.bias for l in lins])
        if any(l.weight.requires_grad for l in lins):
            # ONE launch for the K experts' weight gradients: the expert is the image axis, each image accumulates into its own slice
            dwp = tape.zeros((K, cop, wp.shape[2]), torch.float32, dev)
            ops.conv_wgrad([x.t for x in srcs], segs, ck, dy, dwp, flops=flops, tag="wgrad grouped " + tag)
            _group_to_params(tape, dwp, wpks, [l.weight for l in lins])
        ck_d = ops.choose_ck([cstore])
        for x in srcs:
            if not _rg(x.act):
                continue
            wd = _grouped_pack(lins, [x], cstore, dt, "d")[0]
            if MERGE_SKINNY_DGRAD and cstore <= 16 and len(srcs) == 1 and x.t is x.act.t:
                # 512 -> 4 and 512 -> 1 heads of one feature tensor (action_pred, alpha): their data gradients are two K = 16 GEMMs that
                # each write / accumulate the whole (K, B, 512) gradient; queued here, they run as ONE K = 32 GEMM (Tape.grad_of)
                tape.pending_dgrad.setdefault(id(x.act), []).append((dy, wd, x.act, 2.0 * K * B * cout * x.nlog, tag))
                continue
            g, existed = _grad_buffer(tape, x.act)
            ops.conv([dy], wd, [(0, 0, 0, 0, cstore // ck_d)], ck_d, g, residual=g if existed else None,
                     flops=2.0 * K * B * cout * x.nlog, tag="dgrad grouped " + tag)

    for lin in lins:
        tape.expect(lin.bias, lin.weight)
    tape.record(backward)
    return z


GROUPED_HEADS = True  # tests switch this off to compare against the per-expert launches
MERGE_SKINNY_DGRAD = _os0.environ.get("PMOE_MERGE_SKINNY_DGRAD", "1") != "0"


def _flush_dgrad(tape, key):
    """Run the queued skinny data gradients of one source activation (grouped_linear_op) as one GEMM over their concatenated K."""
    ents = tape.pending_dgrad.pop(key, None)
    if not ents:
        return
    act = ents[0][2]
    g, existed = _grad_buffer(tape, act)
    if len(ents) == 1:
        dy, wd = ents[0][0], ents[0][1]
    else:
        dy = torch.cat([e[0] for e in ents], dim=3)
        wd = torch.cat([e[1] for e in ents], dim=2)
    kd = dy.shape[3]
    ck = ops.choose_ck([kd])
    ops.conv([dy], wd, [(0, 0, 0, 0, kd // ck)], ck, g, residual=g if existed else None, flops=sum(e[3] for e in ents),
             tag="dgrad grouped " + "+".join(e[4] for e in ents))
FUSE_ACT_BWD_BIAS = _os0.environ.get("PMOE_FUSE_ACT_BWD_BIAS", "1") != "0"   # head Linears: activation backward + bias gradient in one pass


def grouped_supported(experts, alt):
    """Grouped heads need the tensor-core path and the plain MLP structure (Linear [+act] [+Dropout], no BatchNorm1d)."""
    if not GROUPED_HEADS or config.precision() != "bf16" or config.FORCE_SIMT or len(experts) < 2:
        return False
    for ex in experts:
        for seq in (ex.speed_encoder, ex.command_encoder, ex.speed_pred, ex.action_features):
            if any(isinstance(m, torch.nn.BatchNorm1d) for m in seq):
                return False
    return True


def grouped_mlp(tape, seqs, srcs, tag="mlp"):
    """make_mlp Sequentials of K experts (same structure) over stacked inputs."""
    mods0 = list(seqs[0])
    x_srcs, y, i = srcs, None, 0
    while i < len(mods0):
        if not isinstance(mods0[i], torch.nn.Linear):
            raise RuntimeError("pmoe_b200 grouped mlp: unexpected layer order at %d: %r" % (i, mods0[i]))
        j = i + 1
        act = drop = None
        if j < len(mods0) and isinstance(mods0[j], (torch.nn.ReLU, torch.nn.ELU, torch.nn.Tanh, torch.nn.Sigmoid)):
            act = {torch.nn.ReLU: "relu", torch.nn.ELU: "elu", torch.nn.Tanh: "tanh", torch.nn.Sigmoid: "sigmoid"}[type(mods0[j])]
            j += 1
        if j < len(mods0) and isinstance(mods0[j], torch.nn.Dropout):
            drop = mods0[j]
            j += 1
        y = grouped_linear_op(tape, x_srcs, [list(sq)[i] for sq in seqs], act, tag="%s.%d" % (tag, i))
        if drop is not None:
            y = dropout_op(tape, y, drop.p, drop.training)
        x_srcs = [y]
        i = j
    return y


def stack_acts(tape, acts):
    """K Acts of shape (1,1,B,C) -> one Act (K,1,B,C); the gradient of the stack flows back to each part."""
    K = len(acts)
    t = torch.empty(K, 1, acts[0].t.shape[2], acts[0].t.shape[3], dtype=acts[0].t.dtype, device=acts[0].t.device)
    for e, a in enumerate(acts):
        _axpy(a.t, t[e:e + 1], 1.0, None, False)
    st = _new_act(tape, t, acts[0].c, any(_rg(a) for a in acts))
    if tape.save and _rg(st):
        def backward():
            g = tape.grad_of(st)
            if g is None:
                return
            for e, a in enumerate(acts):
                if _rg(a):
                    _accumulate_copy(tape, a, g[e:e + 1])
        tape.record(backward)
    return st


def broadcast_act(tape, a, K):
    """The same (1,1,B,C) input for every expert as a (K,1,B,C) Act (inputs: no gradient)."""
    t = a.t.expand(K, -1, -1, -1).contiguous()
    return _new_act(tape, t, a.c, False)


def expert_heads_grouped(tape, experts, img_feats, speed_a, cmd_a, alt):
    """All experts' heads (moe.py:88-101 / 113-128) with one launch per layer. Returns stacked (K,1,B,16) Acts of the raw
    alpha, action_pred and speed_pred outputs."""
    K = len(experts)
    feat = stack_acts(tape, img_feats)
    s = grouped_mlp(tape, [ex.speed_encoder for ex in experts], [broadcast_act(tape, speed_a, K)], "speed_encoder")
    c = grouped_mlp(tape, [ex.command_encoder for ex in experts], [broadcast_act(tape, cmd_a, K)], "command_encoder")
    feats = [feat, s, c]
    sp = grouped_mlp(tape, [ex.speed_pred for ex in experts], feats, "speed_pred")
    af = grouped_mlp(tape, [ex.action_features for ex in experts], feats, "action_features")
    ap = grouped_linear_op(tape, [af], [ex.action_pred for ex in experts], None, tag="action_pred")
    if alt:
        a1 = grouped_linear_op(tape, feats, [ex.alpha[0] for ex in experts], "relu", tag="alpha.0")
        al = grouped_linear_op(tape, [a1], [ex.alpha[2] for ex in experts], None, tag="alpha.2")
    else:
        al = grouped_linear_op(tape, [af], [ex.alpha for ex in experts], None, tag="alpha")  # ReLU applied by the gating kernel
    return al, ap, sp



class GateMixture:
    """softmax over expert logits + sigma = ELU+1 from the strided raw head outputs; backward to those outputs."""

    def __init__(self, tape, alpha_acts, ap_acts, alpha_buf, ap_buf, B, K, relu_alpha, a_sk, p_sk):
        self.tape, self.alpha_acts, self.ap_acts = tape, alpha_acts, ap_acts
        self.alpha_buf, self.ap_buf, self.B, self.K, self.relu = alpha_buf, ap_buf, B, K, relu_alpha
        self.a_sk, self.p_sk = a_sk, p_sk
        dev = alpha_buf.device
        self.probs = torch.empty(B, K, dtype=torch.float32, device=dev)
        self.mean = torch.empty(B, K, 2, dtype=torch.float32, device=dev)
        self.std = torch.empty(B, K, 2, dtype=torch.float32, device=dev)
        self.route = torch.empty(B, dtype=torch.int64, device=dev)
        a_sb, p_sb = alpha_buf.shape[3], ap_buf.shape[3]
        check(profiler.launch("gate_mixture_fwd", lambda: lib().pmoe_gate_mixture_fwd(
            alpha_buf.data_ptr(), a_sb, a_sk, ap_buf.data_ptr(), p_sb, p_sk, dtype_code(alpha_buf), B, K, int(relu_alpha),
            self.probs.data_ptr(), self.mean.data_ptr(), self.std.data_ptr(), self.route.data_ptr(), stream_ptr())),
            "gate_mixture_fwd")

    def backward(self, dprobs, dmean, dstd):
        tape = self.tape
        if not any(_rg(a) for a in self.alpha_acts + self.ap_acts):
            return
        dalpha = torch.zeros_like(self.alpha_buf)
        dap = torch.zeros_like(self.ap_buf)
        f = lambda t: None if t is None else t.contiguous().float()
        dprobs, dmean, dstd = f(dprobs), f(dmean), f(dstd)
        a_sb, p_sb = self.alpha_buf.shape[3], self.ap_buf.shape[3]
        check(profiler.launch("gate_mixture_bwd", lambda: lib().pmoe_gate_mixture_bwd(
            _lib.ptr(dprobs), _lib.ptr(dmean), _lib.ptr(dstd), self.probs.data_ptr(), self.std.data_ptr(),
            self.alpha_buf.data_ptr(), a_sb, self.a_sk, dtype_code(self.alpha_buf), self.B, self.K, int(self.relu),
            dalpha.data_ptr(), dap.data_ptr(), p_sb, self.p_sk, stream_ptr())), "gate_mixture_bwd")
        for k, a in enumerate(self.alpha_acts):
            if _rg(a):
                _accumulate_copy(tape, a, dalpha[..., k * self.a_sk:k * self.a_sk + a.t.shape[3]] if len(self.alpha_acts) > 1 else dalpha)
        for k, a in enumerate(self.ap_acts):
            if _rg(a):
                _accumulate_copy(tape, a, dap[..., k * self.p_sk:k * self.p_sk + a.t.shape[3]] if len(self.ap_acts) > 1 else dap)


# ------------------------------------------------------------------------------------------------ autograd bridge
import os as _os
DIRECT_GRADS = _os.environ.get("PMOE_DIRECT_GRADS", "0") == "1"
class TapeFunction(torch.autograd.Function):
    """forward(runner, *params): runner(tape) -> (list of output tensors, seed_fn). seed_fn(grad_outputs)
    installs the output gradients into the tape; backward then replays the tape and returns the parameter
    gradients accumulated in it."""

    @staticmethod
    def forward(ctx, runner, *params):
        tape = Tape(config.act_dtype(), save=any(p.requires_grad for p in params))
        outs, seed = runner(tape)
        tape.finish_forward()
        tape.registry = packs.current()
        ctx.tape, ctx.seed, ctx.plist = tape, seed, params
        ctx.dp = dp.current()  # a pmoe_b200.dp.DataParallel wrapper when the model runs under one
        ctx.mark_non_differentiable(*[o for o in outs if not o.is_floating_point()])
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts):
        tape = ctx.tape
        if tape is None:
            raise RuntimeError("pmoe_b200: backward through the same forward pass twice (the tape frees its buffers as it replays; "
                               "run forward again)")
        if ctx.dp is not None:
            tape.bucketer = ctx.dp.make_bucketer(ctx.plist, eager_alloc=bool(tape.side_streams))
            if tape.side_streams:
                tape.bucketer.producer_streams = lambda: list(tape.side_streams)
        with packs.resumed(getattr(tape, "registry", None)):
            ctx.seed(tape, gouts)
            # the seed closure holds the forward's output tensors, whose grad_fn is this node: drop it so that the finished graph (and
            # the AccumulateGrad nodes it keeps alive, with the stream they were created on) can be freed
            ctx.seed = None
            tape.backward()
        if tape.bucketer is not None:
            reduced, stats = tape.bucketer.finish()
            ctx.dp.note_reduced([p for p in ctx.plist if p.requires_grad], stats)
            if ctx.dp.as_bucket_view:
                # .grad aliases the reduced bucket slices: nothing is handed to AccumulateGrad
                for p in ctx.plist:
                    v = reduced.get(id(p)) if p.requires_grad else None
                    if v is not None and (p.grad is None or p.grad.data_ptr() != v.data_ptr()):
                        p.grad = v
                packs.bump([p.grad for p in ctx.plist if p.requires_grad and p.grad is not None])
                grads = tuple(None for _ in ctx.plist)
            else:
                grads = tuple(reduced.get(id(p)) if p.requires_grad else None for p in ctx.plist)
        else:
            grads = tuple((tape.pgrads.get(id(p)) if (p.requires_grad and id(p) not in tape.direct) else None) for p in ctx.plist)
            if tape.direct:
                packs.bump([p.grad for p in ctx.plist if id(p) in tape.direct])
        ctx.tape = None
        if DIRECT_GRADS and ctx.dp is None:
            # hand the gradients to .grad ourselves: the engine would clone each of the ~500 tensors it cannot steal
            with torch.no_grad():
                for p, g in zip(ctx.plist, grads):
                    if g is None:
                        continue
                    if p.grad is None:
                        p.grad = g
                    else:
                        p.grad.add_(g)
            return (None,) * (len(ctx.plist) + 1)
        return (None,) + grads


def run(module, runner):
    """Execute `runner(tape)` for `module`: through autograd when gradients are wanted, directly otherwise."""
    params = [p for p in module.parameters()]
    packs.begin_pass(module)
    try:
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            outs = TapeFunction.apply(runner, *params)
        else:
            tape = Tape(config.act_dtype(), save=False)
            outs, _ = runner(tape)
            tape.finish_forward()
    finally:
        packs.end_pass()
    return outs


def _seed_nchw(act):
    """seed function for an NHWC output Act that was exported as an fp32 NCHW tensor."""
    def seed(tape, g):
        if g is None or not _rg(act):
            return
        buf = torch.empty(act.t.shape, dtype=act.t.dtype, device=act.t.device)
        nhwc.from_nchw(g.contiguous(), out=buf)
        _accumulate(tape, act, buf)
    return seed


def unet_module_forward(net, image, body=None):
    """UNet.forward for the module API: image fp32 NCHW -> logits fp32 NCHW (and pooled bottleneck)."""
    body = unet if body is None else body

    def runner(tape):
        x = nhwc.from_nchw(image, dtype=tape.dtype)
        x.rg = False
        logits, inter = body(tape, net, x, want_inter=net.inter_repr)
        out = nhwc.to_nchw(logits.t, logits.c)
        seed_logits = _seed_nchw(logits)
        if net.inter_repr:
            def seed(tp, gouts):
                if gouts[0] is not None:
                    inter.backward(gouts[0])
                seed_logits(tp, gouts[1])
            return [inter.value, out], seed
        return [out], (lambda tp, gouts: seed_logits(tp, gouts[0]))
    outs = run(net, runner)
    return (outs[0], outs[1]) if net.inter_repr else outs[0]


def punet_tape(tape, net, images):
    """PredictiveUnet.forward (punet.py:75-120) on the tape. Every U-Net call is separate (train-mode BatchNorm
    statistics are per call, as in the reference's Python loop). The deque of masks is a sliding window over one ring
    buffer (B,H,W,(P+F)*slot). Returns dict(ring, pools, masks, futures, inter, slot, ncls)."""
    B, T, Cin, H, W = images.shape
    P, Fu = net.n_past_frames, net.n_future_frames
    ncls = net.unet.out.weight.shape[0]
    slot = pad_ch(ncls)
    dev = images.device
    nslots = P + max(Fu, 0)
    ring = torch.zeros(B, H, W, nslots * slot, dtype=tape.dtype, device=dev)
    pools = torch.zeros(B, nslots * slot, dtype=torch.float32, device=dev)
    masks, futures, inter = [], [], None
    for t in range(P):
        x = nhwc.from_nchw(images[:, t], dtype=tape.dtype)
        m, it = unet(tape, net.unet, x, out=ring[..., t * slot:(t + 1) * slot], out_pool=pools[:, t * slot:],
                     pool_stride=nslots * slot, want_inter=(net.unet_inter_repr and Fu == 0 and t == P - 1), tag="unet")
        masks.append(m)
        inter = it
    for f in range(Fu):
        window = ring_window(tape, ring, masks[f:f + P], f * slot, slot, ncls)
        e = eca_conv_block(tape, net.entry_block, window, (P, ncls, slot), pools[:, f * slot:(f + P) * slot], tag="entry")
        m, inter = unet(tape, net.pred_unet, e, out=ring[..., (P + f) * slot:(P + f + 1) * slot],
                        out_pool=pools[:, (P + f) * slot:], pool_stride=nslots * slot, want_inter=net.inter_repr,
                        tag="pred_unet")
        masks.append(m)
        futures.append(m)
    return {"ring": ring, "pools": pools, "masks": masks, "futures": futures, "inter": inter, "slot": slot, "ncls": ncls,
            "P": P, "F": Fu}


def ring_window(tape, ring, parts, c_begin, slot, ncls):
    """A run of consecutive ring slots as ONE multi-group Act; its gradient is split back onto the slot Acts."""
    n = len(parts)
    window = _new_act(tape, ring[..., c_begin:c_begin + n * slot], n * ncls, any(_rg(m) for m in parts))
    if tape.save and _rg(window):
        def split_grad():
            g = tape.grad_of(window)
            if g is None:
                return
            for i, part in enumerate(parts):
                if _rg(part):
                    _accumulate_copy(tape, part, g[..., i * slot:(i + 1) * slot])
        tape.record(split_grad)
    return window


def punet_module_forward(net, images):
    B, T, Cin, H, W = images.shape

    def runner(tape):
        r = punet_tape(tape, net, images)
        ncls, Fu, masks, futures, inter = r["ncls"], r["F"], r["masks"], r["futures"], r["inter"]
        if Fu == 0:
            if net.unet_inter_repr:
                return [inter.value], (lambda tp, g: inter.backward(g[0]) if g[0] is not None else None)
            return [nhwc.to_nchw(masks[-1].t, ncls)], (lambda tp, g: _seed_nchw(masks[-1])(tp, g[0]))
        if net.inter_repr:
            return [inter.value], (lambda tp, g: inter.backward(g[0]) if g[0] is not None else None)
        out = torch.empty(B, Fu, ncls, H, W, dtype=torch.float32, device=images.device)
        for f, m in enumerate(futures):
            nhwc.to_nchw(m.t, ncls, out=out[:, f])

        def seed(tp, gouts):
            g = gouts[0]
            if g is None:
                return
            for f, m in enumerate(futures):
                _seed_nchw(m)(tp, g[:, f])
        return [out], seed

    return run(net, runner)[0]


def _accumulate_copy(tape, act, gview):
    """grads[act] += gview where gview is a strided view with act's geometry."""
    k = id(act)
    tape.presums.pop(k, None)
    if k in tape.grads:
        _axpy(gview, tape.grads[k], 1.0, None, True)
    else:
        buf = torch.empty(act.t.shape[0], act.t.shape[1], act.t.shape[2], act.t.shape[3], dtype=gview.dtype, device=gview.device)
        _axpy(gview, buf, 1.0, None, False)
        tape.grads[k] = buf


def nhwc_module_forward(module, x_nchw, body):
    """Run `body(tape, Act) -> Act` for a stand-alone block called through the NCHW module API."""
    def runner(tape):
        x = nhwc.from_nchw(x_nchw, dtype=tape.dtype)
        y = body(tape, x)
        return [nhwc.to_nchw(y.t, y.c)], (lambda tp, g: _seed_nchw(y)(tp, g[0]))
    return run(module, runner)[0]
