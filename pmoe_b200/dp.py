"""Batch data-parallel training over the GPUs of one box (one process per GPU, torch.distributed / NCCL).

The reference has no distributed path at all: the split is inserted where its training loop calls
`loss.backward()` (trainer/train_2.py:157-158). Each rank runs the model on its shard of the batch — per-rank
BatchNorm statistics, as in the reference run at that batch size — and the gradients are averaged:

* the tape (pmoe_b200.train) reports every parameter gradient the moment its last contribution has been written;
* `GradBucketer` packs them, in that order, into flat fp32 buckets (~32 MB) and starts an asynchronous all-reduce of a
  bucket as soon as it is complete, so the reduction of the late layers' gradients travels over NVLink while the tape is
  still running dgrad/wgrad kernels of the early layers;
* the autograd function returns views of the reduced buckets, so `.grad`, `clip_grad_norm_` and the optimizer of the
  reference's loop see averaged gradients with no further change.

There is no data-path collective in forward. Host-side logic only: all arithmetic stays in the CUDA library / NCCL.
"""
import contextlib
import threading

import torch
import torch.distributed as dist

_tls = threading.local()


def current():
    """The DataParallel wrapper whose forward is running on this thread (None outside)."""
    return getattr(_tls, "dp", None)


def shard(t, rank=None, world=None, dim=0):
    """This rank's contiguous slice of a global batch (rank r gets [r*B/N, (r+1)*B/N))."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    n = t.shape[dim]
    if n % world:
        raise ValueError("global batch %d is not divisible by the %d ranks" % (n, world))
    per = n // world
    return t.narrow(dim, rank * per, per)


class GradBucketer:
    """Flat fp32 buckets over a fixed parameter list; all-reduce(avg) per bucket, launched in bucket order as buckets
    fill up (every rank fills them in the same order because every rank replays the same tape)."""

    def __init__(self, params, group=None, bucket_bytes=32 << 20, device=None, persistent=False):
        self.group = group
        self.persistent = persistent   # keep the flat buffers across backward passes (gradient_as_bucket_view)
        self._keep = None
        self.eager_alloc = False       # set while a multi-stream tape is the producer (pmoe_b200.train.Tape.branch)
        self.producer_streams = None   # callable -> CUDA streams that may hold un-finished writes into the buckets
        self._comm_stream = None
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.params = [p for p in params if p.requires_grad]
        self.device = device if device is not None else (self.params[0].device if self.params else torch.device("cpu"))
        # buckets in REVERSE registration order: backward produces the last layers' gradients first
        self.slots = {}      # id(param) -> (bucket index, offset, numel)
        self.sizes = []
        cur, off = 0, 0
        cap = max(1, bucket_bytes // 4)
        for p in reversed(self.params):
            n = p.numel()
            if off > 0 and off + n > cap:
                self.sizes.append(off)
                cur, off = cur + 1, 0
            self.slots[id(p)] = (cur, off, n)
            off += (n + 3) // 4 * 4  # keep every slot 16-byte aligned for the vectorised optimizer kernels
        if off > 0:
            self.sizes.append(off)
        self.members = [0] * len(self.sizes)
        for (b, _, _) in self.slots.values():
            self.members[b] += 1
        self.reset()

    # ---- one backward pass
    def reset(self):
        if self.persistent:
            if self._keep is None:
                self._keep = [torch.zeros(n, dtype=torch.float32, device=self.device) for n in self.sizes]
            else:
                for buf in self._keep:   # parameters that receive no gradient in this pass must read zero
                    buf.zero_()
            self.flat = list(self._keep)
        else:
            # allocated (and zero-filled) up front on the caller's stream: gradient kernels may write the slots from side streams
            self.flat = [torch.zeros(n, dtype=torch.float32, device=self.device) for n in self.sizes] if self.eager_alloc else \
                [None] * len(self.sizes)
        self.filled = [0] * len(self.sizes)
        self.have = set()
        self.handles = []
        self.next_launch = 0
        self.launched = 0

    def _buffer(self, b):
        if self.flat[b] is None:
            self.flat[b] = torch.zeros(self.sizes[b], dtype=torch.float32, device=self.device)
        return self.flat[b]

    def slot(self, p):
        """p's slice of its flat bucket: the tape's kernels write the gradient straight into it (no staging copy)."""
        b, off, n = self.slots[id(p)]
        return self._buffer(b)[off:off + n]

    def ready(self, p, grad=None):
        """Gradient of `p` is final (already in its slot when `grad` is None, copied there otherwise); launch every bucket
        that became complete (in order)."""
        b, off, n = self.slots[id(p)]
        if id(p) in self.have:
            raise RuntimeError("pmoe_b200.dp: gradient reported twice for one parameter")
        self.have.add(id(p))
        if grad is not None:
            self._buffer(b)[off:off + n].copy_(grad.reshape(-1))
        self.filled[b] += 1
        self._launch_complete()

    def _launch_complete(self):
        while self.next_launch < len(self.sizes) and self.filled[self.next_launch] == self.members[self.next_launch]:
            self._launch(self.next_launch)
            self.next_launch += 1

    def _launch(self, b):
        buf = self._buffer(b)
        self.launched += 1
        if self.world == 1:
            return
        backend = dist.get_backend(self.group)
        if backend == "nccl":
            streams = self.producer_streams() if self.producer_streams is not None else []
            if streams:
                # the bucket was filled from several streams: launch the reduction from a stream that waits for all of them,
                # without stalling any of the producers
                if self._comm_stream is None:
                    self._comm_stream = torch.cuda.Stream(device=self.device)
                comm = self._comm_stream
                comm.wait_stream(torch.cuda.current_stream())
                for st in streams:
                    comm.wait_stream(st)
                with torch.cuda.stream(comm):
                    self.handles.append(dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=self.group, async_op=True))
                self._used_comm = True
                return
            self.handles.append(dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=self.group, async_op=True))
        else:  # gloo (CPU tests): sum, then scale
            h = dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self.handles.append((h, buf))

    def finish(self):
        """Launch what is left (parameters that received no gradient contribute zeros, identically on every rank),
        wait for all reductions and return {id(param): reduced gradient view or None}."""
        while self.next_launch < len(self.sizes):
            self._launch(self.next_launch)
            self.next_launch += 1
        for h in self.handles:
            if isinstance(h, tuple):
                h[0].wait()
                h[1].div_(self.world)
            else:
                h.wait()  # makes the current stream wait for the NCCL stream
        if getattr(self, "_used_comm", False):
            torch.cuda.current_stream().wait_stream(self._comm_stream)   # rejoin (also ends the branch inside a graph capture)
            self._used_comm = False
        out = {}
        for p in self.params:
            b, off, n = self.slots[id(p)]
            out[id(p)] = self.flat[b][off:off + n].view(p.shape) if (id(p) in self.have or self.persistent) else None
        stats = {"buckets": len(self.sizes), "launched": self.launched, "bytes": 4 * sum(self.sizes)}
        if not self.persistent:
            self.reset()
        return out, stats


class DataParallel(torch.nn.Module):
    """Wraps one of the pmoe_b200 models. forward/sample delegate to the module; while forward runs, the tape
    executor picks this wrapper up (dp.current()) and routes its parameter gradients through a GradBucketer.
    Parameters and buffers are broadcast from rank 0 at construction (replicas start identical)."""

    def __init__(self, module, process_group=None, bucket_mb=32, broadcast=True, gradient_as_bucket_view=False):
        """gradient_as_bucket_view: every parameter's .grad IS its slice of the flat all-reduce buckets — the backward kernels
        write there, NCCL reduces in place, the optimizer reads the result: no copy or add between them. Each backward
        OVERWRITES the gradients (no accumulation across backward calls): use it when every optimizer step follows exactly
        one backward pass."""
        super().__init__()
        self.module = module
        self.as_bucket_view = bool(gradient_as_bucket_view)
        self.group = process_group
        self.bucket_bytes = int(bucket_mb * (1 << 20))
        self.last_stats = None
        self._reduced = set()
        self._bucketers = {}
        if dist.is_initialized() and dist.get_world_size(process_group) > 1:
            if broadcast:
                with torch.no_grad():
                    for t in list(module.parameters()) + list(module.buffers()):
                        dist.broadcast(t, src=dist.get_global_rank(process_group, 0) if process_group is not None else 0, group=process_group)
            # gradients that do not come out of a tape (e.g. PMoE's 2->1 combiners, model/moe.py:345-346)
            for p in module.parameters():
                if p.requires_grad:
                    p.register_post_accumulate_grad_hook(self._late_hook)

    def _late_hook(self, p):
        if id(p) in self._reduced or p.grad is None:
            return
        world = dist.get_world_size(self.group)
        dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=self.group)
        p.grad.div_(world)

    def make_bucketer(self, params, eager_alloc=False):
        key = tuple(id(p) for p in params if p.requires_grad)
        b = self._bucketers.get(key)
        if b is None:
            b = self._bucketers[key] = GradBucketer(params, self.group, self.bucket_bytes, persistent=self.as_bucket_view)
        b.eager_alloc = eager_alloc
        b.producer_streams = None
        b.reset()
        return b

    def note_reduced(self, params, stats):
        self._reduced.update(id(p) for p in params)
        self.last_stats = stats

    @contextlib.contextmanager
    def _active(self):
        prev = current()
        _tls.dp = self
        try:
            yield
        finally:
            _tls.dp = prev

    def forward(self, *args, **kwargs):
        self._reduced = set()
        with self._active():
            return self.module(*args, **kwargs)

    def sample(self, *args, **kwargs):
        return self.module.sample(*args, **kwargs)

    def state_dict(self, *args, **kwargs):  # checkpoints keep the reference's key names (no "module." prefix)
        return self.module.state_dict(*args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        return self.module.load_state_dict(*args, **kwargs)
