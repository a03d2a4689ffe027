"""Host-side logic of the data-parallel path on CPU: world_size-2 gloo process group, GradBucketer averaging,
launch order, parameters without gradient, batch sharding. (The NCCL/GPU leg is scripts/gpu_dp_check.py.)"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pmoe_b200 import dp
        torch.manual_seed(0)
        params = [torch.nn.Parameter(torch.zeros(*s)) for s in [(64, 12, 3, 3), (64,), (64,), (300, 7), (5,), (1000, 33)]]
        params[4].requires_grad_(False)          # frozen: never bucketed
        b = dp.GradBucketer(params, bucket_bytes=64 * 1024)
        assert len(b.sizes) >= 2, b.sizes       # several buckets at this size
        # gradients differ per rank; parameter 3 gets no gradient on any rank
        grads = {i: torch.full(p.shape, float(i + 1) * (rank + 1)) for i, p in enumerate(params) if i not in (3, 4)}
        order = [5, 2, 1, 0]                     # backward order = reverse registration order
        for i in order:
            b.ready(params[i], grads[i])
        launched_before_finish = b.launched
        red, stats = b.finish()
        res = {"launched_early": launched_before_finish, "stats": stats}
        for i, p in enumerate(params):
            if i == 4:
                assert id(p) not in red
            elif i == 3:
                assert red[id(p)] is None
            else:
                want = float(i + 1) * (1 + 2) / 2.0   # mean over the two ranks
                assert torch.allclose(red[id(p)], torch.full(p.shape, want)), (i, red[id(p)].flatten()[:3])
                assert red[id(p)].shape == p.shape
        # out-of-order arrival must still launch buckets in index order (same sequence on every rank)
        b2 = dp.GradBucketer(params, bucket_bytes=64 * 1024)
        for i in ([0, 1, 2, 5] if rank == 0 else [5, 2, 1, 0]):
            b2.ready(params[i], grads[i])
        red2, _ = b2.finish()
        assert torch.allclose(red2[id(params[0])], torch.full(params[0].shape, 1.5))
        # shard(): contiguous equal slices
        t = torch.arange(8 * 3).view(8, 3)
        sh = dp.shard(t)
        assert sh.shape[0] == 4 and sh[0, 0].item() == rank * 12
        with pytest.raises(ValueError):
            dp.shard(torch.zeros(7, 2))
        # DataParallel wrapper: broadcast from rank 0 + late hook for gradients that bypass the tape
        lin = torch.nn.Linear(3, 2)
        with torch.no_grad():
            lin.weight.fill_(float(rank + 1))
        w = dp.DataParallel(lin)
        assert torch.allclose(lin.weight, torch.ones_like(lin.weight))  # rank 0's values everywhere
        y = w(torch.full((1, 3), float(rank + 1))).sum()
        y.backward()
        assert torch.allclose(lin.weight.grad, torch.full_like(lin.weight, 1.5))  # mean of x over the ranks
        assert list(w.state_dict().keys()) == ["weight", "bias"]
        out[rank] = res
    finally:
        dist.destroy_process_group()


def test_bucketer_world2_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert set(out.keys()) == {0, 1}
    assert out[0]["launched_early"] >= 1          # at least one bucket went out before finish(): overlap
    assert out[0]["stats"]["buckets"] == out[1]["stats"]["buckets"]


def test_bucketer_single_process_no_group():
    from pmoe_b200 import dp
    params = [torch.nn.Parameter(torch.zeros(10)), torch.nn.Parameter(torch.zeros(3, 3))]
    b = dp.GradBucketer(params, bucket_bytes=1 << 20)
    b.ready(params[1], torch.ones(3, 3))
    b.ready(params[0], torch.arange(10.0))
    red, stats = b.finish()
    assert torch.equal(red[id(params[0])], torch.arange(10.0)) and torch.equal(red[id(params[1])], torch.ones(3, 3))
    assert stats["buckets"] == 1
    with pytest.raises(RuntimeError):
        b.ready(params[0], torch.zeros(10))
        b.ready(params[0], torch.zeros(10))


def test_bucketer_slots_are_written_in_place_and_persistent_buckets_alias():
    """slot(): the backward kernels write a gradient straight into its bucket slice (no staging copy); persistent buckets
    (gradient_as_bucket_view) hand back the SAME storage every pass, zeroed, so .grad can alias it."""
    from pmoe_b200 import dp
    params = [torch.nn.Parameter(torch.zeros(10)), torch.nn.Parameter(torch.zeros(3, 3)), torch.nn.Parameter(torch.zeros(5))]
    b = dp.GradBucketer(params, bucket_bytes=1 << 20, persistent=True)
    s1 = b.slot(params[1])
    s1.copy_(torch.arange(9.0))
    b.ready(params[1])
    b.slot(params[0]).fill_(2.0)
    b.ready(params[0])
    red, _ = b.finish()
    assert torch.equal(red[id(params[1])], torch.arange(9.0).view(3, 3)) and torch.equal(red[id(params[0])], torch.full((10,), 2.0))
    assert torch.equal(red[id(params[2])], torch.zeros(5))       # no gradient this pass: reads zero
    ptr = red[id(params[1])].data_ptr()
    b.reset()
    assert b.slot(params[1]).data_ptr() == ptr and float(b.slot(params[1]).abs().sum()) == 0.0


def test_checkpoint_with_numpy_scalars_loads(tmp_path):
    """The reference trainers store np.mean(...) entries next to the state dicts (train_2.py:346-370); torch >= 2.6 refuses
    them under weights_only=True."""
    import numpy as np
    from pmoe_b200.utils import io
    path = io.save_checkpoint({"epoch": 3, "best": np.float64(0.25), "e_loss": [np.mean([1.0, 2.0])], "model": {"w": torch.ones(2)}},
                              False, str(tmp_path), "model-3")
    ck = io.load_checkpoint(path, "cpu")
    assert float(ck["best"]) == 0.25 and torch.equal(ck["model"]["w"], torch.ones(2))
