"""Copy-only probe for the end-to-end inference ceiling (no compute): N ranks (torchrun), each copying what one B=256 PU-Net
serving step moves — 0.62 GB host->device and 7.09 GB device->host through the same kind of pinned buffers the bench uses
(frame-major, one async copy per future frame) — all ranks at once. Reports per-rank and aggregate GB/s per direction, both
directions overlapped, and the NUMA placement (GPU bus -> local CPU list, CPUs this process may run on).
    python -m torch.distributed.run --nproc-per-node N scripts/gpu_copy_probe.py   ->  gpurun_out/copy_probe_nN.json"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    import datetime
    dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=120))
dev = torch.device("cuda", local)
B, Fu, C, H, W = 256, 6, 23, 224, 224
h_in = torch.empty(B, 4, 3, H, W, dtype=torch.float32).pin_memory()
d_in = torch.empty_like(h_in, device=dev)
d_out = torch.empty(Fu, B, C, H, W, dtype=torch.float32, device=dev)
h_out = torch.empty(Fu, B, C, H, W, dtype=torch.float32).pin_memory()
s_up, s_down = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn, reps=3):
    fn()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def up():
    with torch.cuda.stream(s_up):
        d_in.copy_(h_in, non_blocking=True)


def down():
    with torch.cuda.stream(s_down):
        for f in range(Fu):
            h_out[f].copy_(d_out[f], non_blocking=True)


def both():
    up()
    down()


nb_up, nb_down = h_in.numel() * 4, h_out.numel() * 4
res = {"world": world, "bytes_up_per_rank": nb_up, "bytes_down_per_rank": nb_down}
for name, fn, nb in (("h2d", up, nb_up), ("d2h", down, nb_down), ("both", both, nb_up + nb_down)):
    dt = timed(fn)
    res[name] = {"seconds_max_over_ranks": dt, "gbs_per_rank": nb / dt / 1e9, "gbs_aggregate": world * nb / dt / 1e9}
res["frames_per_s_ceiling_from_d2h"] = world * B / res["d2h"]["seconds_max_over_ranks"]
try:
    p = torch.cuda.get_device_properties(local)
    bus = "%04x:%02x:%02x.0" % (getattr(p, "pci_domain_id", 0), p.pci_bus_id, p.pci_device_id)
    res["numa"] = {"gpu_bus": bus, "local_cpulist": open("/sys/bus/pci/devices/%s/local_cpulist" % bus).read().strip(),
                   "numa_node": open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip(),
                   "cpus_allowed": len(os.sched_getaffinity(0)), "cpus_total": os.cpu_count(),
                   "numa_nodes_online": open("/sys/devices/system/node/online").read().strip()}
except Exception as ex:
    res["numa"] = {"error": str(ex)}
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/copy_probe_n%d.json" % world, "w"), indent=1)
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
