"""Pinned-host <-> device copy bandwidth on this box, with the process bound to the GPU's local CPUs or not (NUMA
placement of the pinned buffers decides whether a D2H stream reaches PCIe line rate)."""
import os
import sys
import torch


def local_cpus(dev=0):
    try:
        p = torch.cuda.get_device_properties(dev)
        bus = "%04x:%02x:%02x.0" % (getattr(p, "pci_domain_id", 0), p.pci_bus_id, p.pci_device_id)
        txt = open("/sys/bus/pci/devices/%s/local_cpulist" % bus).read().strip()
        cpus = set()
        for part in txt.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        return bus, txt, cpus
    except Exception as ex:
        return None, str(ex), set()


def bw(nbytes=2 << 30):
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    out = {}
    for name, fn in (("d2h", lambda: h.copy_(d, non_blocking=True)), ("h2d", lambda: d.copy_(h, non_blocking=True))):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            fn()
        e1.record(); torch.cuda.synchronize()
        out[name] = 3 * nbytes / e0.elapsed_time(e1) / 1e6
    return out


torch.cuda.init()
print("cpus available:", len(os.sched_getaffinity(0)), "of", os.cpu_count())
print("unbound:", bw())
bus, txt, cpus = local_cpus(0)
print("gpu bus", bus, "local cpulist", txt)
cpus &= os.sched_getaffinity(0)
if cpus:
    os.sched_setaffinity(0, cpus)
    print("bound to %d local cpus:" % len(cpus), bw())
try:
    print(open("/sys/devices/system/node/online").read().strip(), "numa nodes online")
except Exception:
    pass
