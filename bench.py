#!/usr/bin/env python
"""Benchmark of the pmoe_b200 hot path (contract: see the task statement / DESIGN.md §Measurement).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K --warmup W   # the reference algorithm on host cores

Default workload = BASELINE.json configs[1]: PU-Net encoder-decoder inference, batch 256 synthetic
frames (4 past frames of 3x224x224 each -> 6 future 23-class masks), bf16 storage / fp32 accumulate,
one replica per GPU (inference does not shard: "replicas only", weak scaling).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PUNET_GF_PER_SAMPLE = 730.82  # forward GFLOP per sample, 2*MACs of conv/convT/linear
PUNET_CFG = dict(past_frames=4, future_frames=6, in_features=3, num_classes=23, gamma=2, b=1, inter_repr=False,
                 unet_inter_repr=False, model_name="unet")


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_sustained": d["bf16_tflops_sustained"], "bf16_burst": d["bf16_tflops"], "hbm": d["hbm_gbs"], "src": "measured"}
    return {"bf16_sustained": 1400.0, "bf16_burst": 1590.0, "hbm": 6650.0, "src": "fallback"}


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons during the timed region (pynvml; nvidia-smi semantics)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._halt = index, [], set(), None, threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap"}
            while not self._halt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.05)
        except Exception as e:  # clocks are evidence, not a dependency
            self.reasons.add("sampler_error:%s" % type(e).__name__)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def build_punet_state(seed=0):
    from oracle import functional as O  # weights only: a seeded, reference-shaped state_dict (random init, no checkpoints offline)
    pc = dict(PUNET_CFG)
    return O.seeded_state_dict(O.make_spec(O.punet_spec, pc), seed)


def synth_images(B, seed=1234):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(B, 4, 3, 224, 224, generator=g)


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args, rank, world):
    """The reference algorithm (oracle port = the same ATen CPU operators the reference modules call) on the
    host cores, all threads, on a bounded sample of the same workload."""
    if rank != 0:
        return
    from oracle import functional as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Bs = args.cpu_batch
    sd = build_punet_state()
    x = synth_images(Bs)
    times = []
    with torch.no_grad():
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            O.punet(x, sd, "", False, 4, 6)
            dt = time.perf_counter() - t0
            if i >= args.warmup:
                times.append(dt)
    total = sum(times)
    val = Bs * len(times) / total
    line = {"impl": "reference", "metric": "infer_frames_per_sec", "value": val, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "punet_infer (BASELINE configs[1]): PredictiveUnet 4->6 frames, 3x224x224, eval", "batch": Bs},
            "cpu_baseline": {"value": val, "unit": "frames/s", "cores": cores, "kind": "port",
                             "sample": "batch %d of the 256-frame workload per step (fp32 ATen/oneDNN, %d threads)" % (Bs, cores)},
            "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline_sample(batch=2, iters=1):
    from oracle import functional as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = build_punet_state()
    x = synth_images(batch)
    with torch.no_grad():
        O.punet(x, sd, "", False, 4, 6)  # warm-up
        t0 = time.perf_counter()
        for _ in range(iters):
            O.punet(x, sd, "", False, 4, 6)
        dt = time.perf_counter() - t0
    return {"value": batch * iters / dt, "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": "batch %d x %d iterations of the same PU-Net forward (fp32, %d threads)" % (batch, iters, cores)}


# ------------------------------------------------------------------------------------------------ CUDA arm
def run_cuda(args, rank, world, local_rank):
    import tempfile
    from pmoe_b200 import _lib, ops, profiler
    from pmoe_b200.model.punet import PredictiveUnet

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    _lib.check(_lib.lib().pmoe_device_check(), "device_check")
    peaks = load_peaks()
    B = args.batch
    sd = build_punet_state()
    with tempfile.TemporaryDirectory() as td:
        ck = os.path.join(td, "unet.pth")
        torch.save({"unet": {k[5:]: v for k, v in sd.items() if k.startswith("unet.")}}, ck)
        net = PredictiveUnet(**dict(PUNET_CFG, model_path=ck))
    net.load_state_dict(sd, strict=True)
    net = net.to(dev).eval()

    host_in = synth_images(B, seed=1234 + rank).pin_memory()
    x = host_in.to(dev, non_blocking=True)
    host_out = torch.empty(B, 6, 23, 224, 224, dtype=torch.float32).pin_memory()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            y = net(x)
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        profiler.reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            y = net(x)
        e1.record()
        barrier()
        clocks = sampler.stop()
        launches = profiler.launch_count()
        ms_total = e0.elapsed_time(e1)

        # end-to-end through the module API with host buffers: H2D of the pinned input and D2H of the
        # full fp32 logits every step.
        for _ in range(1):
            y = net(host_in.to(dev, non_blocking=True))
            host_out.copy_(y, non_blocking=True)
        barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        esteps = max(1, min(args.steps, 3))
        e2.record()
        for _ in range(esteps):
            xin = host_in.to(dev, non_blocking=True)
            y = net(xin)
            host_out.copy_(y, non_blocking=True)
        e3.record()
        barrier()
        ms_e2e = e2.elapsed_time(e3)

        # per-launch CUDA-event pass over one extra step (not part of the timed region): time and
        # algorithmic FLOPs of every tensor-core conv launch -> roofline of the dominant kernel.
        profiler.enable_events(True)
        y = net(x)
        torch.cuda.synchronize()
        prof = profiler.summary()
        profiler.enable_events(False)

    t = torch.tensor([ms_total, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms_total, ms_e2e = t.tolist()
    if rank != 0:
        return
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total / 1e3)
    e2e_val = world * B * esteps / (ms_e2e / 1e3)
    conv = prof.get("conv_tc", {"ms": 0.0, "flops": 0.0, "launches": 0})
    achieved = conv["flops"] / (conv["ms"] * 1e-3) / 1e12 if conv["ms"] > 0 else 0.0
    line = {
        "metric": "infer_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "punet_infer (BASELINE configs[1]): PredictiveUnet 4->6 frames, 3x224x224, eval, random-init weights",
                   "batch_per_gpu": B, "global_batch": B * world, "parallelism": "replicas x%d" % world,
                   "l2": "inputs (%.0f MB) and activations exceed the 126 MB L2 every step" % (host_in.numel() * 4 / 1e6),
                   "whole_step_tflops": PUNET_GF_PER_SAMPLE * B / ms_step / 1e3},
        "e2e": {"value": e2e_val, "unit": "frames/s", "h2d_bytes_per_step": host_in.numel() * 4,
                "d2h_bytes_per_step": host_out.numel() * 4},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "conv_tc_kernel (tcgen05 implicit GEMM)", "achieved": achieved,
                     "peak": peaks["bf16_sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16_sustained"],
                     "peak_source": "%s bf16_tflops_sustained" % peaks["src"], "traffic": None,
                     "launches_per_step": conv["launches"], "kernel_ms_per_step": conv["ms"],
                     "share_of_step": conv["ms"] / ms_step if ms_step > 0 else None,
                     "other_kernels_ms": {k: v["ms"] for k, v in prof.items() if k != "conv_tc"}},
    }
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_sample(args.cpu_batch, 1)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="frames per GPU per step")
    ap.add_argument("--cpu-batch", type=int, default=2, help="bounded CPU sample size")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_cuda(args, rank, world, local_rank)
    finally:
        if world > 1:
            torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
