#!/usr/bin/env python
"""Benchmark of the pmoe_b200 hot path (contract: see the task statement / DESIGN.md §Measurement).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K --warmup W   # the reference algorithm on host cores

Primary line = BASELINE.json configs[2]: the full mixture (6 experts, each a ResNet18-ECA encoder + gating + action / speed
heads) TRAINING step at a global batch of 512, bf16 storage / fp32 accumulate, batch-sharded over the N ranks with the
bucketed gradient all-reduce overlapped with backward (strong scaling): `value` = samples/s with the inputs resident in HBM,
`e2e` = the same with every micro-batch uploaded from pinned host memory and the step's loss read back, `roofline` = the
tensor-core convolution family of the step, `roofline_hbm` = its memory-bound kernels. The nested `infer` object is
BASELINE.json configs[1]: PU-Net encoder-decoder inference, batch 256 frames per GPU (replicas, weak scaling).
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


def load_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu --set full summary (profiles/), or None."""
    p = os.path.join(ROOT, "profiles", "r01_conv_tc_full.json")
    try:
        d = json.load(open(p))
        l = d["launches"][0]
        unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(l[k]["value"].replace(",", "")) * unit[l[k]["unit"]]
        return tot, "%s — %s (profiles/r01_conv_tc_full.json)" % (l["kernel"], d.get("note", ""))
    except Exception:
        return None, None


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


def build_punet(seed=0):
    """PredictiveUnet with seeded random-init weights (PyTorch default inits; there are no checkpoints offline). The
    constructor loads its frozen U-Net from a checkpoint file exactly like the reference (punet.py:40-50), so a
    seeded stage-0 checkpoint is synthesised first."""
    import tempfile
    from pmoe_b200.model.blocks.unet import UNet
    from pmoe_b200.model.punet import PredictiveUnet
    torch.manual_seed(seed)
    with tempfile.TemporaryDirectory() as td:
        ck = os.path.join(td, "unet.pth")
        torch.save({"unet": UNet(3, 23).state_dict()}, ck)
        net = PredictiveUnet(**dict(PUNET_CFG, model_path=ck))
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():  # non-trivial BatchNorm running statistics, as after training
        for name, buf in net.named_buffers():
            if name.endswith("running_mean"):
                buf.copy_(torch.randn(buf.shape, generator=g) * 0.1)
            elif name.endswith("running_var"):
                buf.copy_(torch.rand(buf.shape, generator=g) * 0.5 + 0.75)
    return net


def build_punet_state(seed=0):
    return {k: v.clone() for k, v in build_punet(seed).state_dict().items()}


def synth_images(B, seed=1234):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(B, 4, 3, 224, 224, generator=g)


# ------------------------------------------------------------------------------------------------ workload definitions
MOE_FWD_GF = 17.963        # forward GFLOP per sample per expert (ResNet18-ECA @224^2 + heads)
MOE_STEM_DGRAD_GF = 0.694  # the first conv needs no data gradient
TRAIN_WORKLOAD = ("moe_train (BASELINE configs[2]): get_model(type='moe', n_experts=%d), %d frames x 3x%dx%d + speed + command, "
                  "fwd + moe_loss + bwd + gradient all-reduce + clip + Adam(amsgrad), random-init weights")


def train_gf_per_sample(K, frames, hw):
    """fwd + dgrad + wgrad per sample (17.963 / 0.694 GF per expert at the conf shape; 77.38 / 8.32 GF for
    configs[4]); other geometries scale the conf figure by the pixel count (approximate)."""
    if (frames, hw) == (12, 448):
        return K * (3 * 77.38 - 8.32)
    return K * (3 * MOE_FWD_GF - MOE_STEM_DGRAD_GF) * (hw * hw) / (224.0 * 224.0)


def synth_train_batch(n, frames, hw, seed):
    g = torch.Generator().manual_seed(seed)
    return {"images": torch.rand(n, frames, 3, hw, hw, generator=g),
            "speed": torch.rand(n, 1, generator=g) * 1.2,
            "command": torch.nn.functional.one_hot(torch.randint(0, 6, (n,), generator=g), 6).float(),
            "control": torch.rand(n, 2, generator=g) * 2 - 1,
            "target": torch.rand(n, 1, generator=g)}


class CpuTrainStep:
    """The reference algorithm of the training step on the host cores: the oracle port (the same ATen CPU operators the
    reference modules dispatch to) of the K-expert mixture, moe_loss, backward, clip_grad_norm_ and torch.optim.Adam(amsgrad)
    (trainer/train_2.py:149-165), on a batch of `batch` samples."""

    def __init__(self, K, batch, frames=4, hw=224):
        from oracle import functional as O
        from pmoe_b200 import conf
        self.O, self.batch = O, batch
        self.cfg = conf.stage2_model_cfg("moe", K, n_frames=frames)
        sd = O.seeded_state_dict(O.make_spec(O.moe_spec, self.cfg), 0)
        self.leaf = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and not k.endswith(("running_mean", "running_var"))
                         else v.clone()) for k, v in sd.items()}
        self.params = [v for v in self.leaf.values() if v.requires_grad]
        self.opt = torch.optim.Adam(self.params, lr=2e-4, betas=(0.9, 0.999), eps=1e-8, amsgrad=True)
        self.d = synth_train_batch(batch, frames, hw, 4321)

    def step(self):
        O, d = self.O, self.d
        self.opt.zero_grad()
        o = O.moe(d["images"], d["speed"], d["command"], self.leaf, "", self.cfg, True)
        loss = O.moe_loss(o[0], o[1], o[2], o[3], d["control"], d["target"].clone(), self.cfg["loss_coefs"])
        loss.backward()
        torch.nn.utils.clip_grad_norm_(self.params, 1.0)
        self.opt.step()
        return float(loss)


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args, rank, world):
    """`--impl reference`: the reference's CPU implementation of the path (oracle port; the reference is pure Python on ATen, and
    /root/reference does not exist on the GPU box) on every host core, each step a bounded sample (batch --cpu-batch) of the
    training workload."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    K, Bs = args.train_experts, args.cpu_batch
    job = CpuTrainStep(K, Bs, args.train_frames, args.train_hw)
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        job.step()
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    total = sum(times)
    val = Bs * len(times) / total
    line = {"impl": "reference", "metric": "train_samples_per_sec", "value": val, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": TRAIN_WORKLOAD % (K, args.train_frames, args.train_hw, args.train_hw), "global_batch": Bs,
                       "note": "bounded sample: batch %d per step instead of %d (samples/s is batch-comparable)" % (Bs, args.train_batch)},
            "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cores, "kind": "port",
                             "sample": "batch %d of the %d-sample step, fwd+loss+bwd+clip+Adam (fp32 ATen/oneDNN, %d threads)"
                                       % (Bs, args.train_batch, cores)},
            "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_train_baseline(K, batch, frames, hw, budget_s=12.0):
    """Bounded CPU sample for the `cpu_baseline` object of the CUDA arm: ~budget_s seconds of the oracle training step."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    job = CpuTrainStep(K, batch, frames, hw)
    job.step()  # warm-up
    n, t0 = 0, time.perf_counter()
    while True:
        job.step()
        n += 1
        dt = time.perf_counter() - t0
        if (n >= 2 and dt >= budget_s) or dt > 60.0:
            break
    return {"value": batch * n / dt, "unit": "samples/s", "cores": cores, "kind": "port",
            "sample": "%d steps x batch %d of the same training step (K=%d experts, fwd+loss+bwd+clip+Adam) = %.1f s of CPU work "
                      "(fp32 ATen/oneDNN, %d threads)" % (n, batch, K, dt, cores)}


def cpu_infer_baseline(batch=2, budget_s=8.0):
    from oracle import functional as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = build_punet_state()
    x = synth_images(batch)
    with torch.no_grad():
        O.punet(x, sd, "", False, 4, 6)  # warm-up
        n, t0 = 0, time.perf_counter()
        while True:
            O.punet(x, sd, "", False, 4, 6)
            n += 1
            dt = time.perf_counter() - t0
            if (n >= 2 and dt >= budget_s) or dt > 60.0:
                break
    return {"value": batch * n / dt, "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": "%d iterations x batch %d of the same PU-Net forward = %.1f s of CPU work (fp32 ATen/oneDNN, %d threads)" % (n, batch, dt, cores)}


# ------------------------------------------------------------------------------------------------ CUDA arm: inference leg
def run_infer_leg(args, rank, world, dev, peaks):
    """BASELINE configs[1]: PU-Net inference, B frames per GPU, replicas. Returns the nested `infer` object (rank 0) or None."""
    from pmoe_b200 import profiler
    from pmoe_b200.infer import pinned_output_like
    B = args.batch
    steps = max(2, min(args.steps, args.infer_steps))
    net = build_punet().to(dev).eval()
    host_in = synth_images(B, seed=1234 + rank).pin_memory()
    x = host_in.to(dev, non_blocking=True)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for _ in range(3):
            y = net(x)
        barrier()
        profiler.reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            y = net(x)
        e1.record()
        barrier()
        launches = profiler.launch_count()
        ms_total = e0.elapsed_time(e1)

        # end-to-end through the module API with host buffers: every step copies its input from pinned host memory and its full
        # fp32 output back to pinned host memory; uploads and downloads are double-buffered on their own streams and every copy
        # of every timed step completes inside the timed region.
        copy_stream = torch.cuda.Stream(device=dev)
        h2d_stream = torch.cuda.Stream(device=dev)
        host_out = pinned_output_like(B, 6, 23, 224, 224)
        try:
            host_out2 = pinned_output_like(B, 6, 23, 224, 224) if world <= 4 else host_out
        except RuntimeError:
            host_out2 = host_out
        houts = [host_out, host_out2]
        xbufs = [torch.empty_like(x), torch.empty_like(x)]
        copied, uploaded, consumed = [None, None], [None, None], [None, None]

        def upload(i):
            j = i % 2
            if consumed[j] is not None:
                h2d_stream.wait_event(consumed[j])
            with torch.cuda.stream(h2d_stream):
                xbufs[j].copy_(host_in, non_blocking=True)
                uploaded[j] = torch.cuda.Event()
                uploaded[j].record()

        def e2e_step(i, last):
            j = i % 2
            cur = torch.cuda.current_stream()
            if copied[j] is not None:
                cur.wait_event(copied[j])
            cur.wait_event(uploaded[j])
            if not last:
                upload(i + 1)
            net(xbufs[j], host_out=houts[j], copy_stream=copy_stream)
            consumed[j] = torch.cuda.Event()
            consumed[j].record()
            with torch.cuda.stream(copy_stream):
                copied[j] = torch.cuda.Event()
                copied[j].record()

        upload(0)
        e2e_step(0, True)
        torch.cuda.current_stream().wait_stream(copy_stream)
        barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        h2d_stream.wait_event(e2)
        upload(0)
        for i in range(steps):
            e2e_step(i, i == steps - 1)
        torch.cuda.current_stream().wait_stream(copy_stream)
        e3.record()
        barrier()
        ms_e2e = e2.elapsed_time(e3)

        profiler.enable_events(True)
        net(x)
        torch.cuda.synchronize()
        prof = profiler.summary()
        profiler.enable_events(False)

    h2d_bytes, d2h_bytes = host_in.numel() * 4, B * 6 * 23 * 224 * 224 * 4
    del net, x, y, host_out, host_out2, houts, xbufs
    torch.cuda.empty_cache()
    t = torch.tensor([ms_total, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms_total, ms_e2e = t.tolist()
    if rank != 0:
        return None
    ms_step = ms_total / steps
    conv = prof.get("conv_tc", {"ms": 0.0, "flops": 0.0, "launches": 0})
    achieved = conv["flops"] / (conv["ms"] * 1e-3) / 1e12 if conv["ms"] > 0 else 0.0
    traffic, traffic_launch = load_traffic()
    return {"metric": "infer_frames_per_sec", "value": world * B * steps / (ms_total / 1e3), "unit": "frames/s", "steps": steps,
            "ms_per_step": ms_step, "scaling": "weak",
            "config": {"workload": "punet_infer (BASELINE configs[1]): PredictiveUnet 4->6 frames, 3x224x224, eval, random-init weights",
                       "batch_per_gpu": B, "global_batch": B * world, "parallelism": "replicas x%d" % world,
                       "whole_step_tflops": PUNET_GF_PER_SAMPLE * B / ms_step / 1e3},
            "e2e": {"value": world * B * steps / (ms_e2e / 1e3), "unit": "frames/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": d2h_bytes},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": "conv_tc_kernel (tcgen05 implicit GEMM)", "achieved": achieved,
                         "peak": peaks["bf16_sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16_sustained"],
                         "traffic": traffic, "traffic_launch": traffic_launch, "launches_per_step": conv["launches"],
                         "kernel_ms_per_step": conv["ms"], "share_of_step": conv["ms"] / ms_step if ms_step > 0 else None}}


# ------------------------------------------------------------------------------------------------ CUDA arm: training leg (primary)
TC_KINDS = ("conv_tc", "conv_wgrad_tc")


def run_train_leg(args, rank, world, dev, peaks, local_rank):
    """BASELINE configs[2]/3a: full mixture (K experts, each ResNet18-ECA encoder + gating + action/speed heads) training
    step at a GLOBAL batch of --train-batch, batch-sharded over the ranks (strong scaling): forward, moe_loss, backward with
    the bucketed gradient all-reduce overlapped, grad-norm clip folded into the fused Adam(amsgrad) step. A rank whose shard
    exceeds --train-micro samples accumulates micro-batches (BatchNorm statistics are then per micro-batch, exactly what the
    same shard split over more ranks computes). The micro-step (forward + loss + backward + all-reduces, including the re-pack
    of every weight operand from the live parameters) is ONE captured CUDA graph."""
    from pmoe_b200 import conf, dp, loss as L, optim, profiler
    from pmoe_b200 import train as _train
    from pmoe_b200.model.moe import get_model
    K, Bg = args.train_experts, args.train_batch
    if Bg % world:
        raise SystemExit("global batch %d not divisible by %d ranks" % (Bg, world))
    per = Bg // world
    micro = min(per, args.train_micro)
    if per % micro:
        raise SystemExit("per-rank batch %d not a multiple of the micro-batch %d" % (per, micro))
    n_micro = per // micro
    frames, hw = args.train_frames, args.train_hw
    torch.manual_seed(0)
    cfg = conf.stage2_model_cfg(args.train_type, K, n_frames=frames)
    model = get_model(cfg).to(dev).train()
    wrapped = dp.DataParallel(model, gradient_as_bucket_view=(n_micro == 1)) if world > 1 else model
    opt = optim.FusedAdam([p for p in model.parameters() if p.requires_grad], lr=2e-4, betas=(0.9, 0.999), eps=1e-8, amsgrad=True)
    host = {k: v.pin_memory() for k, v in synth_train_batch(per, frames, hw, 4321 + rank).items()}
    resident = {k: v.to(dev) for k, v in host.items()}   # the device-resident copy `value` is measured on

    def loss_of(d):
        dist_, sp = wrapped(d["images"], d["speed"], d["command"])
        return L.moe_loss(dist_, sp, d["control"], d["target"].clone(), cfg.loss_coefs) / n_micro

    graph, mode, launches_per_micro = None, "eager", 0
    static = {k: torch.empty((micro,) + tuple(v.shape[1:]), dtype=v.dtype, device=dev) for k, v in host.items()}
    for k, v in resident.items():
        static[k].copy_(v[:micro])
    if not args.no_graph:
        try:
            torch.distributions.Distribution.set_default_validate_args(False)  # argument validation synchronises
            _train.DROPOUT_STEP = torch.zeros(1, dtype=torch.int64, device=dev)  # bumped before every replay: fresh dropout masks
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):  # warm-up on a side stream: lazy initialisation, index maps of the packed operands, allocator pool
                    opt.zero_grad(set_to_none=True)
                    loss_of(static).backward()
            torch.cuda.current_stream().wait_stream(side)
            for p in model.parameters():  # static, zeroed .grad: the captured AccumulateGrad adds in place
                if p.grad is not None:
                    p.grad.zero_()
            graph = torch.cuda.CUDAGraph()
            profiler.reset()
            with torch.cuda.graph(graph):  # under DP the bucketed NCCL all-reduces of the tape are captured as well
                static_loss = loss_of(static)
                static_loss.backward()
            launches_per_micro = profiler.launch_count()
            mode = "cuda_graph"
        except Exception as ex:  # capture is an optimisation: fall back to the eager tape, loudly
            import traceback
            print("bench: CUDA-graph capture of the training micro-step failed, running the eager tape instead:\n" + traceback.format_exc(),
                  file=sys.stderr, flush=True)
            graph, mode = None, "eager (graph capture failed: %s)" % str(ex).splitlines()[0][:160]
            static_loss = None
            _train.DROPOUT_STEP = None
            torch.cuda.synchronize()
            import gc
            gc.collect()
            torch.cuda.empty_cache()

    copy_stream = torch.cuda.Stream(device=dev)
    stage = {k: torch.empty_like(v) for k, v in static.items()}
    ev_ready, ev_free = torch.cuda.Event(), torch.cuda.Event()

    def prefetch(m):
        """host -> staging buffer on the copy stream (overlaps the compute of the previous micro-batch)"""
        sl = slice(m * micro, (m + 1) * micro)
        copy_stream.wait_event(ev_free)
        with torch.cuda.stream(copy_stream):
            for k, v in host.items():
                stage[k].copy_(v[sl], non_blocking=True)
            ev_ready.record(copy_stream)

    def micro_step(src):
        for k in static:
            static[k].copy_(src[k], non_blocking=True)
        if graph is not None:
            _train.DROPOUT_STEP.add_(1)
            graph.replay()
            return static_loss.detach()
        loss = loss_of(static)
        loss.backward()
        return loss.detach()

    def zero_grads():
        gs = [p.grad for p in model.parameters() if p.grad is not None]
        if gs:
            torch._foreach_zero_(gs)

    def step_resident():
        """inputs already in HBM: `value`"""
        zero_grads()
        total = None
        for m in range(n_micro):
            sl = slice(m * micro, (m + 1) * micro)
            lv = micro_step({k: v[sl] for k, v in resident.items()})
            total = lv.clone() if total is None else total + lv
        opt.step(max_grad_norm=1.0)
        return total

    def step_e2e():
        """host buffers: every micro-batch is uploaded from pinned memory inside the timed region (pipelined behind the previous
        micro-batch's compute) and the step's loss is read back"""
        cur = torch.cuda.current_stream()
        zero_grads()
        total = None
        for m in range(n_micro):
            cur.wait_event(ev_ready)
            lv_src = stage
            for k in static:
                static[k].copy_(lv_src[k], non_blocking=True)
            ev_free.record(cur)
            prefetch((m + 1) % n_micro)
            if graph is not None:
                _train.DROPOUT_STEP.add_(1)
                graph.replay()
                lv = static_loss.detach()
            else:
                loss = loss_of(static)
                loss.backward()
                lv = loss.detach()
            total = lv.clone() if total is None else total + lv
        opt.step(max_grad_norm=1.0)
        return float(total.item())   # D2H read of the step's loss

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    K_steps, W = args.steps, max(args.warmup, 3)
    for _ in range(W):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    profiler.reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K_steps):
        loss = step_resident()
    e1.record()
    barrier()
    clocks = sampler.stop()
    launches = profiler.launch_count() + launches_per_micro * n_micro * K_steps
    ms_value = e0.elapsed_time(e1)
    loss_value = float(loss.item())

    ev_free.record(torch.cuda.current_stream())
    prefetch(0)
    step_e2e()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(K_steps):
        step_e2e()
    torch.cuda.current_stream().wait_stream(copy_stream)  # the upload issued by the last step lands inside the timed region
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)

    # per-launch CUDA-event pass over ONE eager micro-step (outside the timed regions): time, algorithmic FLOPs and bytes of
    # every launch -> roofline of the tensor-core convolution family and of the memory-bound family
    prof = {}
    graph = static_loss = None   # frees the graph's private pool (~100 GB of activations) for the eager pass below
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    # (every rank runs it: under data parallelism the pass contains the bucket all-reduces, which all ranks must enter)
    _train.DROPOUT_STEP = None
    zero_grads()
    profiler.enable_events(True)
    loss_of(static).backward()
    torch.cuda.synchronize()
    prof = profiler.summary()
    profiler.enable_events(False)

    t = torch.tensor([ms_value, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms_value, ms_e2e = t.tolist()
    stats = getattr(wrapped, "last_stats", None) if world > 1 else None
    params = sum(p.numel() for p in model.parameters())
    del model, wrapped, opt, static, stage, resident
    _train.DROPOUT_STEP = None
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    ms_step = ms_value / K_steps
    gf = train_gf_per_sample(1 if args.train_type == "moe_shared" else K, frames, hw)   # moe_shared: ONE shared encoder
    tc_ms = sum(prof[k]["ms"] for k in TC_KINDS if k in prof)
    tc_fl = sum(prof[k]["flops"] for k in TC_KINDS if k in prof)
    tc_n = sum(prof[k]["launches"] for k in TC_KINDS if k in prof)
    hbm = {k: v for k, v in prof.items() if v["bytes"] > 0 and k not in TC_KINDS}
    hbm_ms, hbm_by = sum(v["ms"] for v in hbm.values()), sum(v["bytes"] for v in hbm.values())
    achieved = tc_fl / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
    hbm_ach = hbm_by / (hbm_ms * 1e-3) / 1e9 if hbm_ms > 0 else 0.0
    micro_ms = ms_step / n_micro
    traffic, traffic_launch = load_traffic()
    line = {
        "metric": "train_samples_per_sec", "value": Bg / (ms_step / 1e3), "unit": "samples/s", "n_gpus": world, "steps": K_steps,
        "warmup": W, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": (TRAIN_WORKLOAD % (K, frames, hw, hw)).replace("type='moe'", "type='%s'" % args.train_type), "global_batch": Bg, "batch_per_gpu": per, "micro_batch": micro,
                   "parallelism": "dp%d (batch-sharded, bucketed NCCL all-reduce inside the captured graph)" % world if world > 1 else "1 GPU",
                   "launch_mode": mode, "params": params,
                   "l2": "activations of one micro-batch (%.0f GB) exceed the 126 MB L2 many times over" % (77e-3 * micro * K),
                   "whole_step_tflops_per_gpu": gf * per / ms_step / 1e3, "loss": loss_value},
        "e2e": {"value": Bg / (ms_e2e / K_steps / 1e3), "unit": "samples/s",
                "h2d_bytes_per_step": sum(v.numel() * v.element_size() for v in host.values()), "d2h_bytes_per_step": 4},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "conv_tc* + conv_wgrad_tc* (tcgen05 implicit GEMM: forward, data gradient, weight gradient)",
                     "achieved": achieved, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16_sustained"],
                     "peak_source": "%s bf16_tflops_sustained" % peaks["src"], "traffic": traffic, "traffic_launch": traffic_launch,
                     "launches_per_micro_step": tc_n, "kernel_ms_per_micro_step": tc_ms,
                     "share_of_step": tc_ms / micro_ms if micro_ms > 0 else None,
                     "measured": "per-launch CUDA events over one eager micro-step of %d samples" % micro},
        "roofline_hbm": {"bound": "hbm", "kernel": "memory-bound family: " + ", ".join(sorted(hbm, key=lambda k: -hbm[k]["ms"])[:6]),
                         "achieved": hbm_ach, "peak": peaks["hbm"], "unit": "GB/s", "frac": hbm_ach / peaks["hbm"],
                         "traffic": load_hbm_traffic(), "kernel_ms_per_micro_step": hbm_ms,
                         "share_of_step": hbm_ms / micro_ms if micro_ms > 0 else None,
                         "per_kernel": {k: {"ms": v["ms"], "gbs": v["bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] > 0 else 0.0,
                                            "launches": v["launches"]} for k, v in hbm.items()}},
        "other_kernels_ms_per_micro_step": {k: v["ms"] for k, v in prof.items() if k not in TC_KINDS and k not in hbm},
    }
    if stats:
        line["allreduce"] = stats
    return line


def load_hbm_traffic():
    """dram bytes per launch of the dominant memory-bound kernel (bn_bwd_apply_fast) from the committed ncu --set full summary."""
    for name in ("r02_bn_bwd_apply_full.json", "r01_bn_bwd_apply_full.json"):
        p = os.path.join(ROOT, "profiles", name)
        try:
            d = json.load(open(p))
            l = d["launches"][0]
            unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            return sum(float(l[k]["value"].replace(",", "")) * unit[l[k]["unit"]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        except Exception:
            continue
    return None


def run_cuda(args, rank, world, local_rank):
    from pmoe_b200 import _lib
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    _lib.check(_lib.lib().pmoe_device_check(), "device_check")
    peaks = load_peaks()
    line = run_train_leg(args, rank, world, dev, peaks, local_rank)
    infer = None if args.no_infer else run_infer_leg(args, rank, world, dev, peaks)
    if rank != 0:
        return
    if infer is not None:
        line["infer"] = infer
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_train_baseline(args.train_experts, args.cpu_batch, args.train_frames, args.train_hw)
        if infer is not None:
            infer["cpu_baseline"] = cpu_infer_baseline(2)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="inference leg: frames per GPU per step")
    ap.add_argument("--infer-steps", type=int, default=5, help="inference leg: timed steps (at most --steps)")
    ap.add_argument("--cpu-batch", type=int, default=4, help="bounded CPU sample size (BASELINE configs[0]: batch 4)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-infer", action="store_true", help="skip the nested inference leg (BASELINE configs[1])")
    ap.add_argument("--train-batch", type=int, default=512, help="GLOBAL training batch, sharded over the ranks")
    ap.add_argument("--train-experts", type=int, default=6)
    ap.add_argument("--train-type", default="moe", choices=["moe", "moe_alt", "moe_shared"],
                    help="mixture flavour of the training leg (moe = BASELINE configs[2] / SURVEY 3a; moe_shared = 3b)")
    ap.add_argument("--train-micro", type=int, default=256, help="largest micro-batch one rank runs at once")
    ap.add_argument("--train-frames", type=int, default=4, help="frames stacked per sample (12 = 3 cameras x 4: BASELINE configs[4])")
    ap.add_argument("--train-hw", type=int, default=224, help="frame height = width (448: BASELINE configs[4])")
    ap.add_argument("--no-graph", action="store_true", help="training leg: issue every launch from Python instead of one CUDA graph")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=180))
    try:
        run_cuda(args, rank, world, local_rank)
    finally:
        if world > 1:
            torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
