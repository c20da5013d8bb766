#!/usr/bin/env python
"""Benchmark of the hot path named by BASELINE.json: the Neural Module Network executor, forward + backward.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): module_training.yml at batch 256 per GPU, synthetic 14x14x1024 features
and ground-truth-style CLEVR programs (seeded grammar, <= 40 tokens), reference-shaped random-init weights.
One step = NeuralModuleNetwork.forward(features, programs, answers) + loss.mean().backward() through the
public nn.Module API (program compilation on the host included); N > 1 adds the gradient all-reduce (NCCL).

Prints ONE JSON line (rank 0).  `value` is measured with inputs resident in HBM, `e2e` with pinned host
inputs copied in and the loss read back every step.  `roofline` is the tcgen05 conv kernel (the dominant
kernel) timed with CUDA events around every launch; `cpu_baseline` is the CPU oracle (a port of the
reference's PyTorch path) on a bounded sample.  `--impl reference` times that CPU path alone.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "questions/sec (NMN executor fwd+bwd, batch 256 per GPU)"
UNIT = "questions/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--length", type=int, default=40)
    ap.add_argument("--cpu-sample", type=int, default=16, help="rows of the workload the CPU baseline runs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary ProgramGenerator / joint-step legs")
    return ap.parse_args()


def workload_config(args):
    return {
        "workload": "module_training.yml batch=256/GPU: NMN stem + module executor + classifier, fwd+bwd "
                    "(BASELINE.json configs[1])",
        "batch_per_gpu": args.batch, "program_length": args.length, "features": [1024, 14, 14],
        "programs": "seeded CLEVR template grammar (probnmn_clevr_b200/synthetic.py), ground-truth style",
        "weights": "reference shapes, He-normal, seed 0",
        "step": "forward + loss.mean().backward(); no optimizer (executor-only config)",
        "l2": "inputs larger than L2 (205 MB of features per step, two alternating batches)",
        "programs": "host-resident int64 (consumed by the host-side program compiler); features/answers in HBM",
        "parallelism": f"dp{args.gpus}",
    }


# ---------------------------------------------------------------------------------------------------------
# CPU path (oracle port of the reference): cpu_baseline and --impl reference
# ---------------------------------------------------------------------------------------------------------
def cpu_reference_step(sd, vocab, feats, programs, answers):
    from oracle import nmn_oracle
    for p in sd.values():
        p.grad = None
    out = nmn_oracle.nmn_forward(sd, vocab, feats, programs, answers)
    out["loss"].mean().backward()
    return float(out["loss"].detach().mean())


def cpu_inputs(args, rows, seed=0):
    from probnmn_clevr_b200.synthetic import ProgramSampler, make_answers, make_features, make_nmn_state_dict
    from probnmn_clevr_b200.vocabulary import Vocabulary
    vocab = Vocabulary.clevr()
    sd = {k: v.requires_grad_(True) for k, v in make_nmn_state_dict(vocab, 0).items()}
    programs = ProgramSampler(vocab, seed=seed).sample(args.batch, args.length)[:rows]
    return vocab, sd, make_features(rows, seed), programs, make_answers(rows, seed)


def run_reference(args):
    """The reference's CPU PyTorch path (oracle port; /root/reference does not exist on the GPU box)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count())
    rows = args.cpu_sample
    vocab, sd, feats, programs, answers = cpu_inputs(args, rows)
    for _ in range(min(args.warmup, 1)):
        cpu_reference_step(sd, vocab, feats, programs, answers)
    steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_reference_step(sd, vocab, feats, programs, answers)
    dt = time.perf_counter() - t0
    value = rows * steps / dt
    sample = f"{rows} rows of the batch-{args.batch} workload per step, fwd+bwd, {steps} steps"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """SM clock / power / throttle reasons sampled DURING the timed region through NVML (nvidia_ml_py; a few hundred
    samples per second), falling back to one nvidia-smi query per 100 ms."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.source = index, [], False, "nvidia-smi"
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml, self.source = pynvml, "nvml"
        except Exception:
            self.nvml = None

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    n = self.nvml
                    mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                    try:
                        mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                    except Exception:
                        mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                    watts = n.nvmlDeviceGetPowerUsage(self.handle) / 1e3
                    self.samples.append((mhz, self.max_mhz, watts, mask))
                    time.sleep(0.004)
                    continue
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 7:
                    mask = sum(bit for (name, bit), v in zip(self.BITS.items(), f[3:7]) if v.lower().startswith("active"))
                    self.samples.append((float(f[0]), float(f[1]), float(f[2]), mask))
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(s[0] for s in self.samples)
        reasons = [name for name, bit in self.BITS.items() if any(s[3] & bit for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.samples[0][1], "reasons": reasons,
                "samples": len(sm), "power_w_max": max(s[2] for s in self.samples), "source": self.source}


# ---------------------------------------------------------------------------------------------------------
# ours
# ---------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from probnmn_clevr_b200 import _lib as L
    from probnmn_clevr_b200.nmn import NeuralModuleNetwork
    from probnmn_clevr_b200.synthetic import ProgramSampler, make_answers, make_features, make_nmn_state_dict
    from probnmn_clevr_b200.vocabulary import Vocabulary

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.allow_tf32 = False  # classifier stays fp32, as in the reference
    torch.backends.cuda.matmul.allow_tf32 = False

    vocab = Vocabulary.clevr()
    model = NeuralModuleNetwork(vocab)
    model.load_state_dict(make_nmn_state_dict(vocab, 0))
    model = model.to(dev).train()

    # two alternating batches per rank; host copies pinned for the end-to-end leg
    host = []
    for i in range(2):
        seed = 100 * rank + i
        host.append((make_features(args.batch, seed).pin_memory(),
                     ProgramSampler(vocab, seed=seed).sample(args.batch, args.length).pin_memory(),
                     make_answers(args.batch, seed).pin_memory()))
    # `value` leg: features / answers resident in HBM; the programs stay in (pinned) host memory because that is
    # where the executor consumes them (its program compiler runs on the host)
    resident = [(h[0].to(dev), h[1], h[2].to(dev)) for h in host]

    if world > 1 and os.environ.get("PNMN_NO_GRAD_OVERLAP") is None:
        model.enable_gradient_overlap()  # classifier gradients are all-reduced underneath the executor's backward

    def step(feats, programs, answers):
        model.zero_grad(set_to_none=True)
        out = model(feats, programs, answers)
        loss = out["loss"].mean()
        loss.backward()
        if world > 1:
            model.allreduce_gradients()
        return loss

    def timed(fn, steps, finish=None):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        t_host = time.perf_counter()
        for i in range(steps):
            fn(i)
        if finish is not None:
            finish()
        timed.host_issue_ms = (time.perf_counter() - t_host) * 1e3  # host time to ISSUE the steps (no device sync inside)
        b.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    lookahead = os.environ.get("PNMN_NO_PRECOMPILE") is None  # diagnostics: compile every plan inline

    # The host may run at most two steps ahead of the device (a training loop reads its loss / metrics with that kind of
    # lag): without a bound the issuing thread gets ~10 steps ahead within the timed region, every staging pool and the
    # caching allocator grow under it (cudaHostAlloc / cudaMalloc synchronise), and the measurement becomes erratic.
    run_ahead = [torch.cuda.Event() for _ in range(3)]
    issued = {"n": 0, "wait_s": 0.0}

    def throttle():
        k = issued["n"]
        if k >= 2:
            t0 = time.perf_counter()
            run_ahead[(k - 2) % 3].synchronize()
            issued["wait_s"] += time.perf_counter() - t0

    def step_issued():
        run_ahead[issued["n"] % 3].record()
        issued["n"] += 1

    def resident_step(i):
        # software pipeline of the host side: the program compiler works on the batches after this one (helper threads)
        # while this thread issues the current step -- the same look-ahead an input pipeline gives the feature copy
        throttle()
        f, p, a = resident[i % 2]
        if os.environ.get("PNMN_DEBUG_INPUTS"):
            bad = ((a < 0) | (a >= 28))
            if bool(bad.any()):
                raise RuntimeError(f"rank {rank} resident step {i}: {int(bad.sum())} bad answers")
        model.zero_grad(set_to_none=True)
        out = model(f, p, a)
        if lookahead:
            model.precompile(resident[(i + 2) % 2][1])  # two steps ahead: two plans in flight on two helper threads
        loss = out["loss"].mean()
        loss.backward()
        if world > 1:
            model.allreduce_gradients()
        step_issued()

    # end-to-end leg: every step's features / answers start in PINNED HOST memory; the copy of step i+1 is issued on a side
    # stream before step i computes (probnmn_clevr_b200/feed.py), so all K copies sit inside the timed region but overlap
    # with compute; the loss is read back (device -> host) every step
    from probnmn_clevr_b200.feed import DevicePrefetcher
    prefetch_ahead = int(os.environ.get("PNMN_PREFETCH_AHEAD", "1"))  # batches copied ahead of the one being computed
    feed = DevicePrefetcher(dev, depth=prefetch_ahead + 2)
    feed_state = {"next": 0}

    def feed_upto(last):
        while feed_state["next"] <= last:
            k = feed_state["next"]
            feed.submit(k, (host[k % 2][0], host[k % 2][2]))
            feed_state["next"] += 1
    e2e_total = {"n": 0, "first": 0}

    n_slots = args.steps + max(args.warmup, 8) + 8
    loss_host = torch.zeros(n_slots, dtype=torch.float32).pin_memory()
    loss_events = [torch.cuda.Event() for _ in range(n_slots)]
    loss_values = []

    def read_loss(i):
        loss_events[i].synchronize()
        loss_values.append(float(loss_host[i]))

    e2e_trace = [] if os.environ.get("PNMN_E2E_TRACE") else None  # diagnostics: per-step device / host timestamps
    e2e_lag = int(os.environ.get("PNMN_E2E_LAG", "1"))  # a step's loss is read this many steps later
    e2e_variant = os.environ.get("PNMN_E2E_VARIANT", "")  # diagnostics only: "nocopy" / "noread" drop one part of the leg

    def e2e_step(i):
        if e2e_variant == "nocopy":
            model.zero_grad(set_to_none=True)
            out = model(resident[i % 2][0], host[i % 2][1], resident[i % 2][2])
            loss = out["loss"].mean()
            loss.backward()
            loss_host[i:i + 1].copy_(loss.detach().reshape(1), non_blocking=True)
            loss_events[i].record()
            if i > 0:
                read_loss(i - 1)
            return
        if e2e_trace is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            e2e_trace.append((i, ev, time.perf_counter()))
        if feed_state["next"] <= i:  # (first step of a sequence: nothing was prefetched)
            feed_state["next"] = i
            feed_upto(i)
        f, a = feed.get(i)
        if os.environ.get("PNMN_DEBUG_INPUTS"):  # diagnostics: the batch as the device sees it (synchronises)
            bad = ((a < 0) | (a >= 28))
            if bool(bad.any()):
                idx = bad.nonzero().flatten()
                raise RuntimeError(f"rank {rank} e2e step {i}: {int(bad.sum())} bad answers, rows {idx[:4].tolist()}..{idx[-4:].tolist()}, "
                                   f"values {a[idx[:4]].tolist()}, host ok {bool(((host[i % 2][2] >= 0) & (host[i % 2][2] < 28)).all())}")
        model.zero_grad(set_to_none=True)
        out = model(f, host[i % 2][1], a)
        if lookahead and i + 2 < e2e_total["n"]:
            model.precompile(host[(i + 2) % 2][1])  # two steps ahead (the input pipeline knows its next two batches)
        # (queued after the forward pass has been issued; measured alternatives -- first thing in the step, two batches ahead,
        # loss read two steps later -- were no faster on average and more erratic: profiles/r1c_notes.md)
        feed_upto(min(i + prefetch_ahead, e2e_total["n"] - 1))
        loss = out["loss"].mean()
        loss.backward()
        if world > 1:
            model.allreduce_gradients()
        # every step's loss goes device -> pinned host; it is READ one step later (the usual logging lag of a training
        # loop) so that the host can prepare step i+1 while step i still runs; the last one is read by e2e_finish()
        if e2e_variant != "nod2h":  # (diagnostics: "nod2h" keeps the per-step synchronisation but drops the 4-byte copy)
            loss_host[i:i + 1].copy_(loss.detach().reshape(1), non_blocking=True)
        loss_events[i].record()
        if i - e2e_lag >= e2e_total["first"] and e2e_variant != "noread":
            read_loss(i - e2e_lag)

    def e2e_finish():
        if e2e_variant == "noread":
            for j in range(e2e_total["first"], e2e_total["n"]):
                read_loss(j)
            return
        for j in range(max(e2e_total["first"], e2e_total["n"] - e2e_lag), e2e_total["n"]):
            read_loss(j)

    for i in range(max(args.warmup, 8)):  # (at least 8: the pinned staging pool of the plan uploads settles during warm-up)
        resident_step(i)
    # everything allocated so far (model, vocabulary, torch internals) leaves the cyclic collector's working set: a full
    # collection over it costs milliseconds, which at ~2 ms of host work per step shows up as a device bubble
    import gc
    gc.collect()
    gc.freeze()
    host_ms = (ctypes.c_double * 4)()
    L.lib().pnmn_debug_host_times(host_ms)
    sampler = ClockSampler(local)
    sampler.start()
    L.lib().pnmn_launch_count(1)
    issued["wait_s"] = 0.0
    ms = timed(resident_step, args.steps)
    host_issue_ms = timed.host_issue_ms - issued["wait_s"] * 1e3  # without the time spent waiting for the device
    own_launches = int(L.lib().pnmn_launch_count(1))
    sampler.stop_flag = True
    L.lib().pnmn_debug_host_times(host_ms)
    host_ms_per_step = {"plan_create": host_ms[0] / args.steps, "forward_call": host_ms[1] / args.steps,
                        "backward_call": host_ms[2] / args.steps,
                        "issue_total": host_issue_ms / args.steps,  # wall time the host needs to issue one step
                        "run_ahead_wait": issued["wait_s"] * 1e3 / args.steps}  # waiting for step i-2 (run-ahead bound)
    stats = model.last_plan_stats
    # warm-up of the end-to-end leg (W steps like the resident leg: the prefetcher's device ring, the upload stream's
    # allocator pool and the pinned staging buffers are first touched here)
    # The pipeline stays warm across the warm-up / timed boundary (one continuous sequence of steps): when the timer starts,
    # the features of the first timed step and the plans of the first two are already in flight, as they are for every
    # later step; each timed step issues the copy of the next one.
    W = max(args.warmup, 8)
    e2e_total["n"], e2e_total["first"] = W + args.steps, 0
    for i in range(W):
        e2e_step(i)
    loss_values.clear()
    ms_e2e = timed(lambda j: e2e_step(W + j), args.steps, e2e_finish)
    assert len(loss_values) == args.steps + e2e_lag and all(v == v for v in loss_values), "every step's loss must have been read back"
    if e2e_trace is not None and rank == 0:
        torch.cuda.synchronize()
        for (i0, a0, h0), (i1, a1, h1) in zip(e2e_trace, e2e_trace[1:]):
            print(f"e2e step {i0}: device start->start {a0.elapsed_time(a1):6.2f} ms, host {1e3 * (h1 - h0):6.2f} ms", file=sys.stderr)
        e2e_trace.clear()

    # the end-to-end leg moves 205.5 MB of fp32 features per step: what the host -> device link alone sustains for that copy
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(3):
        resident[0][0].copy_(host[0][0], non_blocking=True)
    ev1.record()
    torch.cuda.synchronize()
    h2d_ms = ev0.elapsed_time(ev1) / 3

    value = world * args.batch * args.steps / (ms * 1e-3)
    e2e_value = world * args.batch * args.steps / (ms_e2e * 1e-3)
    h2d = sum(t.numel() * t.element_size() for t in host[0])

    # per-kernel device time of a few profiled steps (CUDA events around every launch, same stream)
    lib = L.lib()
    prof_steps = min(args.steps, 5)
    kinds = ["elementwise", "conv_tc<2,2>", "conv_tc<1,3>", "wgrad_tc", "bias_grad", "pack_weights", "nchw_to_planes", "other"]
    lib.pnmn_profile_enable(1)
    for i in range(prof_steps):
        resident_step(i)
    pms, pln = (ctypes.c_double * 8)(), (ctypes.c_int64 * 8)()
    lib.pnmn_profile_read(pms, pln)
    lib.pnmn_profile_enable(0)
    kernel_ms = {k: pms[i] / prof_steps for i, k in enumerate(kinds)}
    # the same kernels while the next batch's 205 MB host -> device copy is in flight (end-to-end leg)
    e2e_total["n"], e2e_total["first"] = prof_steps, 0
    feed_state["next"] = 0
    loss_values.clear()
    lib.pnmn_profile_enable(1)
    for i in range(prof_steps):
        e2e_step(i)
    e2e_finish()
    lib.pnmn_profile_read(pms, pln)
    lib.pnmn_profile_enable(0)
    e2e_kernel_ms = {k: pms[i] / prof_steps for i, k in enumerate(kinds) if pms[i] > 0}
    conv_flops = stats[8] + stats[10]  # forward + dgrad FLOPs executed by conv_tc<2,2> per step
    conv_ms = kernel_ms["conv_tc<2,2>"]
    achieved = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    extras = {} if args.no_extras else extra_legs(args, model, vocab, dev, resident, timed)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("bf16_tflops_sustained", 1590.0)
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1590 (of fallback)"

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 8), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f16", "data": "synthetic", "config": workload_config(args),
        "clocks": sampler.summary(),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps,
                "kernel_ms_per_step": e2e_kernel_ms,  # the library's kernels with the feature copy in flight
                "h2d_copy_alone_ms": h2d_ms,  # bare pinned -> device copy of one step's features: the floor of this leg
                "pipeline": "pinned host buffers; the copy of step i+1 runs on a side stream during step i (feed.DevicePrefetcher); "
                            "every step's loss is copied to pinned host memory and read one step later (all K reads inside the timed region); the pipeline is warm when the timer starts (warm-up and timed steps are one continuous sequence)"},
        "gpu_launches": own_launches,
        "roofline": {
            "bound": "tensor", "kernel": "exec_kernel (persistent tcgen05 kind::f16 shift-GEMM executor: forward + dgrad launches)",
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "peak_source": peak_src,
            "note": "fp16 operands, fp32 accumulation in TMEM; peak is the bf16 cuBLAS figure; the kernel's time includes "
                    "its CUDA-core tasks and dependency waits",
            "flops_per_step": conv_flops, "kernel_ms_per_step": conv_ms, "launches_per_step": pln[1] / prof_steps,
            "traffic": traffic_per_launch(),
        },
        "kernel_ms_per_step": kernel_ms, "host_ms_per_step": host_ms_per_step,
        "plan": {"valid_programs": stats[0], "conv3x3_instances": stats[1], "module_tokens": stats[2],
                 "forward_launches": stats[3], "backward_launches": stats[4], "wgrad_flops": stats[12]},
        "classifier_math": {"split": "split bf16 x2 (one cuBLAS tensor-core GEMM over the 3x contraction, fp32 accumulate)",
                            "tf32": "tf32 (cuBLAS/cuDNN)", "ieee": "ieee fp32 (cuBLAS/cuDNN)"}[model.classifier_math],
    }
    line.update(extras)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count())
        rows = args.cpu_sample
        cvocab, sd, feats, programs, answers = cpu_inputs(args, rows)
        cpu_reference_step(sd, cvocab, feats, programs, answers)
        best = 1e30
        for _ in range(2):
            t0 = time.perf_counter()
            cpu_reference_step(sd, cvocab, feats, programs, answers)
            best = min(best, time.perf_counter() - t0)
        line["cpu_baseline"] = {"value": rows / best, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                "sample": f"{rows} rows of the same batch, fwd+bwd, best of 2 after 1 warm-up"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def traffic_per_launch():
    """dram__bytes_read.sum + dram__bytes_write.sum of exec_kernel per launch, from the committed ncu --set full capture
    of the same workload (profiles/<round>_traffic.json, written by scripts/summarize_profiles.py); None if absent."""
    try:
        files = sorted(f for f in os.listdir(os.path.join(ROOT, "profiles")) if f.endswith("_traffic.json"))
        return json.load(open(os.path.join(ROOT, "profiles", files[-1])))["exec_kernel_bytes_per_launch"]
    except Exception:
        return None


def extra_legs(args, nmn, vocab, dev, resident, timed):
    """Secondary numbers (not the headline): BASELINE.json configs[2] (ProgramGenerator, question_coding mix) and
    configs[3] (joint step: ProgramGenerator sampling + NMN + REINFORCE-weighted objective) at the same batch."""
    from probnmn_clevr_b200.seq2seq import ProgramGenerator
    from probnmn_clevr_b200.synthetic import ProgramSampler, make_questions, make_seq2seq_state_dict
    B = args.batch
    pg = ProgramGenerator(vocab)
    pg.load_state_dict(make_seq2seq_state_dict(vocab.get_vocab_size("questions"), vocab.get_vocab_size("programs"), seed=0))
    pg = pg.to(dev).train()
    questions = make_questions(B, vocab.get_vocab_size("questions"), seed=0, max_length=40).to(dev)
    gt_programs = ProgramSampler(vocab, seed=0).sample(B, 26).to(dev)
    half = B // 2

    def pg_step(i):
        # question_coding "ours" mix (question_coding_trainer.py:120-152): supervised half teacher-forced, rest sampled
        pg.zero_grad(set_to_none=True)
        sup = pg(questions[:half], gt_programs[:half], decoding_strategy="sampling")
        uns = pg(questions[half:], decoding_strategy="sampling")
        (sup["loss"].mean() + uns["loss"].mean()).backward()

    def joint_step(i):
        # modules/elbo.py:230-275 restricted to the hot path: sample programs, answer with the NMN, REINFORCE the
        # generator with the (detached) answer log-likelihood.  A random-init generator samples mostly invalid
        # programs, so the NMN runs the batch's ground-truth-style programs (what a trained generator emits).
        feats, programs, answers = resident[i % 2]
        pg.zero_grad(set_to_none=True)
        nmn.zero_grad(set_to_none=True)
        gen = pg(questions, decoding_strategy="sampling")
        out = nmn(feats, programs, answers)
        reward = (-out["loss"]).detach()
        objective = out["loss"].mean() + (gen["loss"] * (reward - reward.mean())).mean()
        objective.backward()

    res = {}
    for name, fn in (("pg", pg_step), ("joint", joint_step)):
        for i in range(3):
            fn(i)
        steps = max(3, min(args.steps, 10))
        ms = timed(fn, steps)
        world = int(os.environ.get("WORLD_SIZE", "1"))
        res[name] = {"value": world * B * steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps, "steps": steps}
    res["pg"]["workload"] = ("question_coding_ours.yml mix at batch %d: %d rows teacher-forced + %d rows sampled (26 steps), "
                             "questions <= 40 tokens, fwd+bwd (BASELINE.json configs[2])" % (B, half, B - half))
    res["joint"]["workload"] = ("joint step at batch %d: ProgramGenerator sampling fwd+bwd + NMN fwd+bwd on GT-style programs + "
                                "REINFORCE-weighted objective (BASELINE.json configs[3] restricted to the hot path)" % B)
    return res


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
