#!/usr/bin/env python
"""Benchmark of the hot path named by BASELINE.json: one joint_training iteration (joint_training_ours.yml) at batch 256
per GPU -- ProgramGenerator sampling -> QuestionReconstructor + NeuralModuleNetwork + ProgramPrior on the SAMPLED
programs -> REINFORCE / ELBO objective + the supervised terms -> backward -> gradient clamp -> Adam.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload joint|executor]

Workload "joint" (default; BASELINE.json configs[3], configs[4] at N = 8): synthetic JointTrainingDataset batches
(questions <= 40 tokens that determine their programs, 14x14x1024 features, answers, Bernoulli(0.5) supervision flags),
reference-shaped weights (He-normal module network, default-initialised reconstructor / prior, a program generator
pre-trained on the synthetic question -> program mapping: the config starts from checkpoints, and a random-init generator
samples programs the executor cannot run).  One step = probnmn_clevr_b200.joint.JointTrainingStep.step: what
JointTrainingTrainer.step does per batch (trainers/_trainer.py:172-196, joint_training_trainer.py:128-198); N > 1 adds the
NCCL gradient average before the clamp.
Workload "executor" (BASELINE.json configs[1]): NeuralModuleNetwork forward + backward on ground-truth-style programs of
up to 40 tokens; reported under "extra" in the default run.

Prints ONE JSON line (rank 0).  `value` is measured with inputs resident in HBM, `e2e` with pinned host inputs copied in
and the objective read back every step.  `roofline` is the persistent tcgen05 executor kernel (the dominant kernel) timed
with CUDA events around every launch; `cpu_baseline` is the CPU oracle port of the same step on a bounded sample.
`--impl reference` times that CPU path alone.
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

METRIC = "questions/sec (joint_training fwd+bwd, batch 256 per GPU)"
METRIC_EXECUTOR = "questions/sec (NMN executor fwd+bwd, batch 256 per GPU)"
UNIT = "questions/s"
PG_ASSET = os.path.join(ROOT, "probnmn_clevr_b200", "assets", "pg_synthetic_fp16.npz")
# configs/joint_training_ours.yml:6-25
JOINT = dict(alpha=100.0, beta=0.1, gamma=1.0, delta=0.99, objective="ours", lr=1e-6, weight_decay=0.0, clamp=5.0)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="joint", choices=["joint", "executor"])
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--length", type=int, default=40)
    ap.add_argument("--cpu-sample", type=int, default=32, help="rows of the workload the CPU baseline runs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary legs (executor-only, ProgramGenerator)")
    return ap.parse_args()


def executor_config(args):
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
        "impl": "reference", "metric": METRIC_EXECUTOR, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": executor_config(args),
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
def init_ours():
    """process group / device of this rank (one process per GPU)"""
    import torch.distributed as dist
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
    return {"world": world, "rank": rank, "local": local, "dev": dev}


def make_timed(ctx):
    import torch.distributed as dist
    world, dev = ctx["world"], ctx["dev"]

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
    return timed


def load_peaks():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        peaks = {}
    peak = peaks.get("bf16_tflops_sustained", 1590.0)
    src = "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1590 (of fallback)"
    return peak, src, peaks


def run_executor(args, ctx, extra=False):
    """BASELINE.json configs[1]: the module executor alone (stem + modules + classifier, forward + backward, no optimizer)
    on ground-truth-style programs.  The default run reports it under "extra"."""
    import torch.distributed as dist
    from probnmn_clevr_b200 import _lib as L
    from probnmn_clevr_b200.nmn import NeuralModuleNetwork
    from probnmn_clevr_b200.synthetic import ProgramSampler, make_answers, make_features, make_nmn_state_dict
    from probnmn_clevr_b200.vocabulary import Vocabulary

    world, rank, local, dev = ctx["world"], ctx["rank"], ctx["local"], ctx["dev"]

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

    timed = make_timed(ctx)

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
    peak, peak_src, _ = load_peaks()

    line = {
        "metric": METRIC_EXECUTOR, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 8), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f16", "data": "synthetic", "config": executor_config(args),
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
        "classifier_math": {"split": "pnmn_gemm_split: bf16 (hi, lo) split operands, three tcgen05 MMAs per k step, fp32 accumulate (no library GEMM)",
                            "tf32": "tf32 (cuBLAS/cuDNN, comparison only)", "ieee": "ieee fp32 (cuBLAS/cuDNN, comparison only)"}[model.classifier_math],
    }
    model._drop_precompiled()
    if extra:
        return line
    if not args.no_extras:
        line["pg"] = pg_leg(args, ctx, vocab, timed)
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
    return line


def traffic_per_launch():
    """dram__bytes_read.sum + dram__bytes_write.sum of exec_kernel per launch, from the committed ncu --set full capture
    of the same workload (profiles/<round>_traffic.json, written by scripts/summarize_profiles.py); None if absent."""
    try:
        files = sorted(f for f in os.listdir(os.path.join(ROOT, "profiles")) if f.endswith("_traffic.json"))
        return json.load(open(os.path.join(ROOT, "profiles", files[-1])))["exec_kernel_bytes_per_launch"]
    except Exception:
        return None


def pg_leg(args, ctx, vocab, timed):
    """BASELINE.json configs[2]: ProgramGenerator alone on the question_coding "ours" mix."""
    from probnmn_clevr_b200.seq2seq import ProgramGenerator
    from probnmn_clevr_b200.synthetic import ProgramSampler, make_questions, make_seq2seq_state_dict
    B, dev = args.batch, ctx["dev"]
    pg = ProgramGenerator(vocab)
    pg.load_state_dict(make_seq2seq_state_dict(vocab.get_vocab_size("questions"), vocab.get_vocab_size("programs"), seed=0))
    pg = pg.to(dev).train()
    questions = make_questions(B, vocab.get_vocab_size("questions"), seed=0, max_length=40).to(dev)
    gt_programs = ProgramSampler(vocab, seed=0).sample(B, 26).to(dev)
    half = B // 2

    # question_coding "ours" mix (question_coding_trainer.py:120-152): the supervised half is teacher-forced, the rest is
    # sampled.  The reference calls the model once per kind; here both kinds share ONE pass (Seq2SeqBase.forward_mixed: per
    # row the results of the separate calls), as in the joint-training step
    free = pg._max_decoding_steps
    width = max(gt_programs.shape[1], free - 1)
    targets = torch.zeros(B, width, dtype=torch.int64, device=dev)
    targets[:half, : gt_programs.shape[1]] = gt_programs[:half]
    teacher_rows = torch.zeros(B, dtype=torch.uint8, device=dev)
    teacher_rows[:half] = 1

    def pg_step(i):
        pg.zero_grad(set_to_none=True)
        out = pg.forward_mixed(questions, targets, teacher_rows, free_steps=free)
        (out["loss"][:half].mean() + out["loss"][half:].mean()).backward()

    for i in range(3):
        pg_step(i)
    steps = max(3, min(args.steps, 10))
    ms = timed(pg_step, steps)
    # the LSTM is bound by its chain of dependent steps, not by a roofline -- reported as achieved
    # FLOP/s (algorithmic forward FLOPs per sample: encoder 2 layers x (T_q + 1) x 2 x (256 + 256) x 1024, decoder
    # steps x [2 x 768 x 1024 + 2 x 2 x T_src x 256 + 2 x 256 x 44]; forward + backward = 3 x forward) and as time per
    # dependent step (encoder ticks + decoder steps, forward and backward)
    t_src, dec_steps = 41, free
    f_fwd = 2 * t_src * 2 * 512 * 1024 + dec_steps * (2 * 768 * 1024 + 4 * t_src * 256 + 2 * 256 * 44)
    chain = 2 * ((t_src + 1) + dec_steps)     # wavefront of the two encoder layers + decoder steps, both directions
    return {"value": ctx["world"] * B * steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps, "steps": steps,
            "achieved_tflops": 3.0 * f_fwd * B / (ms / steps * 1e-3) / 1e12,
            "dependent_steps_per_pass": chain, "us_per_dependent_step": 1e3 * (ms / steps) / chain,
            "bound": "latency (chain of dependent steps; SURVEY.md 8d: no roofline fraction target)",
            "workload": ("question_coding_ours.yml mix at batch %d: %d rows teacher-forced + %d rows sampled (26 steps) in one mixed pass, "
                         "questions <= 40 tokens, fwd+bwd (BASELINE.json configs[2])" % (B, half, B - half))}


def eval_leg(args, ctx, vocab, timed, pg, nmn):
    """Validation-shaped pass (evaluators/joint_training_evaluator.py:98-103 under _evaluator.py:67-115): eval mode, no
    gradients, the generator decodes greedily (teacher-forced on the ground-truth program, as the evaluator calls it) and the
    module network executes its predictions; metrics accumulate on the device and are read once at the end."""
    from probnmn_clevr_b200.synthetic import make_joint_batch
    B, dev = args.batch, ctx["dev"]
    batches = []
    for i in range(2):
        b = make_joint_batch(vocab, B, seed=900 + i)
        batches.append({k: b[k].to(dev) for k in ("question", "program", "image", "answer")})
    was_training = pg.training, nmn.training
    pg.eval(); nmn.eval()

    def eval_step(i):
        b = batches[i % 2]
        with torch.no_grad():
            g = pg(b["question"], b["program"], decoding_strategy="greedy")
            nmn(b["image"], g["predictions"], b["answer"])

    for i in range(3):
        eval_step(i)
    nmn.get_metrics(reset=True); pg.get_metrics(reset=True)
    steps = max(3, min(args.steps, 10))
    ms = timed(eval_step, steps)
    metrics = {**{k: float(v) for k, v in nmn.get_metrics(reset=True).items()},
               **{k: float(v) for k, v in pg.get_metrics(reset=True).items() if k in ("sequence_accuracy", "perplexity")}}
    pg.train(was_training[0]); nmn.train(was_training[1])
    return {"value": ctx["world"] * B * steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps, "steps": steps,
            "metrics": metrics,
            "workload": "joint_training evaluator iteration at batch %d: ProgramGenerator greedy (teacher-forced) -> "
                        "NeuralModuleNetwork on its predictions, eval mode, no_grad; inputs resident" % B}


# ---------------------------------------------------------------------------------------------------------
# the joint_training iteration (headline)
# ---------------------------------------------------------------------------------------------------------
def joint_config(args, world):
    return {
        "workload": "joint_training_ours.yml batch=256/GPU: ProgramGenerator sampling -> QuestionReconstructor + NeuralModuleNetwork "
                    "+ ProgramPrior on the sampled programs, REINFORCE/ELBO + supervised terms, backward, clamp, Adam "
                    "(BASELINE.json configs[3]; configs[4] at 8 GPUs)",
        "batch_per_gpu": args.batch, "questions": "<= 40 tokens, synthetic, determine their programs",
        "programs": "sampled by the ProgramGenerator (26 decoding steps); ground-truth programs (<= 26 tokens) on supervised rows",
        "features": [1024, 14, 14], "supervision": "Bernoulli(0.5) per row (SupervisionWeightedRandomSampler balances the two kinds)",
        "hyper": JOINT,
        "weights": "module network He-normal seed 0; reconstructor / prior default init; program generator pre-trained on the "
                   "synthetic question->program mapping (probnmn_clevr_b200/assets/pg_synthetic_fp16.npz, scripts/pretrain_pg.py)",
        "l2": "inputs larger than L2 (~103 MB of features per step, two alternating batches)",
        "parallelism": f"dp{world}",
        "step_pipelining": "the module network's backward pass, gradient average and clamp + Adam of step i are issued at the "
                           "start of step i+1, next to the generator's forward pass (JointTrainingStep(defer_nmn=True); same "
                           "arithmetic, flushed inside every timed region; PNMN_JOINT_DEFER_NMN=0 switches it off)",
    }


def load_pg_asset():
    import numpy as np
    if not os.path.exists(PG_ASSET):
        raise SystemExit(f"{PG_ASSET} is missing (scripts/pretrain_pg.py writes it)")
    z = np.load(PG_ASSET)
    return {k: torch.from_numpy(z[k].astype("float32")) for k in z.files}


def joint_state_dicts(vocab):
    from probnmn_clevr_b200.synthetic import make_nmn_state_dict, make_prior_state_dict, make_seq2seq_state_dict
    vq, vp = vocab.get_vocab_size("questions"), vocab.get_vocab_size("programs")
    return {"program_generator": load_pg_asset(), "question_reconstructor": make_seq2seq_state_dict(vp, vq, seed=1),
            "nmn": make_nmn_state_dict(vocab, 0), "program_prior": make_prior_state_dict(vp, seed=0)}


def run_joint(args, ctx):
    import numpy as np
    import torch.distributed as dist
    from probnmn_clevr_b200 import _lib as L
    from probnmn_clevr_b200.feed import DevicePrefetcher
    from probnmn_clevr_b200.joint import JointTrainingStep, split_batch
    from probnmn_clevr_b200.nmn import NeuralModuleNetwork
    from probnmn_clevr_b200.program_prior import ProgramPrior
    from probnmn_clevr_b200.seq2seq import ProgramGenerator, QuestionReconstructor
    from probnmn_clevr_b200.synthetic import make_joint_batch
    from probnmn_clevr_b200.vocabulary import Vocabulary

    world, rank, local, dev = ctx["world"], ctx["rank"], ctx["local"], ctx["dev"]
    torch.manual_seed(0)
    vocab = Vocabulary.clevr()
    sds = joint_state_dicts(vocab)
    models = {}
    for name, cls in (("program_generator", ProgramGenerator), ("question_reconstructor", QuestionReconstructor),
                      ("nmn", NeuralModuleNetwork), ("program_prior", ProgramPrior)):
        m = cls(vocab)
        m.load_state_dict(sds[name])
        models[name] = m.to(dev).train()
    if world > 1 and os.environ.get("PNMN_NO_GRAD_OVERLAP") is None:
        models["nmn"].enable_gradient_overlap()
    # defer_nmn: the module network's backward pass + update of step i is issued with step i + 1 (same arithmetic;
    # JointTrainingStep.flush() issues what is pending -- the timed regions below start and end with a flush, so each holds
    # exactly K forward and K backward passes of every model).  PNMN_JOINT_DEFER_NMN=0: every step complete in itself
    js = JointTrainingStep(models["program_generator"], models["question_reconstructor"], models["nmn"],
                           models["program_prior"], concurrent=os.environ.get("PNMN_NO_STREAMS") is None,
                           defer_nmn=os.environ.get("PNMN_JOINT_DEFER_NMN", "1") != "0", **JOINT)
    nmn = models["nmn"]

    # two alternating batches per rank, split on the host the way an input pipeline would (joint.split_batch), pinned
    KEYS = (("unsup", "question"), ("unsup", "image"), ("unsup", "answer"), ("sup", "question"), ("sup", "program"))
    host = []
    for i in range(2):
        parts = split_batch(make_joint_batch(vocab, args.batch, seed=100 * rank + i))
        host.append([parts[a][b].contiguous().pin_memory() for a, b in KEYS])
    as_parts = lambda ts: {"unsup": {"question": ts[0], "image": ts[1], "answer": ts[2]}, "sup": {"question": ts[3], "program": ts[4]}}
    resident = [as_parts([t.to(dev) for t in h]) for h in host]
    timed = make_timed(ctx)

    run_ahead = [torch.cuda.Event() for _ in range(3)]
    issued = {"n": 0, "wait_s": 0.0}

    def throttle():   # the host runs at most two steps ahead of the device (see run_executor)
        k = issued["n"]
        if k >= 2:
            t0 = time.perf_counter()
            run_ahead[(k - 2) % 3].synchronize()
            issued["wait_s"] += time.perf_counter() - t0

    plan_flops = {"n": 0, "flops": 0, "valid": 0, "convs": 0, "tokens": 0, "unsup": 0}

    def account():
        st = nmn.last_plan_stats
        plan_flops["n"] += 1
        plan_flops["flops"] += st[8]      # forward convolution FLOPs of the plan the forward pass just ran from
        plan_flops["valid"] += st[0]; plan_flops["convs"] += st[1]; plan_flops["tokens"] += st[2]

    def account_backward(st):
        # (called when a backward pass of the module network is launched -- with a deferred backward pass that is one step
        # later than its forward pass, and with the split compile the forward plan does not know the backward FLOPs)
        plan_flops["flops"] += st[10]

    nmn.backward_stats_hook = account_backward

    def resident_step(i):
        throttle()
        out = js.step(resident[i % 2])
        account()
        run_ahead[issued["n"] % 3].record()
        issued["n"] += 1
        return out

    # ---- warm-up, then a parity check of what is about to be timed: the module network's per-row losses on the programs the
    # generator just sampled, against the CPU oracle, on a 16-row slice
    W = max(args.warmup, 3)
    for i in range(W - 1):
        resident_step(i)
    # (weights as the last warm-up step sees them: its optimizer update moves every weight by ~lr along sign(gradient), which
    # shifts the logits by ~1e-2 -- the oracle must run on the weights the forward pass used)
    js.flush()   # (a deferred update of the module network belongs to the weights "as the last warm-up step sees them")
    sd_now = {k: v.detach().cpu().clone() for k, v in nmn.state_dict().items()} if rank == 0 else None
    resident_step(W - 1)
    torch.cuda.synchronize()
    parity = None
    if rank == 0:
        from oracle import nmn_oracle
        last = js.elbo.last_outputs
        rows = min(16, last["nmn"]["loss"].shape[0])
        b = resident[(W - 1) % 2]["unsup"]
        torch.set_num_threads(os.cpu_count())
        with torch.no_grad():
            ref = nmn_oracle.nmn_forward(sd_now, vocab, b["image"][:rows].cpu(), last["program_generator"]["predictions"][:rows].cpu(),
                                         b["answer"][:rows].cpu())
        mine = last["nmn"]["loss"][:rows].detach().cpu()
        err = float((mine - ref["loss"]).abs().max())
        same_valid = bool(torch.equal(last["nmn"]["predictions"][:rows].cpu() == 28, ref["valid"] == 0))
        parity = {"rows": rows, "max_abs_nmn_loss_err_vs_oracle": err, "validity_identical": same_valid,
                  "valid_in_slice": int(ref["valid"].sum())}
        assert same_valid and err < 5e-3, f"bench parity check failed: {parity}"

    import gc
    gc.collect()
    gc.freeze()
    lib = L.lib()
    host_ms = (ctypes.c_double * 4)()
    lib.pnmn_debug_host_times(host_ms)
    sampler = ClockSampler(local)
    sampler.start()
    lib.pnmn_launch_count(1)
    issued["wait_s"] = 0.0
    for k in plan_flops:
        plan_flops[k] = 0
    # (JointTrainingStep.defer_nmn: a step may leave the module network's backward pass + update to be issued with the next
    # step; what is pending is issued BEFORE the timer starts and again before it stops, so the region holds exactly K of them)
    js.flush()
    ms = timed(resident_step, args.steps, js.flush)
    host_issue_ms = timed.host_issue_ms - issued["wait_s"] * 1e3
    own_launches = int(lib.pnmn_launch_count(1))
    sampler.stop_flag = True
    lib.pnmn_debug_host_times(host_ms)
    host_ms_per_step = {"plan_create": host_ms[0] / args.steps, "issue_total": host_issue_ms / args.steps,
                        "run_ahead_wait": issued["wait_s"] * 1e3 / args.steps}
    per_step = {k: plan_flops[k] / max(plan_flops["n"], 1) for k in ("valid", "convs", "tokens")}
    unsup_rows = sum(h[0].shape[0] for h in host) / 2

    # ---- end to end: pinned host inputs -> device on a side stream, objective read back every step (one step later)
    feed = DevicePrefetcher(dev, depth=3)
    n_slots = args.steps + W + 8
    loss_host = torch.zeros(n_slots, dtype=torch.float32).pin_memory()
    loss_events = [torch.cuda.Event() for _ in range(n_slots)]
    loss_values = []
    total = {"n": 0}

    def read_loss(i):
        loss_events[i].synchronize()
        loss_values.append(float(loss_host[i]))

    def e2e_step(i):
        if i == 0:
            feed.submit(0, host[0])
        ts = feed.get(i)
        if i + 1 < total["n"]:
            feed.submit(i + 1, host[(i + 1) % 2])
        out = js.step(as_parts(ts))
        loss_host[i:i + 1].copy_(out["objective"].reshape(1), non_blocking=True)
        loss_events[i].record()
        if i >= 1:
            read_loss(i - 1)

    total["n"] = W + args.steps
    for i in range(W):
        e2e_step(i)
    loss_values.clear()
    js.flush()
    ms_e2e = timed(lambda j: e2e_step(W + j), args.steps, lambda: (js.flush(), read_loss(total["n"] - 1)))
    assert len(loss_values) == args.steps + 1 and all(v == v for v in loss_values), "every step's objective must have been read back"
    h2d = sum(t.numel() * t.element_size() for t in host[0])
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(3):
        resident[0]["unsup"]["image"].copy_(host[0][1], non_blocking=True)
    ev1.record()
    torch.cuda.synchronize()
    h2d_ms = ev0.elapsed_time(ev1) / 3

    # ---- end to end with a device-resident feature cache (feed.ImageFeatureCache): the features
    # of the split's images live in HBM as fp16 operand values (the reference keeps them in host RAM, readers.py:85-89); a step
    # sends only its tokens, answers and IMAGE INDICES over PCIe and gathers the feature batch on the device
    e2e_cached = None
    if not args.no_extras:
        from probnmn_clevr_b200.feed import ImageFeatureCache
        images = torch.cat([h[1] for h in host])                      # the "split": the images of both batches
        cache = ImageFeatureCache(images, dev, split="train")
        offs = [0, host[0][1].shape[0]]
        host_idx = []
        for i in range(2):
            n = host[i][1].shape[0]
            # (questions share images in CLEVR: 10 questions per image; here every question has its own, the worst case)
            idx = (torch.arange(n, dtype=torch.int64) + offs[i]).pin_memory()
            host_idx.append([host[i][0], idx, host[i][2], host[i][3], host[i][4]])
        feed2 = DevicePrefetcher(dev, depth=3)

        def cached_step(i):
            if i == 0:
                feed2.submit(0, host_idx[0])
            ts = feed2.get(i)
            if i + 1 < total["n"]:
                feed2.submit(i + 1, host_idx[(i + 1) % 2])
            parts = {"unsup": {"question": ts[0], "image": cache.gather(ts[1]), "answer": ts[2]},
                     "sup": {"question": ts[3], "program": ts[4]}}
            out = js.step(parts)
            loss_host[i:i + 1].copy_(out["objective"].reshape(1), non_blocking=True)
            loss_events[i].record()
            if i >= 1:
                read_loss(i - 1)

        for i in range(W):
            cached_step(i)
        loss_values.clear()
        js.flush()
        ms_c = timed(lambda j: cached_step(W + j), args.steps, lambda: (js.flush(), read_loss(total["n"] - 1)))
        e2e_cached = {"value": world * args.batch * args.steps / (ms_c * 1e-3), "unit": UNIT, "ms_per_step": ms_c / args.steps,
                      "h2d_bytes_per_step": sum(t.numel() * t.element_size() for t in host_idx[0]), "d2h_bytes_per_step": 4,
                      "cache_bytes": cache.features.numel() * 2,
                      "pipeline": "as e2e, but the image features come from feed.ImageFeatureCache (device-resident fp16 operand "
                                  "values keyed by image index, filled once before the timed region); results identical"}
        del cache

    # ---- per-kernel device time of a few profiled steps (CUDA events around the library's launches, same stream)
    prof_steps = min(args.steps, 5)
    kinds = ["elementwise", "exec_kernel", "conv_tc<1,3>", "wgrad_tc", "bias_grad", "pack_weights", "nchw_to_planes", "other"]
    for k in plan_flops:
        plan_flops[k] = 0
    js.flush()
    lib.pnmn_profile_enable(1)
    for i in range(prof_steps):
        resident_step(i)
    js.flush()   # (the last step's deferred backward pass belongs to the profiled steps)
    pms, pln = (ctypes.c_double * 8)(), (ctypes.c_int64 * 8)()
    lib.pnmn_profile_read(pms, pln)
    lib.pnmn_profile_enable(0)
    kernel_ms = {k: pms[i] / prof_steps for i, k in enumerate(kinds) if pms[i] > 0}
    conv_flops = plan_flops["flops"] / prof_steps
    conv_ms = pms[1] / prof_steps
    achieved = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    peak, peak_src, _ = load_peaks()

    # ---- phases of the step, each timed alone on the device (diagnostics for DESIGN.md; not part of `value`)
    phases = joint_phases(js, resident, timed) if not args.no_extras else None

    value = world * args.batch * args.steps / (ms * 1e-3)
    e2e_value = world * args.batch * args.steps / (ms_e2e * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16", "data": "synthetic", "config": joint_config(args, world), "clocks": sampler.summary(),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps, "h2d_copy_alone_ms": h2d_ms,
                "pipeline": "input pipeline splits the batch on the host (joint.split_batch: features of supervised rows are not "
                            "needed by the step) into pinned buffers; the copy of step i+1 runs on a side stream during step i "
                            "(feed.DevicePrefetcher); every step's objective is copied to pinned host memory and read one step later"},
        "gpu_launches": own_launches,
        "roofline": {
            "bound": "tensor", "kernel": "exec_kernel (persistent tcgen05 kind::f16 shift-GEMM executor: forward + dgrad launches)",
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "peak_source": peak_src,
            "note": "fp16 operands, fp32 accumulation in TMEM; peak is the bf16 cuBLAS figure; the kernel's time includes its "
                    "CUDA-core tasks and dependency waits; M is padded 196 -> 256 rows per sample (ceiling 0.766 of peak)",
            "flops_per_step": conv_flops, "kernel_ms_per_step": conv_ms, "launches_per_step": pln[1] / prof_steps,
            "m_padding_ceiling": 196.0 / 256.0, "traffic": traffic_per_launch(),
            # in the joint step the executor's persistent CTAs leave some SMs to the LSTM passes that run next to it
            # (JointTrainingStep.reserved_sms): `frac` stays against the whole chip's peak
            "sms_left_to_other_streams": js.reserved_sms,
        },
        "kernel_ms_per_step": kernel_ms, "host_ms_per_step": host_ms_per_step,
        "programs_per_step": {"unsupervised_rows": unsup_rows, "executable": per_step["valid"],
                              "conv3x3_instances": per_step["convs"], "module_tokens": per_step["tokens"]},
        "optimizer_launches_per_step": js.optimizer.launches_last_step,
        "parity_check": parity, "phases_ms": phases,
        "classifier_math": "pnmn_gemm_split: bf16 (hi, lo) split operands, three tcgen05 MMAs per k step, fp32 accumulate (no library GEMM)",
    }
    if not args.no_extras:
        sub = argparse.Namespace(**vars(args))
        sub.steps = max(5, min(args.steps, 30))
        ex = run_executor(sub, ctx, extra=True)
        line["extra"] = {"e2e_feature_cache": e2e_cached, "executor": {k: ex[k] for k in ("metric", "value", "ms_per_step", "steps", "e2e", "roofline", "kernel_ms_per_step", "host_ms_per_step", "plan")},
                         "pg": pg_leg(args, ctx, vocab, timed),
                         "eval": eval_leg(args, ctx, vocab, timed, models["program_generator"], nmn)}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_joint_baseline(args, best_of=2)
    return line


def joint_phases(js, resident, timed):
    """device time of the step's parts run one after the other on ONE stream (no overlap), forward + backward each"""
    pg, qr, nmn, prior = js.program_generator, js.question_reconstructor, js.nmn, js.program_prior
    b = resident[0]
    un, su = b["unsup"], b["sup"]
    with torch.no_grad():
        programs = pg(un["question"], decoding_strategy="sampling")["predictions"]
    res = {}

    def run(name, fn, steps=10):
        for _ in range(2):
            fn(0)
        res[name] = timed(fn, steps) / steps

    def pg_unsup(i):
        pg.zero_grad(); pg(un["question"], decoding_strategy="sampling")["loss"].mean().backward()

    def qr_unsup(i):
        qr.zero_grad(); qr(programs, un["question"], decoding_strategy="sampling")["loss"].mean().backward()

    def pg_sup(i):
        pg.zero_grad(); pg(su["question"], su["program"], decoding_strategy="sampling")["loss"].mean().backward()

    def qr_sup(i):
        qr.zero_grad(); qr(su["program"], su["question"], decoding_strategy="sampling")["loss"].mean().backward()

    def nmn_fb(i):
        nmn.zero_grad(); nmn(un["image"], programs, un["answer"])["loss"].mean().backward()

    def prior_f(i):
        with torch.no_grad():
            prior(programs)

    def opt(i):
        js.optimizer.step()

    run("program_generator_sampling_fwd_bwd", pg_unsup); run("question_reconstructor_unsup_fwd_bwd", qr_unsup)
    run("program_generator_supervised_fwd_bwd", pg_sup); run("question_reconstructor_supervised_fwd_bwd", qr_sup)
    run("nmn_fwd_bwd", nmn_fb); run("program_prior_fwd", prior_f)
    js.optimizer.zero_grad(); js.do_iteration(b)
    run("clamp_adam", opt)
    return res


# ---------------------------------------------------------------------------------------------------------
# CPU path of the joint step (oracle ports): cpu_baseline and --impl reference
# ---------------------------------------------------------------------------------------------------------
def cpu_joint_setup(args, rows):
    from oracle import joint_oracle
    from probnmn_clevr_b200.synthetic import make_joint_batch
    from probnmn_clevr_b200.vocabulary import Vocabulary
    vocab = Vocabulary.clevr()
    sds = {name: {k: v.clone().requires_grad_(name != "program_prior") for k, v in sd.items()}
           for name, sd in joint_state_dicts(vocab).items()}
    sds["program_prior"]["_output_layer.weight"] = sds["program_prior"]["_embedder.token_embedder_programs.weight"]
    batch = {k: v[:rows] for k, v in make_joint_batch(vocab, args.batch, seed=0).items()}
    opt = torch.optim.Adam(joint_oracle.trained_parameters(sds), lr=JOINT["lr"], weight_decay=JOINT["weight_decay"])
    state = joint_oracle.ElboState()
    gen = torch.Generator().manual_seed(0)

    def step():
        opt.zero_grad()
        out = joint_oracle.joint_iteration(sds, vocab, batch, state, alpha=JOINT["alpha"], beta=JOINT["beta"], gamma=JOINT["gamma"],
                                           delta=JOINT["delta"], objective=JOINT["objective"], generator=gen)
        joint_oracle.clamp_and_step(sds, opt, JOINT["clamp"])
        return out
    return step, batch


def cpu_joint_baseline(args, best_of=2):
    torch.set_num_threads(os.cpu_count())
    rows = args.cpu_sample
    step, batch = cpu_joint_setup(args, rows)
    step()
    best = 1e30
    for _ in range(best_of):
        t0 = time.perf_counter()
        out = step()
        best = min(best, time.perf_counter() - t0)
    return {"value": rows / best, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
            "sample": f"{rows} rows of the batch-{args.batch} workload ({int(batch['supervision'].sum())} supervised), one full joint "
                      f"iteration (fwd+bwd+clamp+Adam) on the CPU oracle ports, best of {best_of} after 1 warm-up; "
                      f"{int(out['rows']['nmn_valid'].sum())} of {len(out['rows']['nmn_valid'])} sampled programs executable"}


def run_reference_joint(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count())
    rows = args.cpu_sample
    step, batch = cpu_joint_setup(args, rows)
    warm = min(args.warmup, 1)
    for _ in range(warm):
        step()
    steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    value = rows * steps / dt
    sample = (f"{rows} rows of the batch-{args.batch} workload per step ({int(batch['supervision'].sum())} supervised), full joint "
              f"iteration on the CPU oracle ports (fwd+bwd+clamp+Adam), {steps} steps")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": joint_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_ours(args):
    import torch.distributed as dist
    ctx = init_ours()
    if args.workload == "joint":
        line = run_joint(args, ctx)
    else:
        line = run_executor(args, ctx)
    if ctx["rank"] == 0:
        print(json.dumps(line))
    if ctx["world"] > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference_joint(a) if a.workload == "joint" else run_reference(a)
    else:
        run_ours(a)
