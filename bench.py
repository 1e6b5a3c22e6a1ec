#!/usr/bin/env python
"""bench.py — G+D train-step throughput (sequences/s) of the B200-native Multi-StyleGAN hot path.

  python bench.py --gpus N --steps K --warmup W            # this package (CUDA, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on the host CPUs

Workload (BASELINE.json config 3): full generator + U-Net discriminator training iteration of
model_wrapper._gan_training — D step, lazy R1 (every 16th), G step, lazy path length (every 16th, half
batch), EMA — default 512-channel / 256x256 model, batch 8 per GPU (weak scaling), synthetic data,
random-init weights, no ADA / CutMix (epoch-0 behaviour).  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "train_step_sequences_per_sec"
UNIT = "sequences/s"


VERBOSE = bool(os.environ.get("MSG_BENCH_VERBOSE"))


def workload_config(args, world, reference=False):
    cfg = {"workload": "full G+D train step (BASELINE config 3): default 512ch/256x256 Multi-StyleGAN generator + "
                       "U-Net discriminator, NS-logistic loss, lazy R1 + path length every 16th iteration, EMA",
           "per_gpu_batch": args.batch, "global_batch": args.batch * world, "resolution": 256,
           "parallelism": "dp%d" % world, "ada": bool(args.ada),
           "cuda_graphs": bool(not args.no_graphs),
           "l2_policy": "working set per step (>= 10 GiB of activations) >> 126 MB L2; inputs rotate over a pool"}
    if reference:
        # the CPU arm runs a bounded sample of the same workload: one plain iteration at batch 1, scaled per sequence
        cfg.update(per_gpu_batch=1, global_batch=1, parallelism="cpu", cuda_graphs=False, ada=False, lazy=False,
                   l2_policy="n/a (host cores)")
    return cfg


def measure_mma_issue_rate(dev, sustained_s=1.5):
    """Tensor-core TF32 issue-rate roofline (csrc/mma_rate.cu): back-to-back tcgen05.mma.kind::tf32 128x256x8 on every SM
    with operands resident in shared memory.  Returns (burst, sustained) TFLOP/s: best of 10 launches, and launches back
    to back for `sustained_s` seconds so that the power-capped clock of a long run applies."""
    import torch
    from multi_stylegan_b200 import _C
    iters = 16384                                             # ~2 ms per launch
    flops = _C.tf32_mma_rate_probe(256, dev)
    torch.cuda.synchronize(dev)
    best = 1e30
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        flops = _C.tf32_mma_rate_probe(iters, dev)
        e1.record()
        torch.cuda.synchronize(dev)
        best = min(best, e0.elapsed_time(e1))
    reps = max(10, int(sustained_s * 1e3 / best))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        _C.tf32_mma_rate_probe(iters, dev)
    e1.record()
    torch.cuda.synchronize(dev)
    sust = e0.elapsed_time(e1) / reps
    return flops / (best * 1e-3) / 1e12, flops / (sust * 1e-3) / 1e12


def measure_tf32_peak(dev, sustained_s=2.0):
    """Dense TF32 tensor-core peak measured the way MEASURED_PEAKS.json measures bf16: torch.matmul (cuBLAS, fp32 operands
    with TF32 allowed) on 8192^3, best of 10 (burst) and back to back for `sustained_s` seconds (sustained)."""
    import torch
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a = torch.randn(n, n, device=dev)
        b = torch.randn(n, n, device=dev)
        c = torch.empty(n, n, device=dev)
        flops = 2.0 * n ** 3
        for _ in range(3):
            torch.matmul(a, b, out=c)
        torch.cuda.synchronize(dev)
        best = 1e30
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b, out=c)
            e1.record()
            torch.cuda.synchronize(dev)
            best = min(best, e0.elapsed_time(e1))
        reps = max(10, int(sustained_s * 1e3 / best))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize(dev)
        sust = e0.elapsed_time(e1) / reps
        del a, b, c
        return flops / (best * 1e-3) / 1e12, flops / (sust * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return p.get("hbm_gbs", 6650.0), p.get("bf16_tflops", 1590.0), p.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    FIELDS = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown," \
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown," \
             "clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle port) on the host cores
# ----------------------------------------------------------------------------------------------------
def cpu_reference_step_time(max_steps: int, warmup: int, budget_s: float):
    """One *plain* training iteration (D step + G step + EMA, no lazy regularisers) of the reference
    algorithm at batch 1 on the default model, all host threads.  Returns (seconds per step, steps timed)."""
    import torch
    from multi_stylegan_b200 import config
    import multi_stylegan_b200.multi_stylegan_generator as G_mod
    import multi_stylegan_b200.u_net_2d_discriminator as D_mod
    from oracle.train_step import OracleTrainer
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    G = G_mod.Generator(config.multi_style_gan_generator_config)        # parameter containers only (CPU)
    D = D_mod.Discriminator(config.u_net_2d_discriminator_config, no_rfp=True)
    hp = dict(config.generation_hyperparameters)
    hp["lazy_discriminator_regularization"] = 10 ** 9
    hp["lazy_generator_regularization"] = 10 ** 9
    tr = OracleTrainer(dict(G.state_dict()), dict(D.state_dict()), (2e-4, 2e-6), 6e-4, hp["betas"], hp,
                       dead_branch=True)
    del G, D
    times = []
    t_start = time.time()
    for i in range(warmup + max_steps):
        real = torch.rand(1, 2, 3, 256, 256)
        z = [[torch.randn(1, 512), torch.randn(1, 512)] for _ in range(2)]
        t0 = time.time()
        tr.step(real, z[0], z[1], None, None, None, None, 7, None)
        dt = time.time() - t0
        if i >= warmup:
            times.append(dt)
        if time.time() - t_start > budget_s and times:
            break
    return min(times), len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    sec, n = cpu_reference_step_time(max(1, args.steps), min(args.warmup, 1) if args.steps > 1 else 0, args.cpu_budget)
    value = 1.0 / sec
    cores = os.cpu_count() or 1
    sample = "oracle port of the reference trainer (oracle/train_step.py over oracle/model.py; the reference's own " \
             "modules are not available on this box), one plain iteration (D step + G step, dead second branch " \
             "evaluated as the reference does, no lazy regularisers) at batch 1, value = best of %d timed iteration(s) " \
             "(fastest, i.e. favourable to the CPU arm), torch %s CPU, %d threads" % (n, torch.__version__, cores)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": n, "warmup": args.warmup,
            "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, 1, reference=True), "impl": "reference",
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ----------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as tdist
    from multi_stylegan_b200 import _C, _lib, config, dist as mdist
    import multi_stylegan_b200.multi_stylegan_generator as G_mod
    import multi_stylegan_b200.u_net_2d_discriminator as D_mod
    from multi_stylegan_b200.adaptive_discriminator_augmentation import AdaptiveDiscriminatorAugmentation
    from multi_stylegan_b200.model_wrapper import ModelWrapper

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout carries the one JSON line, NCCL's banner goes to stderr
    _lib.lib()
    local_rank = mdist.init_from_env()
    world, rank = mdist.world_size(), mdist.rank()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    torch.manual_seed(0)
    B = args.batch
    G = G_mod.Generator(config.multi_style_gan_generator_config, compute_dead_branch=False).to(dev)
    D = D_mod.Discriminator(config.u_net_2d_discriminator_config, no_rfp=True).to(dev)
    mdist.broadcast_parameters([G, D])
    hp = dict(config.generation_hyperparameters)
    lazy_every = hp["lazy_generator_regularization"]
    graphs = not args.no_graphs
    # same Adam as train_multi_stylegan.py:53-57; fused=True only selects PyTorch's single-kernel implementation
    # (capturable=True keeps Adam's step counters on the device so that an iteration can be replayed as a CUDA graph)
    opt_g = torch.optim.Adam(G.get_parameters(lr_main=2e-4, lr_style=2e-6), betas=hp["betas"], fused=True, capturable=True)
    opt_d = torch.optim.Adam(D.parameters(), lr=6e-4, betas=hp["betas"], fused=True, capturable=True)

    def make_wrapper(ada: bool):
        Dw = AdaptiveDiscriminatorAugmentation(D) if ada else D
        if ada:
            Dw.p = 0.5                          # config 4: fixed p for timing (the controller is exercised by the tests)
            Dw.p_step = 0.0
        w = ModelWrapper(G, Dw, opt_g, opt_d, hyperparameters=hp, device=dev, cuda_graphs=graphs)
        w._d_params = lambda: list(D.parameters())
        return w
    mw = make_wrapper(bool(args.ada))
    torch.manual_seed(1234 + rank)
    pool = [torch.rand(B, 2, 3, 256, 256, device=dev) for _ in range(4)]
    host_pool = [p.cpu().pin_memory() for p in pool]
    state = {"graphs": graphs}

    def barrier():
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    def agree(flag: bool) -> bool:
        """True on every rank iff `flag` is true on every rank (a fallback decision must not desynchronise collectives)."""
        if world == 1:
            return flag
        t = torch.tensor([1 if flag else 0], device=dev)
        tdist.all_reduce(t, op=tdist.ReduceOp.MIN)
        return bool(t.item())

    # warm-up: W plain iterations + iterations with both lazy regularisers (their kernels and shapes)
    # (with CUDA graphs: the first lazy / plain iteration runs eagerly, the second one of each kind is captured)
    def warm_up(w, setup, plain):
        for i in range(setup + plain):
            is_lazy = i < setup and i % 2 == 0
            w.iteration = lazy_every - 1 if is_lazy else 0  # train_step increments first: iteration 16 runs R1 + PL
            out = w.train_step(pool[i % len(pool)])
            if VERBOSE:
                torch.cuda.synchronize()
                print("[bench rank %d] warm-up iteration %d done (replays so far: %d)" % (rank, i, w.graph_replays),
                      file=sys.stderr, flush=True)
            if is_lazy:
                assert "loss_path_length_regularization" in out and "loss_discriminator_regularization" in out, \
                    "warm-up did not exercise the lazy regularisers"

    def warm_up_checked(w, plain):
        """Graph mode degrades in steps that every rank takes together: NCCL all-reduces captured inside the iteration's
        graph -> graph segments with eager collectives in between -> eager issue."""
        while True:
            ok = True
            try:
                warm_up(w, 4 if state["graphs"] else 1, plain)
            except Exception as exc:                     # a capture that fails on this box must not cost the measurement
                if not state["graphs"]:
                    raise
                print("[bench rank %d] CUDA-graph capture failed (%s: %s)" % (rank, type(exc).__name__, str(exc)[:300]),
                      file=sys.stderr, flush=True)
                if VERBOSE:
                    import traceback
                    traceback.print_exc(file=sys.stderr)
                ok = False
            if not state["graphs"] or agree(ok):
                return
            w._graphs.clear()
            torch.cuda.synchronize()
            if world > 1 and os.environ.get("MSG_B200_NCCL_IN_GRAPH", "0") == "1":
                print("[bench rank %d] retrying with graph segments around eager collectives" % rank, file=sys.stderr, flush=True)
                os.environ["MSG_B200_NCCL_IN_GRAPH"] = "0"
                continue
            print("[bench rank %d] continuing with eager issue on all ranks" % rank, file=sys.stderr, flush=True)
            state["graphs"] = False
            args.no_graphs = True
            w.cuda_graphs = False
    warm_up_checked(mw, max(args.warmup, 3))
    graphs = state["graphs"]
    barrier()
    if graphs:
        assert mw.graph_replays >= 2 + max(args.warmup, 3), "CUDA graphs requested but the iterations ran eagerly"

    def timed(w, steps, e2e: bool, profile: bool = False, sample_clocks: bool = False):
        # profile=True: the same K iterations issued eagerly with a CUDA-event pair around every conv launch (roofline)
        w.cuda_graphs = graphs and not profile
        if graphs and profile:
            # graph capture emptied the caching allocator: refill the eager pool (one lazy + one plain iteration) untimed
            for it0 in (lazy_every - 1, 0):
                w.iteration = it0
                w.train_step(pool[0])
        w.iteration = 0
        launches0 = _C.launch_count() + w.graph_launches
        _C.profile_enable(profile)
        sampler = ClockSampler(local_rank)
        if sample_clocks and rank == 0:
            sampler.start()
        barrier()
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        t0 = time.time()
        marks[0].record()
        last = None
        torch.cuda.nvtx.range_push("timed_e2e" if e2e else "timed")            # main-thread kernels (forward passes)
        nvtx_id = torch.cuda.nvtx.range_start("step_e2e" if e2e else "step")    # process-wide: autograd thread too
        for i in range(steps):
            if e2e:
                real = host_pool[i % len(host_pool)].to(dev, non_blocking=True)
                out = w.train_step(real)
                last = {k: float(v) for k, v in out.items()}           # device -> host read of the step's losses
            else:
                w.train_step(pool[i % len(pool)])
            marks[i + 1].record()
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_end(nvtx_id)
        torch.cuda.nvtx.range_pop()
        barrier()
        wall = time.time() - t0
        ms = marks[0].elapsed_time(marks[-1])
        per_step = [marks[i].elapsed_time(marks[i + 1]) for i in range(steps)]
        clocks = sampler.stop() if (sample_clocks and rank == 0) else None
        t = torch.tensor([ms], device=dev)
        if world > 1:
            tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
        if VERBOSE:
            print("[bench rank %d] pass e2e=%s profile=%s: %.1f ms/step" % (rank, e2e, profile, ms / steps),
                  file=sys.stderr, flush=True)
        prof = _C.profile_summary() if profile else None
        _C.profile_enable(False)
        w.cuda_graphs = graphs
        return {"ms": float(t.item()), "launches": _C.launch_count() + w.graph_launches - launches0, "clocks": clocks,
                "prof": prof, "last": last, "wall": wall, "per_step": per_step}

    main = timed(mw, args.steps, False, profile=not graphs, sample_clocks=True)
    e2e = timed(mw, args.steps, True)
    pr = timed(mw, args.steps, False, profile=True) if graphs else main
    ms, launches, clocks, prof, ms_prof = main["ms"], main["launches"], main["clocks"], pr["prof"], pr["ms"]
    ms_e2e, last_losses = e2e["ms"], e2e["last"]

    # ---- config 5: EMA-generator sampling throughput (no_grad, single-style z, fresh noise; get_gan_samples.py:40-42) ----
    ema_line = None
    if not args.no_extras:
        try:
            ema_line = {"unit": UNIT, "points": []}
            g_ema = mw.generator_ema
            with torch.no_grad():
                for sb in (8, 16, 32):
                    z = torch.randn(sb, 512, device=dev)
                    for _ in range(3):
                        g_ema(z)
                    barrier()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    reps = 8
                    e0.record()
                    for _ in range(reps):
                        g_ema(torch.randn(sb, 512, device=dev))
                    e1.record()
                    torch.cuda.synchronize()
                    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
                    if world > 1:
                        tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
                    ema_line["points"].append({"per_gpu_batch": sb, "global_batch": sb * world,
                                               "value": world * sb * reps / (float(t.item()) / 1e3)})
            ema_line["note"] = "EMA generator forward under no_grad (eager issue), sequences/s over all ranks"
        except Exception as exc:
            ema_line = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:200])}
            if world > 1:
                raise

    tf32_burst = tf32_sust = mma_burst = mma_sust = None
    if not args.no_extras:
        try:
            tf32_burst, tf32_sust = measure_tf32_peak(dev)
            mma_burst, mma_sust = measure_mma_issue_rate(dev)
        except Exception as exc:
            print("[bench rank %d] TF32 peak measurement failed: %s" % (rank, str(exc)[:200]), file=sys.stderr, flush=True)
    # ---- config 4: the same step with ADA (p = 0.5) wrapped around the discriminator ------------------------------------
    ada_line = None
    if not args.ada and not args.no_extras:
        try:
            # the captured iterations of the plain wrapper own large private memory pools: release them first (an
            # allocation that has to free cached blocks while a capture is open invalidates the capture)
            import gc
            g_ema = None
            if VERBOSE:
                print("[bench rank %d] memory with the plain wrapper's graphs: %.1f GiB allocated, %.1f GiB reserved" %
                      (rank, torch.cuda.memory_allocated() / 2 ** 30, torch.cuda.memory_reserved() / 2 ** 30),
                      file=sys.stderr, flush=True)
            if os.environ.get("MSG_BENCH_KEEP_GRAPHS") != "1":
                mw.reset_cuda_graphs()
                gc.collect()
                torch.cuda.empty_cache()
            if VERBOSE:
                print("[bench rank %d] memory before the ADA variant: %.1f GiB allocated, %.1f GiB reserved" %
                      (rank, torch.cuda.memory_allocated() / 2 ** 30, torch.cuda.memory_reserved() / 2 ** 30),
                      file=sys.stderr, flush=True)
            mwa = make_wrapper(True)
            state["graphs"] = graphs
            warm_up_checked(mwa, 3)
            if state["graphs"] == graphs:
                ra = timed(mwa, args.steps, False)
                ada_line = {"p": 0.5, "ms_per_step": ra["ms"] / args.steps, "value": world * B * args.steps / (ra["ms"] / 1e3),
                            "unit": UNIT, "cuda_graphs": bool(graphs), "steps": args.steps,
                            "relative_to_plain": (ra["ms"] / args.steps) / (ms / args.steps),
                            "note": "BASELINE config 4: same train step, D wrapped in AdaptiveDiscriminatorAugmentation, "
                                    "fixed p = 0.5, 3 augmentation pipelines per iteration (real, fake, fake)"}
            else:
                ada_line = {"error": "graph capture failed for the ADA variant"}
            del mwa
        except Exception as exc:
            ada_line = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:200])}
            if world > 1:
                raise

    if world > 1:
        tdist.barrier()
        tdist.destroy_process_group()
    if rank != 0:
        return
    if args.profile_out and prof:
        with open(args.profile_out, "w") as f:
            for e in sorted(prof, key=lambda e: -e["ms_total"]):
                e = dict(e, avg_ms=e["ms_total"] / e["launches"],
                         tflops=e["flops_per_launch"] * e["launches"] / (e["ms_total"] * 1e-3) / 1e12)
                f.write(json.dumps(e) + "\n")
    seqs = world * B * args.steps
    value = seqs / (ms / 1e3)
    lazy_idx = [i for i in range(args.steps) if (i + 1) % lazy_every == 0]
    plain_ms = [t for i, t in enumerate(main["per_step"]) if i not in lazy_idx]
    lazy_ms = [main["per_step"][i] for i in lazy_idx]
    cfg = workload_config(args, world)
    cfg["cuda_graphs"] = bool(graphs)
    if world > 1:
        cfg["gradient_all_reduce"] = ("NCCL, captured inside the iteration's CUDA graph" if graphs and mw._collectives_in_graph()
                                      else "NCCL, issued eagerly between graph segments" if graphs else "NCCL, eager")
    cfg["lazy_r1_and_pl_steps_in_timed_region"] = len(lazy_idx)
    hbm, bf16_burst, bf16_sust, src = peaks()
    if mma_sust is not None:
        tf32_peak, peak_source = max(mma_sust, mma_burst), "measured in this run: tcgen05.mma.kind::tf32 128x256x8 issue rate on all SMs, operands " \
            "in shared memory (csrc/mma_rate.cu), launches back to back for 1.5 s: sustained %.0f TFLOP/s (best of 10 burst " \
            "%.0f).  The method of MEASURED_PEAKS.json applied to TF32 through cuBLAS (torch.matmul 8192^3, TF32 allowed) " \
            "gives %.0f sustained / %.0f burst TFLOP/s on this box — slower than this kernel, so it cannot serve as a peak; " \
            "half of MEASURED_PEAKS' bf16 sustained figure is %.0f" % (mma_sust, mma_burst, tf32_sust, tf32_burst, bf16_sust / 2.0)
    else:
        tf32_peak, peak_source = bf16_sust / 2.0, "1/2 of the %s bf16 sustained rate in MEASURED_PEAKS.json (TF32 not measured: --no-extras)" % src
    roof = None
    traffic_table = {}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")      # dram bytes per launch from the committed ncu captures
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic_table = json.load(f)
    if prof:
        top = max(prof, key=lambda e: e["ms_total"])
        achieved = top["flops_per_launch"] * top["launches"] / (top["ms_total"] * 1e-3) / 1e12
        conv_ms = sum(e["ms_total"] for e in prof)
        conv_flops = sum(e["flops_per_launch"] * e["launches"] for e in prof)
        kname = "tc_%s (tcgen05 kind::tf32) taps=%d K=%d N=%d pixels=%d" % (
            top["kind"], top["taps"], top["k_channels"], top["n_channels"], top["pixels"])
        roof = {"bound": "tensor", "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s",
                "frac": achieved / tf32_peak, "traffic": traffic_table.get(kname),
                "traffic_source": "profiles/traffic.json (dram bytes per launch from the committed ncu --set full capture of this launch)",
                "kernel": kname, "peak_source": peak_source,
                "algorithmic_flops_per_launch": top["flops_per_launch"],
                "frac_of_burst_peak": achieved / mma_burst if mma_burst else None,
                "peak_cublas_tf32_sustained": tf32_sust, "peak_cublas_tf32_burst": tf32_burst,
                "peak_half_bf16_sustained": bf16_sust / 2.0,
                "launches": top["launches"], "avg_ms": top["ms_total"] / top["launches"],
                "share_of_step": top["ms_total"] / ms_prof,
                "measured_in": ("an eager pass of the same %d iterations (%.1f ms/step) with a CUDA-event pair around each conv "
                                "launch; the timed region itself replays CUDA graphs" % (args.steps, ms_prof / args.steps))
                if graphs else "the timed region",
                "all_tcgen05_conv_kernels": {"ms": conv_ms, "share_of_step": conv_ms / ms_prof,
                                             "tflops": conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms else None}}
        # the same launches split by roofline regime: algorithmic bytes of a launch = its two activation-sized tensors at
        # the GEMM's pixel count, 4 * pixels * (K + N) (filters are negligible, halo / tap re-reads come out of L2); a shape
        # whose flops / byte is below the machine balance cannot reach the tensor peak whatever the kernel does
        balance = tf32_peak * 1e12 / (hbm * 1e9)
        regimes = {"tensor_bound": [0.0, 0.0, 0.0], "hbm_bound": [0.0, 0.0, 0.0]}
        for e in prof:
            nbytes = 4.0 * e["pixels"] * (e["k_channels"] + e["n_channels"])
            r = regimes["tensor_bound" if e["flops_per_launch"] / nbytes >= balance else "hbm_bound"]
            r[0] += e["ms_total"]
            r[1] += e["flops_per_launch"] * e["launches"]
            r[2] += nbytes * e["launches"]
        roof["all_tcgen05_conv_kernels"]["by_regime"] = {
            "machine_balance_flop_per_byte": balance,
            "tensor_bound": {"ms": regimes["tensor_bound"][0], "share_of_step": regimes["tensor_bound"][0] / ms_prof,
                             "tflops": regimes["tensor_bound"][1] / (regimes["tensor_bound"][0] * 1e-3) / 1e12
                             if regimes["tensor_bound"][0] else None},
            "hbm_bound": {"ms": regimes["hbm_bound"][0], "share_of_step": regimes["hbm_bound"][0] / ms_prof,
                          "gbps": regimes["hbm_bound"][2] / (regimes["hbm_bound"][0] * 1e-3) / 1e9
                          if regimes["hbm_bound"][0] else None, "hbm_peak_gbps": hbm,
                          "note": "bytes = the GEMM's two activation tensors only; residual operands and second outputs of "
                                  "the epilogues are not counted, so this is a lower bound on the bandwidth these launches run at"}}
    # secondary roofline (north star: upfirdn2d against HBM): the generator's 256^2 blur, CUDA events, inputs rotate over
    # 3 x 1 GiB (> L2); algorithmic bytes = 4 * (N_in + N_out)
    roof_hbm = None
    if world == 1 and not args.no_extras:
        try:
            k4 = torch.tensor([1., 3., 3., 1.], device=dev)
            k2d = (k4[None] * k4[:, None]) / 16
            xs = [torch.randn(B, 256, 256, 512, device=dev) for _ in range(3)]
            for i in range(3):
                _C.upfirdn2d(xs[i], k2d, 1, 1, 1, 1, 2, 1, 2, 1)
            torch.cuda.synchronize()
            groups = []
            for _ in range(5):                                   # median of 5 groups of 12 launches
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(12):
                    _C.upfirdn2d(xs[i % 3], k2d, 1, 1, 1, 1, 2, 1, 2, 1)
                e1.record()
                torch.cuda.synchronize()
                groups.append(e0.elapsed_time(e1) / 12)
            blur_ms = sorted(groups)[len(groups) // 2]
            nbytes = 2 * xs[0].numel() * 4
            roof_hbm = {"bound": "hbm", "kernel": "fir_cl_blur_kernel: upfirdn2d 4x4 pad (2,1) on [%d,512,256,256] fp32 channels-last" % B,
                        "achieved": nbytes / (blur_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                        "frac": nbytes / (blur_ms * 1e-3) / 1e9 / hbm, "traffic": traffic_table.get("fir_cl_blur 256"),
                        "avg_ms": blur_ms, "algorithmic_bytes_per_launch": nbytes}
            del xs
        except Exception as exc:                                         # never lose the headline line to the extra
            roof_hbm = {"error": str(exc)[:200]}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "tf32 (fp32 storage, fp32 accumulate)", "data": "synthetic", "config": cfg,
            "e2e": {"value": seqs / (ms_e2e / 1e3), "unit": UNIT,
                    "h2d_bytes_per_step": B * 2 * 3 * 256 * 256 * 4, "d2h_bytes_per_step": 4 * len(last_losses or {})},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "roofline_hbm": roof_hbm,
            "conv_engine": _C.conv2d_last_engine()}
    if plain_ms:
        t_plain = sum(plain_ms) / len(plain_ms)
        line["plain_iteration_ms"] = t_plain
        if lazy_ms:
            t_lazy = sum(lazy_ms) / len(lazy_ms)
            line["lazy_iteration_ms"] = t_lazy
            am = ((lazy_every - 1) * t_plain + t_lazy) / lazy_every
            line["amortised_16"] = {"ms_per_step": am, "value": world * B / (am / 1e3), "unit": UNIT,
                                    "note": "15 plain + 1 lazy (R1 + path length) iteration, from this rank's per-step CUDA events; "
                                            "equals `value` when --steps is a multiple of 16"}
    if ada_line is not None:
        line["ada"] = ada_line
    if ema_line is not None:
        line["ema_sampling"] = ema_line
    if world == 1 and not args.no_cpu_baseline:
        sec, n = cpu_reference_step_time(1, 0, args.cpu_budget)
        line["cpu_baseline"] = {"value": 1.0 / sec, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                                "sample": "oracle port, one plain iteration (D step + G step, no lazy regularisers, dead "
                                          "branch evaluated like the reference) at batch 1, %d run(s)" % n}
    emit(line)


_REAL_STDOUT = None


def _capture_stdout():
    """Route everything libraries print on fd 1 (NCCL's version banner, ...) to stderr; emit() writes the one JSON line
    to the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="per-GPU batch")
    ap.add_argument("--ada", action="store_true", help="wrap D in adaptive discriminator augmentation (config 4)")
    ap.add_argument("--no-graphs", action="store_true", help="issue every iteration eagerly instead of replaying CUDA graphs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the ADA / EMA-sampling / TF32-peak / blur sub-measurements")
    ap.add_argument("--profile-out", default=None, help="write the per-shape tcgen05 conv kernel timings (JSON lines)")
    ap.add_argument("--cpu-budget", type=float, default=150.0, help="seconds of CPU work allowed for the reference arm")
    args = ap.parse_args()
    _capture_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
