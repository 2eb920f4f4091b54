#!/usr/bin/env python
"""Benchmark of the MAE loss-and-evaluation hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Default workload:  C3 = BASELINE.json configs[2], "CIFAR-10 32x32x3, discretized logistic mixture (10 comps) decoder,
latent 64, batch 64, full MAE loss fwd+bwd" - the loss-head configuration with a real HBM roofline (configs[1],
OMNIGLOT, is 1.6 MB / 10 MFLOP per step and purely launch-latency-bound; it is reported under "workloads" together with
C4 = IW-NLL eval and C5 = MPD microbench at every N).  One step = reparam+KL, MPD, DMOL likelihood, objective assembly,
forward AND backward, on synthetic encoder/decoder outputs (the conv stacks stay in PyTorch and are not part of the
measured path).  Prints ONE JSON line on rank 0.

Timing rules followed: at least --warmup (and >= 3) warm-up steps, extended to ~0.4 s so that clocks have ramped; a ring
of input slabs larger than 3x L2 is cycled so every step reads from HBM; CUDA events on the launching stream; max over
ranks; the timed block of K steps is repeated (5x when K < 200) and the MEDIAN block is reported; nvidia-smi clocks are
sampled during the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

ETA, GAMMA, FREE_BITS = 1.0, 1.0, 0.0
L2_BYTES = 126e6
BATCH, LATENT = 64, 64
# identical in both arms (the driver compares the strings)
WORKLOAD = ("C3: CIFAR-10 32x32x3 head, DMOL nmix=10, latent 64, batch 64 per GPU, MPD on (eta=gamma=1), "
            "full MAE loss fwd+bwd on synthetic encoder/decoder outputs")
METRIC = "mae_loss_head_fwd_bwd_samples_per_s"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_bf16_tflops():
    """sustained dense bf16 tensor throughput of this pool's B200s (MEASURED_PEAKS.json; fallback of the profiling guide)"""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["bf16_tflops_sustained"])
    except Exception:
        return 1400.0


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the benched kernel (dmol_bwd_kernel<10,1024,5,true>, 160
    threads) from the committed ncu --set full capture; None if no capture of THAT kernel is committed."""
    path = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
    try:
        with open(path) as f:
            k = json.load(f)["dmol_fwd_bwd"]
        return k["dram_bytes_read"] + k["dram_bytes_write"]
    except Exception:
        return None


# ----------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples nvidia-smi SM clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------- synthetic inputs
def c3_host_inputs(slab: int, batch=BATCH, latent=LATENT):
    """One slab of C3 inputs on the host (pinned): encoder outputs, noise, decoder output, images."""
    from mae_b200 import synth
    seed = synth.SEED + 1000 * slab
    mu, lv = synth.posterior_params(batch, (latent,), seed=seed, clamp=True)
    enc = torch.cat([mu, lv], dim=1)                       # [B, 2D]: mu/logvar are chunk views of it
    eps = synth.noise(batch, 1, (latent,), seed=seed)
    gz = synth.noise(batch, 1, (latent,), seed=seed + 7) * 0.01   # d loss / d z as the decoder's backward would deliver it
    head = synth.color_head(batch, 1, 10, 32, 32, seed=seed)
    return {k: v.contiguous().pin_memory() if torch.cuda.is_available() else v
            for k, v in dict(enc=enc, eps=eps, gz=gz, raw=head["raw"], x=head["x"]).items()}


def c3_bytes(batch=BATCH, ns=1, nmix=10, hw=1024):
    fwd = 4 * (10 * nmix * hw * batch * ns + 3 * hw * batch + batch * ns)
    bwd = 4 * (2 * 10 * nmix * hw * batch * ns + 3 * hw * batch + batch * ns)
    return fwd, bwd


# -------------------------------------------------------------------------------------------------- our arm
class C3Step:
    """One slab's fwd+bwd of the colour loss head, captured into a CUDA graph."""

    def __init__(self, host, dev, dist_ctx=None):
        from mae_b200 import heads
        self.heads = heads
        self.dist = dist_ctx
        self.enc = host["enc"].to(dev)
        # mu / logvar are strided chunk views of the encoder output (encoder_cores/resnet.py:43), made leaves here
        # because the encoder itself is not part of the measured path
        self.mu = self.enc[:, :self.enc.size(1) // 2].detach().requires_grad_()
        self.lv = self.enc[:, self.enc.size(1) // 2:].detach().requires_grad_()
        self.one = torch.ones((), device=dev)
        self.eps = host["eps"].to(dev)
        self.gz = host["gz"].to(dev)
        self.raw = host["raw"].to(dev).requires_grad_()
        self.x = host["x"].to(dev)
        self.graph = None
        self.loss = None
        self.stats = None
        self.launches = 0

    def eager(self):
        self.mu.grad = None
        self.lv.grad = None
        self.raw.grad = None
        mu, lv = self.mu, self.lv
        if self.dist is None:
            loss, recon, kl, stats, z = self.heads.loss_head_color(mu, lv, self.eps, self.raw, self.x, 10, ETA, GAMMA, FREE_BITS,
                                                                   defer_join=True)
        else:
            loss, recon, kl, stats, z = self.dist.loss_head_color(mu, lv, self.eps, self.raw, self.x, 10, ETA, GAMMA, FREE_BITS)
        # the decoder is not part of the measured path: its backward is represented by the gradient it would hand to z
        torch.autograd.backward([loss, z], [self.one, self.gz])
        self.stats = stats
        return loss

    def capture(self):
        from mae_b200 import ops
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                self.eager()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        ops.reset_launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self.eager()
        self.launches = ops.launch_count()

    def run(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self.loss = self.eager()
        return self.loss


def time_kernel(fn, n_iter, stream, reps=1):
    """median over ``reps`` of the average device time of fn() over n_iter back-to-back launches (CUDA events on the
    launching stream)"""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0.record(stream)
        for i in range(n_iter):
            fn(i)
        e1.record(stream)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / n_iter * 1e-3)
    return statistics.median(ts)


def _dbg(msg):
    if os.environ.get("MAE_BENCH_DEBUG"):
        print("[bench rank %s %.2f] %s" % (os.environ.get("RANK", "0"), time.perf_counter(), msg), file=sys.stderr, flush=True)


def parity_across_ranks(dist_ctx, hosts, dev, rank, world):
    """N > 1: the sharded step of slab 0 must reproduce the single-GPU head on the concatenated global batch: global loss
    to 2e-6 relative, this rank's own-row gradients to 1e-5 of the tensor scale.  Every rank evaluates the reference."""
    import torch.distributed as dist
    from mae_b200 import heads, ops
    h = hosts[0]
    loc = {k: v.to(dev) for k, v in h.items()}
    glob = {}
    for k, v in loc.items():
        out = torch.empty((world * v.size(0),) + tuple(v.shape[1:]), device=dev, dtype=v.dtype)
        dist.all_gather_into_tensor(out, v.contiguous())
        glob[k] = out
    half = loc["enc"].size(1) // 2

    def leaves(t):
        return t[:, :half].detach().requires_grad_(), t[:, half:].detach().requires_grad_()

    mu, lv = leaves(loc["enc"])
    raw = loc["raw"].detach().requires_grad_()
    loss, recon, kl, stats, z = dist_ctx.loss_head_color(mu, lv, loc["eps"], raw, loc["x"], 10, ETA, GAMMA, FREE_BITS)
    torch.autograd.backward([loss, z], [torch.ones((), device=dev), loc["gz"]])
    g_loss = dist_ctx.global_loss(loss, stats)
    mu_a, lv_a = leaves(glob["enc"])
    raw_a = glob["raw"].detach().requires_grad_()
    ref = heads.loss_head_color(mu_a, lv_a, glob["eps"], raw_a, glob["x"], 10, ETA, GAMMA, FREE_BITS)
    torch.autograd.backward([ref[0], ref[4]], [torch.ones((), device=dev), glob["gz"]])
    ops.join_side(dev)
    torch.cuda.synchronize()
    b = loc["enc"].size(0)
    sl = slice(rank * b, (rank + 1) * b)

    def rel(a, r):
        return float((a.double() - r.double()).abs().max()) / max(float(r.double().abs().max()), 1e-30)

    errs = {"loss": rel(g_loss, ref[0].detach()), "dmu": rel(mu.grad, mu_a.grad[sl]), "dlogvar": rel(lv.grad, lv_a.grad[sl]),
            "draw": rel(raw.grad, raw_a.grad[sl])}
    ok = errs["loss"] <= 2e-6 and errs["dmu"] <= 1e-5 and errs["dlogvar"] <= 1e-5 and errs["draw"] <= 1e-5
    t = torch.tensor([1.0 if ok else 0.0] + [errs[k] for k in ("loss", "dmu", "dlogvar", "draw")], device=dev, dtype=torch.float64)
    lo = t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist_ctx.check()
    return bool(lo[0] > 0.5), {"loss": float(t[1]), "dmu": float(t[2]), "dlogvar": float(t[3]), "draw": float(t[4]),
                               "global_loss": float(g_loss), "single_gpu_loss": float(ref[0].detach())}


def bench_ours(args):
    import torch.distributed as dist
    from mae_b200 import _lib, ops
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.load()
    dist_ctx = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        from mae_b200.distributed import ShardedHead
        dist_ctx = ShardedHead(defer_join=True, check_every=0)
    hbm_gbs, peak_src = measured_peaks()
    batch, latent = BATCH, LATENT
    fwd_b, bwd_b = c3_bytes(batch)
    step_bytes = fwd_b + bwd_b
    ring = max(4, int(L2_BYTES * 3 / bwd_b) + 1)                     # > 3x L2 of distinct data actually moved per step
    hosts = [c3_host_inputs(i + 17 * rank, batch, latent) for i in range(min(ring, 4))]
    parity_n = None
    if world > 1:
        parity_n = parity_across_ranks(dist_ctx, hosts, dev, rank, world)
        _dbg("parity across ranks: %s" % (parity_n,))
    # d loss == 1 is created by this script (torch.autograd.backward([loss, z], [one, gz])): the conditional rescale launch
    # of the fused DMOL node is skipped (ops.set_unit_loss_grad)
    ops.set_unit_loss_grad(True)
    steps_obj = []
    use_graph = not args.no_graph
    _dbg("init done, ring=%d" % ring)
    for i in range(ring):
        st = C3Step(hosts[i % len(hosts)], dev, dist_ctx)
        if use_graph:
            st.capture()
            _dbg("captured %d" % i)
        steps_obj.append(st)
    launches_per_step = steps_obj[0].launches if use_graph else None
    if launches_per_step is None:
        ops.reset_launch_count()
        steps_obj[0].run()
        launches_per_step = ops.launch_count()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stream = torch.cuda.current_stream()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # warm-up: at least W (>= 3) steps, extended to ~0.4 s of replay so that clocks have ramped and nvidia-smi (100 ms
    # period) sees the load.  Multi-rank: a FIXED count (every rank must issue the same number of collectives).
    n_warm = 0
    t_w = time.perf_counter()
    fixed = max(args.warmup, 3000) if world > 1 else None
    while (n_warm < fixed) if fixed is not None else (n_warm < max(args.warmup, 3) or time.perf_counter() - t_w < 0.4):
        for i in range(50):
            steps_obj[(n_warm + i) % ring].run()
        n_warm += 50
        torch.cuda.synchronize()
    _dbg("warm-up done (%d steps)" % n_warm)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    repeats = 5 if args.steps < 200 else 3
    block_ms = []
    for _ in range(repeats):
        barrier()
        e0.record(stream)
        for i in range(args.steps):
            steps_obj[i % ring].run()
        e1.record(stream)
        barrier()
        t_blk = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([t_blk], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_blk = float(t)
        block_ms.append(t_blk)
    dt = statistics.median(block_ms) * 1e-3
    _dbg("timed region done")
    ops.join_side(dev)
    loss_val = float(steps_obj[(args.steps - 1) % ring].loss.detach())

    # ---- end to end through the public API with HOST buffers: H2D of the step's inputs, step, D2H of the loss
    e2e_steps = max(3, min(args.steps, 50))
    h = hosts[0]
    h2d = sum(v.numel() * 4 for v in h.values())
    host_loss = torch.empty((), dtype=torch.float32).pin_memory()
    # two device slabs alternate: the H2D copy of step i+1 (copy stream) overlaps with the kernels of step i
    copy_stream = torch.cuda.Stream()
    pipe = steps_obj[:2] if len(steps_obj) >= 2 else steps_obj[:1]
    done_ev = [None] * len(pipe)
    e2e_i = [0]

    def e2e_once():
        k = e2e_i[0] % len(pipe)
        e2e_i[0] += 1
        st = pipe[k]
        hh = hosts[k % len(hosts)]
        if done_ev[k] is not None:
            copy_stream.wait_event(done_ev[k])          # the slab's previous step has consumed its inputs
        with torch.cuda.stream(copy_stream), torch.no_grad():
            st.enc.copy_(hh["enc"], non_blocking=True)
            st.eps.copy_(hh["eps"], non_blocking=True)
            st.gz.copy_(hh["gz"], non_blocking=True)
            st.raw.copy_(hh["raw"], non_blocking=True)
            st.x.copy_(hh["x"], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        stream.wait_event(ev)
        st.run()
        host_loss.copy_(st.loss.detach(), non_blocking=True)
        done_ev[k] = torch.cuda.Event()
        done_ev[k].record(stream)

    for _ in range(3):
        e2e_once()
    _dbg("e2e warm done")
    e2e_ms = []
    for _ in range(3):
        barrier()
        e0.record(stream)
        for _ in range(e2e_steps):
            e2e_once()
        e1.record(stream)
        barrier()
        t_blk = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([t_blk], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_blk = float(t)
        e2e_ms.append(t_blk)
    dt_e2e = statistics.median(e2e_ms) * 1e-3
    _dbg("e2e done")
    clocks = sampler.stop() if rank == 0 else None

    # ---- roofline of the dominant kernel, timed alone on the launching stream: the fused DMOL forward+backward (one pass:
    #      reads raw + x, writes err + d raw = 53.2 MB moved; the work it replaces is SURVEY.md §8d's fwd 27.0 MB + bwd 53.2 MB).
    #      Inputs AND outputs cycle through rings of > 3x L2 each, so reads come from and writes go to HBM.
    lib = _lib.load()
    raws = [s_.raw.detach() for s_ in steps_obj]
    xs = [s_.x for s_ in steps_obj]
    g = torch.full((batch, 1), 1.0 / batch, device=dev)
    draw = [torch.empty_like(raws[0]) for _ in range(ring)]
    err = torch.empty(batch, 1, device=dev)
    err_part = torch.empty(batch, max(int(lib.mae_dmol_err_chunks(batch, 1, 10, 1024)), 1), device=dev)
    ws = torch.zeros(max(int(lib.mae_dmol_workspace_bytes(batch, 1024)), 256), dtype=torch.uint8, device=dev)
    sp = stream.cuda_stream

    def k_fused(i):
        j = i % ring
        lib.mae_dmol_fwd_bwd(raws[j].data_ptr(), xs[j].data_ptr(), 1.0 / batch, batch, 1, 10, 1024, err_part.data_ptr(),
                             draw[j].data_ptr(), sp)

    def k_bwd(i):
        j = i % ring
        lib.mae_dmol_bwd(raws[j].data_ptr(), xs[j].data_ptr(), g.data_ptr(), 1, 1.0, batch, 1, 10, 1024, draw[j].data_ptr(), sp)

    def k_fwd(i):
        j = i % ring
        lib.mae_dmol_fwd(raws[j].data_ptr(), xs[j].data_ptr(), batch, 1, 10, 1024, err.data_ptr(), ws.data_ptr(), ws.numel(), sp)

    n_k = max(20, min(200, args.steps))
    _dbg("kernel timing start")
    t_fused = time_kernel(k_fused, n_k, stream, 5)
    # the same kernel launched one at a time (a synchronize between launches): no head/tail overlap with its neighbour
    ei0, ei1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iso = []
    for i in range(20):
        torch.cuda.synchronize()
        ei0.record(stream)
        k_fused(i)
        ei1.record(stream)
        torch.cuda.synchronize()
        iso.append(ei0.elapsed_time(ei1) * 1e-3)
    t_iso = statistics.median(iso)
    t_bwd = time_kernel(k_bwd, n_k, stream, 3)
    t_fwd = time_kernel(k_fwd, n_k, stream, 3)
    del draw
    _dbg("kernel timing done")
    step_s = dt / args.steps
    roof = {"bound": "hbm", "kernel": "dmol_bwd_kernel<10,1024,5,ERR=true> (fused DMOL forward+backward, mae_dmol_fwd_bwd)",
            "achieved": bwd_b / t_fused / 1e9, "peak": hbm_gbs, "unit": "GB/s", "frac": bwd_b / t_fused / 1e9 / hbm_gbs,
            "traffic": ncu_traffic(), "peak_source": peak_src, "bytes_per_launch": bwd_b, "us_per_launch": t_fused * 1e6,
            "us_per_launch_isolated": t_iso * 1e6, "frac_isolated": bwd_b / t_iso / 1e9 / hbm_gbs,
            "algorithmic_bytes_replaced": step_bytes, "frac_of_algorithmic_bytes": step_bytes / t_fused / 1e9 / hbm_gbs,
            "note": "achieved/frac count the 53.2 MB the kernel MOVES per launch (reads raw + x, writes d raw + err partials); "
                    "the separate forward + backward it replaces are 27.0 + 53.2 = 80.2 MB (SURVEY 8d, frac_of_algorithmic_bytes); "
                    "back-to-back launches over input and output rings of %d slabs each (> 3x L2), median of 5 blocks; "
                    "us_per_launch_isolated = one launch between synchronizes" % ring,
            "dmol_bwd_alone": {"achieved": bwd_b / t_bwd / 1e9, "frac": bwd_b / t_bwd / 1e9 / hbm_gbs, "bytes_per_launch": bwd_b,
                               "us_per_launch": t_bwd * 1e6},
            "dmol_fwd_alone": {"achieved": fwd_b / t_fwd / 1e9, "frac": fwd_b / t_fwd / 1e9 / hbm_gbs, "bytes_per_launch": fwd_b,
                               "us_per_launch": t_fwd * 1e6},
            "step": {"bytes_algorithmic": step_bytes, "bytes_moved": bwd_b, "us": step_s * 1e6,
                     "frac_algorithmic": step_bytes / step_s / 1e9 / hbm_gbs, "frac_moved": bwd_b / step_s / 1e9 / hbm_gbs}}

    extra = {}
    peaks = None
    if not args.no_extra:
        if rank == 0:
            from mae_b200 import peaks as P
            peaks = P.measure(dev)
        barrier()
        extra.update(bench_c4(dev, hbm_gbs, world))                       # every rank: its shard of the evaluation
        extra.update(bench_n2(dev, hbm_gbs, world))
        extra.update(bench_n3(dev, world, rank))
        if peaks is not None and "c4_iw_nll_eval" in extra:
            extra["c4_iw_nll_eval"]["roofline"]["sfu_coroof"] = c4_sfu_coroof(peaks)
        ffma_peak = [peaks["ffma_tflops"] if peaks else 0.0]
        if world > 1:
            t = torch.tensor(ffma_peak, device=dev, dtype=torch.float64)
            dist.broadcast(t, 0)
            ffma_peak = [float(t)]
        extra.update(bench_c5(dev, world, rank, dist_ctx, ffma_peak[0], quick=args.quick))
        if world == 1:
            extra.update(bench_c1(dev, hbm_gbs))
            extra.update(bench_full_model(dev))
            extra.update(bench_n4(dev))
            extra["gpu_eager_baseline"] = gpu_eager_baseline(dev, value_ours=batch * args.steps / dt)
    out = None
    if rank == 0:
        cpu = cpu_baseline_c3(budget_s=12.0) if (not args.no_cpu and world == 1) else None
        value = world * batch * args.steps / dt
        out = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": n_warm, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "global_batch": world * batch, "parallelism": "dp%d (batch-sharded; MPD over the all-gathered batch)" % world,
                       "l2": "ring of %d input slabs (%.0f MB read, %.0f MB written per cycle) > 3x 126 MB L2" % (
                           ring, ring * (fwd_b) / 1e6, ring * (bwd_b - fwd_b) / 1e6),
                       "cuda_graph": use_graph,
                       "timing": "median of %d blocks of %d steps; warmup = max(--warmup=%d, 3, ~0.4 s of replay%s)" % (
                           repeats, args.steps, args.warmup, ", 3000 at N>1" if world > 1 else "")},
            "block_ms": block_ms,
            "e2e": {"value": world * batch * e2e_steps / dt_e2e, "unit": "samples/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 4, "ms_per_step": dt_e2e / e2e_steps * 1e3, "steps": e2e_steps,
                    "note": "median of 3 blocks; pinned host inputs -> H2D (copy stream, double-buffered) -> graph replay -> D2H of the loss"},
            "gpu_launches": launches_per_step * args.steps, "launches_per_step": launches_per_step,
            "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "loss": loss_val, "measured_peaks": peaks,
            "workloads": extra,
        }
        if parity_n is not None:
            out["parity_n"] = parity_n[0]
            out["parity_n_rel_err"] = parity_n[1]
    if world > 1:
        dist_ctx.check()
        dist.barrier()
        torch.cuda.synchronize()
    return out, steps_obj


# ------------------------------------------------------------------------------------- secondary workloads
def c4_sfu_coroof(peaks):
    """SFU co-roof of the DMOL forward: 17 SFU instructions per (pixel, component) on the common path (hot_math.cuh: 3 ex2
    of the scales, 6 ex2 of |hi| / |lo|, 3 ex2 + 1 rcp of the tanh coefficients, 1 rcp + 1 lg2 shared by the three channels,
    2 ex2 of the two log-sum-exp accumulations) -> 170 per pixel-sample + ~4 to finish the two log-sum-exps."""
    per_pix = 17 * 10 + 4
    per_img = per_pix * 512 * 1024
    rate = min(peaks["mufu_ex2_gops"], peaks["mufu_lg2_gops"], peaks["mufu_rcp_gops"]) * 1e9
    return {"sfu_ops_per_image": per_img, "sfu_peak_gops_measured": rate / 1e9, "images_per_s_at_sfu_peak": rate / per_img}


def bench_c4(dev, hbm_gbs, world=1):
    """C4 (IW-NLL eval) on every rank's own shard of synthetic images (no communication but the final reduction, §8e);
    throughput = images of all ranks / max time over ranks."""
    import torch.distributed as dist
    from mae_b200 import heads
    res = {}
    stream = torch.cuda.current_stream()
    # ---- C4: K=512 samples per image, 8 images per launch, ring of 3 slabs (3 x 1.68 GB) > L2; each slab's evaluation
    #      (DMOL forward + fused IW kernel) is captured in a CUDA graph so the figure is device throughput, not the
    #      Python launch rate
    k, nimg, d = 512, 8, 64
    slabs, graphs, outs = [], [], []
    g = torch.Generator(device=dev).manual_seed(524287)
    for i in range(3):
        raw = torch.randn(nimg * k, 100, 32, 32, device=dev, generator=g)
        x = (torch.randint(0, 256, (nimg, 3, 32, 32), device=dev, generator=g).float() / 127.5 - 1.0)
        mu = torch.randn(nimg, d, device=dev, generator=g)
        lv = torch.randn(nimg, d, device=dev, generator=g).clamp(-7, 7)
        z = mu[:, None] + torch.randn(nimg, k, d, device=dev, generator=g) * (0.5 * lv).exp()[:, None]
        slabs.append((raw, x, z, mu, lv))
    per_img = 4 * (k * 100 * 1024 + 3 * 1024 + k * (2 * d + 3) + 2 * d)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for sl in slabs:
            heads.iw_eval_color(*sl, 10)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    for sl in slabs:
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            outs.append(heads.iw_eval_color(*sl, 10))
        graphs.append(gr)

    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    t = time_kernel(lambda i: graphs[i % 3].replay(), 30, stream, 3)
    if world > 1:
        tt = torch.tensor([t], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t = float(tt)
    res["c4_iw_nll_eval"] = {"metric": "iw_nll_eval_images_per_s", "value": world * nimg / t, "unit": "images/s", "n_gpus": world,
                             "scaling": "weak (8 images x K=512 per launch per GPU; test images are sharded, one final all-reduce)",
                             "config": "CIFAR shapes, K=512, %d images/launch, ring of 3 slabs (%.1f GB) > L2, CUDA graph" % (nimg, 3 * nimg * k * 100 * 1024 * 4 / 1e9),
                             "us_per_image": t / nimg * 1e6,
                             "roofline": {"bound": "hbm", "achieved_per_gpu": nimg * per_img / t / 1e9, "peak": hbm_gbs,
                                          "frac": nimg * per_img / t / 1e9 / hbm_gbs, "bytes_per_image": per_img}}
    del graphs, outs
    del slabs
    torch.cuda.empty_cache()
    return res


def bench_n2(dev, hbm_gbs, world=1):
    """N2: the colour decoder's last 1x1 convolution fused with the DMOL likelihood (csrc/dmol_head.cu: TMA ->
    3xTF32 tcgen05.mma with the operand and the accumulators in TMEM -> DMOL epilogue).  Evaluation shape of C4: 8 images x
    K = 512 samples of the PixelCNN++ CIFAR decoder (hidden 96 channels, image_decoders/pixelcnnpp.py:28-34), inputs are
    the INPUT of the last ELU -> Conv2dWeightNorm(96, 100, 1) pair.  Beside it: the same work unfused (ATen ELU + cuDNN/cuBLAS
    1x1 convolution writing the raw tensor, then this repo's DMOL forward kernel reading it), and the training shape of C3
    (batch 64, kernel also emits d err / d raw)."""
    import torch.distributed as dist
    import torch.nn.functional as F
    from mae_b200 import ops
    res = {}
    stream = torch.cuda.current_stream()
    k, nimg, c, nmix = 512, 8, 96, 10
    g = torch.Generator(device=dev).manual_seed(8191)
    w = torch.randn(100, c, device=dev, generator=g) * (0.5 / c ** 0.5)
    bias = torch.randn(100, device=dev, generator=g) * 0.1
    slabs = []
    for i in range(3):       # 3 x 1.61 GB of hidden > L2
        hid = torch.randn(nimg * k, c, 32, 32, device=dev, generator=g)
        x = (torch.randint(0, 256, (nimg, 3, 32, 32), device=dev, generator=g).float() / 127.5 - 1.0)
        slabs.append((hid, x))

    def fused(i):
        hid, x = slabs[i % 3]
        return ops.dmol_head_recon_error(hid, w, bias, x, k, nmix, True)

    w4 = w.view(100, c, 1, 1)

    def unfused(i):
        hid, x = slabs[i % 3]
        return ops.dmol_recon_error(F.conv2d(F.elu(hid), w4, bias), x, k, nmix)

    with torch.no_grad():
        a, b_ = fused(0), unfused(0)
        torch.cuda.synchronize()
        agree = float((a - b_).abs().max() / b_.abs().max())
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        t_f = time_kernel(fused, 12, stream, 3)
        t_u = time_kernel(unfused, 6, stream, 3)
        if world > 1:
            tt = torch.tensor([t_f, t_u], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t_f, t_u = float(tt[0]), float(tt[1])
    px = nimg * k * 1024
    bytes_fused = 4 * (px * c + nimg * 3 * 1024 + nimg * k)
    bytes_unfused = 4 * (px * (2 * c) + px * (c + 100) + px * 100 + nimg * 3 * 1024 + nimg * k)   # ELU r+w, conv r+w, DMOL r
    mma_flop = px * 2.0 * 112 * (c + 8) * 3                                   # executed: 3 tf32 products, N padded to 112
    res["n2_fused_head_eval"] = {
        "metric": "fused_conv1x1_dmol_eval_images_per_s", "value": world * nimg / t_f, "unit": "images/s", "n_gpus": world,
        "config": "PixelCNN++ CIFAR head: hidden [8 x 512, 96, 32, 32] fp32 -> ELU -> 1x1 conv (100) -> DMOL nmix=10, ring of 3 slabs (4.8 GB) > L2",
        "us_per_image": t_f / nimg * 1e6, "ms_per_launch": t_f * 1e3,
        "unfused": {"what": "ATen ELU + cuDNN 1x1 convolution (raw tensor written) + this repo's DMOL forward kernel",
                    "images_per_s": world * nimg / t_u, "ms_per_launch": t_u * 1e3, "speedup_fused": t_u / t_f,
                    "hbm_bytes_per_image": bytes_unfused // nimg},
        "hbm_bytes_per_image": bytes_fused // nimg, "hbm_bytes_saved_per_image": (bytes_unfused - bytes_fused) // nimg,
        "fused_vs_unfused_rel_diff": agree,
        "roofline": {"bound": "issue / SFU (epilogue) over hbm", "achieved_per_gpu": bytes_fused / t_f / 1e9, "peak": hbm_gbs,
                     "frac": bytes_fused / t_f / 1e9 / hbm_gbs, "tf32_mma_tflops_executed": mma_flop / t_f / 1e12,
                     "note": "per 128-pixel tile: 38 tcgen05.mma (128 x 112 x 8, 3xTF32) = 2.1 k tensor-pipe cycles, ~13 k warp "
                             "instructions (9.2 k DMOL epilogue, 3.3 k ELU + split) = 3.3 k issue cycles at 4 IPC, 24 k + 12 k SFU "
                             "ops = 2.3 k cycles; profiles/README.md has the per-role cycle breakdown"}}
    del slabs
    torch.cuda.empty_cache()

    # training shape (C3): batch 64, one sample; the kernel also writes d err / d raw (26 MB)
    if world == 1:
        bt = 64
        hid = [torch.randn(bt, c, 32, 32, device=dev, generator=g) for _ in range(8)]      # 8 x 25 MB
        x = (torch.randint(0, 256, (bt, 3, 32, 32), device=dev, generator=g).float() / 127.5 - 1.0)
        lib = ops._lib_()
        err = torch.empty(bt, 1, device=dev)
        draws = [torch.empty(bt, 100, 32, 32, device=dev) for _ in range(8)]
        ws = torch.zeros(int(lib.mae_dmol_head_workspace_bytes(bt, 1024)) + 256, dtype=torch.uint8, device=dev)
        raws = [F.conv2d(F.elu(h), w4, bias) for h in hid]
        chunks = int(lib.mae_dmol_err_chunks(bt, 1, nmix, 1024))
        errp = torch.empty(bt, 1, chunks, device=dev)

        def fused_train(i):
            lib.mae_dmol_head(hid[i % 8].data_ptr(), w.data_ptr(), bias.data_ptr(), x.data_ptr(), bt, 1, nmix, c, 1024, 1, 1.0 / bt,
                              err.data_ptr(), draws[i % 8].data_ptr(), ws.data_ptr(), ws.numel(), stream.cuda_stream)

        def dmol_train(i):
            lib.mae_dmol_fwd_bwd(raws[i % 8].data_ptr(), x.data_ptr(), 1.0 / bt, bt, 1, nmix, 1024, errp.data_ptr(),
                                 draws[i % 8].data_ptr(), stream.cuda_stream)

        def conv_only(i):
            F.conv2d(F.elu(hid[i % 8]), w4, bias)

        with torch.no_grad():
            t_ft = time_kernel(fused_train, 50, stream, 5)
            t_dt = time_kernel(dmol_train, 50, stream, 5)
            t_cv = time_kernel(conv_only, 50, stream, 5)
        res["n2_fused_head_train_c3"] = {
            "config": "batch 64 x 32 x 32, hidden 96: last conv + DMOL forward + d err/d raw in one launch", "us_fused": t_ft * 1e6,
            "us_unfused": (t_dt + t_cv) * 1e6, "us_unfused_parts": {"elu_conv1x1": t_cv * 1e6, "dmol_fwd_bwd": t_dt * 1e6},
            "note": "us_unfused = ATen ELU + cuDNN 1x1 convolution (TF32 allowed) writing the raw tensor + this repo's one-pass DMOL "
                    "forward+backward on it.  At 512 tiles (3.5 per SM) the fused kernel's persistent pipeline has little depth, so "
                    "it is 2.6x the DMOL kernel alone but 1.5x faster than the pair it replaces; training still keeps the "
                    "raw-tensor path by default (ops.set_dmol_head('eval')) because the decoder stacks' own backward wants d raw "
                    "either way and the two-node loss head overlaps the regulariser with it; 'always' is the opt-in"}
    return res


def bench_full_model(dev):
    """The caller's view (experiments/color_image.py:153-160, binary_image.py:150-160): whole ``MAE.loss`` + backward of the two
    shipped configurations (encoder / decoder conv stacks in PyTorch, loss head on this repo's kernels), eagerly and as one
    CUDA graph (mae_b200.graph.capture_step).  Configurations come from tests/golden/models.pt (the reference's JSON files with
    one replica and dropout 0), weights from synth.fill_state; batch sizes are the training scripts' defaults."""
    import copy
    from mae_b200 import graph, ops, synth
    from mae_b200.modules import MAE
    res = {}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests", "golden", "models.pt")
    if not os.path.exists(path):
        return {"full_model_step": {"unavailable": "tests/golden/models.pt not found"}}
    fixtures = torch.load(path, weights_only=False)
    stream = torch.cuda.current_stream()
    for name, batch in (("binary_image_json", 100), ("color_image32x32_json", 64)):
        try:
            c = fixtures[name]
            model = synth.fill_state(MAE.from_params(copy.deepcopy(c["params"]))).to(dev).train()
            shape = (batch,) + tuple(c["x"].shape[1:])
            g = torch.Generator(device=dev).manual_seed(17)
            if shape[1] == 1:
                x = (torch.rand(shape, device=dev, generator=g) < 0.2).float()
            else:
                x = torch.randint(0, 256, shape, device=dev, generator=g).float() / 127.5 - 1.0
            args = dict(eta=1.0, gamma=1.0, free_bits=0.0)

            def eager(i):
                for p_ in model.parameters():
                    p_.grad = None
                model.loss(x, 1, **args)[0].backward()

            n0 = ops.launch_count()
            eager(0)
            n_ours = ops.launch_count() - n0
            t_eager = time_kernel(eager, 5, stream, 3)
            step = graph.capture_step(model, x, 1, **args)
            t_graph = time_kernel(lambda i: step(x), 10, stream, 3)
            res["full_model_step_" + name] = {
                "config": "%s, batch %d, nsamples 1, eta = gamma = 1: MAE.loss + backward (no optimizer)" % (name, batch),
                "ms_eager": t_eager * 1e3, "ms_cuda_graph": t_graph * 1e3, "speedup_graph": t_eager / t_graph,
                "samples_per_s_cuda_graph": batch / t_graph, "launches_of_ours_per_step": n_ours,
                "parameters": sum(p_.numel() for p_ in model.parameters())}
            del step, model
            torch.cuda.empty_cache()
        except Exception as e:      # never lose the headline line over a secondary workload
            res["full_model_step_" + name] = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}
    return res


def bench_n3(dev, world, rank):
    """N3: ``evaluate.calc_nll`` - the reference's test loop (experiments/color_image.py:252-277: one image per
    MAE.nll call) batched (8 images x K = 512 per call), sharded over ranks with one final all-reduce, the per-batch body
    replayed as a CUDA graph - through a REAL decoder: the ResNet colour configuration of tests/golden/models.pt (ResNet
    encoder / decoder conv stacks in PyTorch, last 1x1 convolution fused into the likelihood kernel).  A bounded sample of 64
    images per rank; the figure includes the final all-reduce and the host read of the result."""
    import copy
    import torch.distributed as dist
    from mae_b200 import evaluate, synth
    from mae_b200.modules import MAE
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests", "golden", "models.pt")
    if not os.path.exists(path):
        return {"n3_calc_nll": {"unavailable": "tests/golden/models.pt not found"}}
    try:
        c = torch.load(path, weights_only=False)["color_resnet"]
        model = synth.fill_state(MAE.from_params(copy.deepcopy(c["params"]))).to(dev)
        per_rank, k = 160, 512
        g = torch.Generator().manual_seed(23)
        data = torch.randint(0, 256, (per_rank * world, 3, 32, 32), generator=g).float() / 127.5 - 1.0
        head = None
        if world > 1:
            from mae_b200.distributed import ShardedHead
            head = ShardedHead(check_every=0, overlap=False)
        out = {}

        def timed(n_images, use_graph):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = evaluate.calc_nll(model, data[:n_images * world], k, images_per_step=8, sharded=head, cuda_graph=use_graph)
            torch.cuda.synchronize()
            t = time.perf_counter() - t0
            if world > 1:
                tt = torch.tensor([t], device=dev, dtype=torch.float64)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                t = float(tt)
            return t, r

        for label, use_graph in (("eager", False), ("cuda_graph", True)):
            timed(24, use_graph)                                     # warm-up
            t_small, _ = timed(32, use_graph)
            t_big, res = timed(per_rank, use_graph)
            # steady-state rate from the difference of two runs: the one-time graph capture and the final all-reduce /
            # host read are in both
            out[label] = {"images_per_s": (per_rank - 32) * world / max(t_big - t_small, 1e-9), "s": t_big, "bits_per_dim": res[2]}
        return {"n3_calc_nll": {
            "metric": "calc_nll_images_per_s", "value": out["cuda_graph"]["images_per_s"], "unit": "images/s", "n_gpus": world,
            "config": "ResNet colour MAE (tests/golden/models.pt color_resnet), K = 512, 8 images per MAE.nll call; steady-state rate from "
                      "two runs (32 and %d images per rank, wall clock incl. the final all-reduce and host read)" % per_rank,
            "eager_images_per_s": out["eager"]["images_per_s"],
            "graph_speedup": out["cuda_graph"]["images_per_s"] / out["eager"]["images_per_s"],
            "whole_call_s": {"eager": out["eager"]["s"], "cuda_graph_incl_capture": out["cuda_graph"]["s"]},
            "bits_per_dim": out["cuda_graph"]["bits_per_dim"]}}
    except Exception as e:
        return {"n3_calc_nll": {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}}


def bench_n4(dev):
    """N4: where the time of one ``MAE.nll`` call goes for the shipped CIFAR configuration (8 images x K = 512):
    encoder, prior-flow inverse on [B K, 16, 8, 8] (``prior_flow.backward``, gaussian_encoder.py:254: 4 AF2d-MADE blocks whose
    masked convolutions are re-normalised per call, networks/masked.py:102-106), decoder trunk (PixelCNN++), and this path's
    kernels (fused last convolution + DMOL, IW-NLL).  Shows which callers of the path dominate evaluation."""
    import copy
    from mae_b200 import ops, synth
    from mae_b200.modules import MAE
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests", "golden", "models.pt")
    if not os.path.exists(path):
        return {"n4_nll_breakdown": {"unavailable": "tests/golden/models.pt not found"}}
    try:
        c = torch.load(path, weights_only=False)["color_image32x32_json"]
        model = synth.fill_state(MAE.from_params(copy.deepcopy(c["params"]))).to(dev).eval()
        b, k = 2, 128            # 256 decoder evaluations per call (the PixelCNN++ trunk holds ~100 MB of activations per 64)
        g = torch.Generator(device=dev).manual_seed(29)
        x = torch.randint(0, 256, (b, 3, 32, 32), device=dev, generator=g).float() / 127.5 - 1.0
        enc, dec = model.encoder, model.decoder
        stream = torch.cuda.current_stream()
        with torch.no_grad():
            z, (mu, logvar, z_normal, logdet) = model.sample_from_posterior(x, nsamples=k)
            zf = z.view(b * k, *enc.prior_flow.input_size())
            zd = z.view(b, k, *dec.z_shape())
            trunk, last, act = dec.head_parts()
            xe = x.repeat_interleave(k, dim=0)
            hidden = trunk(xe, zd.view(b * k, *dec.z_shape()))
            w, bias = last.conv.weight().flatten(1).contiguous(), last.conv.bias
            t = {
                "encoder_and_sampling": time_kernel(lambda i: model.sample_from_posterior(x, nsamples=k), 3, stream, 3),
                "prior_flow_backward": time_kernel(lambda i: enc.prior_flow.backward(zf), 3, stream, 3),
                "decoder_trunk_pixelcnnpp": time_kernel(lambda i: trunk(xe, zd.view(b * k, *dec.z_shape())), 2, stream, 3),
                "fused_last_conv_dmol (ours)": time_kernel(lambda i: ops.dmol_head_recon_error(hidden, w, bias, x, k, dec.nmix, act), 5, stream, 3),
                "whole_nll_call": time_kernel(lambda i: model.nll(x, k), 2, stream, 3),
            }
        tot = t["whole_nll_call"]
        return {"n4_nll_breakdown": {
            "config": "color_image32x32.json, %d images x K = %d per MAE.nll call" % (b, k),
            "ms": {kk: v * 1e3 for kk, v in t.items()}, "share_of_call": {kk: v / tot for kk, v in t.items() if kk != "whole_nll_call"},
            "note": "the prior-flow inverse and the PixelCNN++ trunk are PyTorch conv stacks (callers of the path, SURVEY 8f N4 / N1); "
                    "the path's own kernels are the last entry"}}
    except Exception as e:
        return {"n4_nll_breakdown": {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}}


def bench_c1(dev, hbm_gbs):
    """C1/C2 (binary head, B = 100, D = 64): launch-latency-bound (1.57 MB / 10 MFLOP per step)."""
    from mae_b200 import heads, ops, synth
    res = {}
    stream = torch.cuda.current_stream()
    h = synth.binary_head(100, 1, 784)
    mu0, lv0 = synth.posterior_params(100, (64,))
    enc = torch.cat([mu0, lv0], 1).to(dev)
    mu_l, lv_l = enc[:, :64].detach().requires_grad_(), enc[:, 64:].detach().requires_grad_()
    one = torch.ones((), device=dev)
    eps = synth.noise(100, 1, (64,)).to(dev)
    probs = h["probs"].to(dev).requires_grad_()
    xb = h["x"].to(dev)

    def c1_step():
        mu_l.grad = None
        lv_l.grad = None
        probs.grad = None
        loss = heads.loss_head_binary(mu_l, lv_l, eps, probs, xb, ETA, GAMMA, FREE_BITS, defer_join=True)[0]
        loss.backward(gradient=one)
        return loss

    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            c1_step()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    ops.reset_launch_count()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        c1_step()
    n_launch = ops.launch_count()
    t = time_kernel(lambda i: graph.replay(), 200, stream, 5)
    c1_bytes = 4 * (100 * 784 + 100 * 784 + 100) + 4 * (2 * 100 * 784 + 100 * 784 + 100)
    res["c1_binary_head"] = {"metric": METRIC, "value": 100 / t, "unit": "samples/s",
                             "config": "MNIST/OMNIGLOT 28x28 head, B=100, latent 64, Bernoulli + MPD fwd+bwd, CUDA graph",
                             "us_per_step": t * 1e6, "launches_per_step": n_launch,
                             "roofline": {"bound": "launch-latency", "hbm_floor_us": c1_bytes / (hbm_gbs * 1e9) * 1e6,
                                          "frac": c1_bytes / t / 1e9 / hbm_gbs}}
    return res


C5_SHAPES = ((4096, 64), (16384, 64), (65536, 64), (4096, 256), (16384, 256), (65536, 256))


def bench_c5(dev, world, rank, dist_ctx, ffma_peak_tflops, quick=False):
    """C5: MPD regulariser microbench, forward + backward, 16*B^2*D algorithmic flop.  At N > 1 the batch is
    STRONG-scaled: every rank owns B/N rows, the packed (mu | logvar) rows are all-gathered, each rank evaluates its row
    block of the B x B tile pass (one fp64 all-reduce of sum (KL - m)^2) and the complete gradient of its own rows."""
    import torch.distributed as dist
    from mae_b200 import ops, synth
    res = {}
    stream = torch.cuda.current_stream()
    head = None
    if world > 1:
        from mae_b200.distributed import ShardedHead
        head = ShardedHead(check_every=0, overlap=False)     # results are consumed by torch ops on the main stream
    shapes = C5_SHAPES[:2] + C5_SHAPES[3:4] if quick else C5_SHAPES
    for b5, d5 in shapes:
        if b5 % world:
            continue
        b_loc = b5 // world
        mu_g, lv_g = synth.posterior_params(b5, (d5,), seed=5)
        sl = slice(rank * b_loc, (rank + 1) * b_loc)
        enc = torch.cat([mu_g[sl], lv_g[sl]], 1).to(dev)
        del mu_g, lv_g
        mu5 = enc[:, :d5].detach().requires_grad_()
        lv5 = enc[:, d5:].detach().requires_grad_()

        def c5(i):
            mu5.grad = None
            lv5.grad = None
            if head is None:
                m_d, s_ = ops.mpd(mu5, lv5)
            else:
                _, m_d, s_, _ = head.kl_mpd(mu5, lv5)
            (-torch.nn.functional.logsigmoid(m_d).sum() + s_).backward()

        flop = 16.0 * b5 * b5 * d5
        n_it = 2 if flop > 2e12 else (5 if flop > 1e11 else 20)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        t = time_kernel(c5, n_it, stream, 3)
        if world > 1:
            tt = torch.tensor([t], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t = float(tt)
        if world > 1:
            lib_ = ops._lib_()
            extra = lib_.mae_mpd_workspace_bytes_rows(b5, d5, rank * b_loc, b_loc) - lib_.mae_mpd_workspace_bytes(b5, d5, 0)
            path = "sharded row block, " + ("stored-KL backward (KL_ij and KL_ji of the own rows: two tile passes)"
                                            if extra <= ops.MPD_KL_BUDGET_BYTES else "recompute backward")
        else:
            path = "stored-KL backward" if ops._mpd_pick_path(b5, d5, 0) != 2 else "recompute backward"
        # tensor-core tiles (3xTF32 tcgen05.mma, csrc/mpd_tc.cuh) serve padded batches >= 2048 with K = 2D a multiple of 128, <= 512
        tc_prev = ops._lib_().mae_set_mpd_tensor_cores(1)
        ops._lib_().mae_set_mpd_tensor_cores(tc_prev)
        tensor = bool(tc_prev) and b5 >= 2048 and (2 * d5) % 128 == 0 and 2 * d5 <= 512 and "stored" in path
        entry = {"metric": "mpd_fwd_bwd_pairs_per_s", "value": b5 * b5 / t, "unit": "pairs/s", "n_gpus": world, "scaling": "strong",
                 "ms": t * 1e3, "tflops_algorithmic": flop / t / 1e12, "path": path + (", tensor-core tiles (3xTF32 tcgen05)" if tensor else ", FP32 FFMA tiles")}
        if tensor:
            # executed tensor work: 3 tf32 products per algorithmic product; forward 2 B^2 D MACs (x2 for a row block's transposed
            # pass), stored-KL backward 4 B^2 D MACs.  tf32 dense peak = half the measured bf16 peak (MEASURED_PEAKS.json).
            macs = (2.0 * (2 if world > 1 else 1) + 4.0) * b5 * b5 * d5
            tf32_peak = 0.5 * measured_bf16_tflops() * world
            entry["roofline"] = {"bound": "tf32 tensor (3xTF32)", "tensor_tflops_executed": 3 * 2 * macs / t / 1e12,
                                 "peak_tflops": tf32_peak, "peak_source": "0.5 x measured bf16 dense (MEASURED_PEAKS.json)",
                                 "frac": 3 * 2 * macs / t / 1e12 / tf32_peak,
                                 "vs_ffma_peak": (flop / t / 1e12 / (ffma_peak_tflops * world)) if ffma_peak_tflops > 0 else None}
        else:
            entry["roofline"] = {"bound": "fp32 ffma", "peak_tflops_measured": ffma_peak_tflops * world,
                                 "frac": (flop / t / 1e12 / (ffma_peak_tflops * world)) if ffma_peak_tflops > 0 else None}
        res["c5_mpd_b%d_d%d" % (b5, d5)] = entry
        del enc, mu5, lv5
        torch.cuda.empty_cache()
    if head is not None:
        head.check()
    return res


# ------------------------------------------------------------------------------------------- CPU / reference
def c3_oracle_step_fn(device, batch=BATCH, latent=LATENT):
    """The oracle port of the reference's ATen path for one C3 step (fwd+bwd) on ``device`` (host cores or, for the
    same-box eager bar, the B200)."""
    from oracle import mae_oracle as O
    h = {k: v.to(device) for k, v in c3_host_inputs(0, batch, latent).items()}
    enc = h["enc"].clone().requires_grad_()
    raw = h["raw"].clone().requires_grad_()
    one = torch.ones((), device=device)

    def step():
        enc.grad = None
        raw.grad = None
        mu, lv = enc.chunk(2, 1)
        out, z = O.loss_head_color(mu, lv, h["eps"], lambda z_: raw, h["x"], 10, ETA, GAMMA, FREE_BITS)
        torch.autograd.backward([out[0], z], [one, h["gz"]])
        return out[0].detach()

    return step


def gpu_eager_baseline(dev, value_ours):
    """SURVEY 8d / 2a 'same-box bar': the reference's own eager-PyTorch formulation of the path (oracle port = the same
    ATen ops in the same order) executed on the B200 with CUDA tensors, timed with CUDA events.  C3 fwd+bwd and one C4
    image (K = 512).  Not the product path: reported beside it."""
    from oracle import mae_oracle as O
    stream = torch.cuda.current_stream()
    step = c3_oracle_step_fn(dev)
    t_c3 = time_kernel(lambda i: step(), 10, stream, 3)
    g = torch.Generator(device=dev).manual_seed(3)
    k, d = 512, 64
    raw = torch.randn(k, 100, 32, 32, device=dev, generator=g)
    x = (torch.randint(0, 256, (1, 3, 32, 32), device=dev, generator=g).float() / 127.5 - 1.0)
    mu = torch.randn(1, d, device=dev, generator=g)
    lv = torch.randn(1, d, device=dev, generator=g).clamp(-7, 7)
    z = mu[:, None] + torch.randn(1, k, d, device=dev, generator=g) * (0.5 * lv).exp()[:, None]
    with torch.no_grad():
        t_c4 = time_kernel(lambda i: O.iw_eval_color(raw, x, z, mu, lv, 10), 5, stream, 3)
    out = {"what": "oracle port (the reference's ATen op sequence) on CUDA tensors, eager, CUDA events; median of 3 blocks",
           "c3_samples_per_s": BATCH / t_c3, "c3_ms_per_step": t_c3 * 1e3, "c3_speedup_ours_device": value_ours * t_c3 / BATCH,
           "c4_images_per_s": 1.0 / t_c4, "c4_ms_per_image": t_c4 * 1e3}
    del raw
    torch.cuda.empty_cache()
    return out


def cpu_baseline_c3(budget_s=12.0):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step = c3_oracle_step_fn("cpu")
    step()
    t0 = time.perf_counter()
    n = 0
    while n < 3 or (time.perf_counter() - t0 < budget_s and n < 200):
        step()
        n += 1
    dt = (time.perf_counter() - t0) / n
    return {"value": BATCH / dt, "unit": "samples/s", "cores": cores, "kind": "port",
            "sample": "%d full C3 steps (B=64) of oracle/mae_oracle.py (torch CPU restatement of the reference's ATen path), "
                      "%.1f ms/step" % (n, dt * 1e3)}


def bench_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle port: the reference is pure
    Python/ATen and /root/reference does not exist on the GPU box) on all host cores, same config and metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step = c3_oracle_step_fn("cpu")
    warm = max(1, min(args.warmup, 3))
    for _ in range(warm):
        step()
    steps = max(1, min(args.steps, 40))
    t0 = time.perf_counter()
    for _ in range(steps):
        loss = float(step())
    dt = time.perf_counter() - t0
    value = BATCH * steps / dt
    world = int(os.environ.get("WORLD_SIZE", "1"))
    return {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s",
        "n_gpus": world, "steps": steps, "warmup": warm, "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": BATCH, "parallelism": "host cpu (one process, all cores; one batch of 64)"},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": "%d C3 steps (B=64) of the oracle port on %d host threads" % (steps, cores)},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "loss": loss,
    }


def _teardown(steps_obj, timeout_s=20.0):
    """Destroy the captured graphs, then the NCCL communicator.  A watchdog leaves the process if the communicator
    teardown does not return (seen with torch 2.11 / NCCL 2.28.9 when graphs that captured peer / collective work are
    still alive at destroy time)."""
    import gc
    import torch.distributed as dist
    done = threading.Event()

    def watchdog():
        if not done.wait(timeout_s):
            print("[bench] NCCL teardown did not return within %.0f s; leaving without it" % timeout_s, file=sys.stderr, flush=True)
            os._exit(0)

    threading.Thread(target=watchdog, daemon=True).start()
    for st in steps_obj:
        st.graph = None
    steps_obj.clear()
    gc.collect()
    torch.cuda.synchronize()
    dist.destroy_process_group()
    done.set()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="launch eagerly instead of replaying CUDA graphs")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary workloads (C1, C4, C5, GPU-eager bar)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--quick", action="store_true", help="C5: three small shapes only")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: library chatter written to fd 1 (NCCL prints its version banner there at
    # N > 1) is sent to stderr while the benchmark runs
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    steps_obj = None
    if args.impl == "reference":
        out = bench_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (the hot path has no CPU fallback); use --impl reference for the CPU arm")
        out, steps_obj = bench_ours(args)
    sys.stdout.flush()
    sys.stderr.flush()
    if out is not None:
        os.write(real_stdout, (json.dumps(out) + "\n").encode())
    os.close(real_stdout)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and args.impl == "ours":
        _teardown(steps_obj)


if __name__ == "__main__":
    main()
