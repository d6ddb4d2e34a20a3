#!/usr/bin/env python
"""bench.py -- throughput of the B200-native MultiTypeTree likelihood hot path.

Headline workload (BASELINE.json configs[1], "cfg2"): batched structured-coalescent logP of
simulated 2-deme 100-tip multi-type trees, density only.  One step = one pass of the density
kernel over one batch of B trees per GPU (B = 2^19 by default = 128 x the 4096-tree batch the
config names, so that the ~1.2 GB of inputs exceed the 126 MB L2; the literal 4096-tree batch
is timed as well, with an L2 flush before every launch, and reported under "cfg2_b4096").

  value         trees/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e           same metric through the reference-facing C-ABI call mtt_sc_logp_owner_batch
                with pinned HOST buffers: host->device copies and the device->host read of logP are
                inside the timed region
  roofline      algorithmic bytes (12 B per event + 8 B per tree) / kernel time vs
                the measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline  the oracle's faithful restatement of the reference algorithm on the host cores

  --impl reference   times the reference's CPU algorithm (oracle port; no JVM exists here) on
                     the same config with all host threads.
  --workload peel|peel_chains|peel_sharded|retype|untyped   the other pieces (configs 3, 4 and
                     5, incremental peeling on resident chains, and the untyped-tree particle
                     density) as separate bench lines.

Multi-GPU: one process per GPU (torchrun); trees are sharded across ranks, no collective on
the data path (weak scaling).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "structured-coalescent+tree logP evals/sec (batched trees) at 1/2/4/8 B200"
FALLBACK_HBM_GBS = 6650.0


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        # One poller for the whole job: rank 0 watches every GPU in use (eight concurrent
        # nvidia-smi loops take driver locks that show up in steps with a host sync).
        self.index, self.lines, self.proc = index, [], None
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.active = int(os.environ.get("RANK", "0")) == 0

    def __enter__(self):
        if not self.active:
            return self
        try:
            gpus = ",".join(str(i) for i in range(self.world)) if self.world > 1 else str(self.index)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", gpus, "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_setup(n_gpus):
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return world, rank, local


def max_over_ranks(x, world):
    if world == 1:
        return x
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(world):
    import torch
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def to_dev(a):
    """device copy with 64 elements of slack (bulk copies over-read <= 15 B)."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a).ravel())
    out = torch.zeros(t.numel() + 64, dtype=t.dtype, device="cuda")
    out[:t.numel()] = t.cuda()
    return out


# ------------------------------------------------------------------ cfg2: density

def make_cfg2(B, rank, n=100, d=2, threads=8):
    import multitypetree_b200 as mtt
    m, N = mtt.synthetic_model(d)
    tb = mtt.simulate_trees(B, n, d, m, N, seed=2000 + rank * B, nThreads=threads)
    return tb, m, N


def cfg2_config(tb, B, world):
    """The `config` object of the cfg2 line: built by BOTH arms from the same generated batch,
    so that the driver sees the same workload description on either side."""
    ev = tb.n_events()
    return {
        "workload": "cfg2: batched structured-coalescent logP, 2-deme 100-tip trees, "
                    "density only (event extraction + time ordering + interval scan)",
        "batch_per_gpu": B, "tips": 100, "demes": 2,
        "events_per_tree_mean": float(ev.mean()), "events_per_tree_max": int(ev.max()),
        "input_bytes_per_gpu": int(tb.nbytes()),
        "l2": "inputs (%.2f GB) exceed the 126 MB L2; no flush needed" % (tb.nbytes() / 1e9),
        "parallelism": "trees sharded across %d GPU(s), no data-path collective" % world,
    }


def cpu_density(tb, m, N, sample, threads):
    """oracle faithful path (the reference's O(E x active) algorithm) on `sample` trees."""
    from oracle import oracle as orc
    sub = tb.slice(0, sample)
    orc.sc_logp_batch(tb.slice(0, min(sample, 256)), m, N, variant=orc.FAITHFUL, nThreads=threads)
    # repeat the sample until about 20 CPU-seconds of work have been timed
    reps, dt = 0, 0.0
    t0 = time.perf_counter()
    while reps < 64:
        orc.sc_logp_batch(sub, m, N, variant=orc.FAITHFUL, nThreads=threads)
        reps += 1
        dt = time.perf_counter() - t0
        if dt * threads >= 20.0:
            break
    return sample * reps / dt, dt


def fp64_peak():
    """Sustained DFMA throughput of one B200 in TFLOP/s: measured by tools/ubench/dfma_peak.cu
    (profiles/r2_dfma_peak.json), else nominal 148 SM x 64 DFMA/clk x 2 x 1.965 GHz."""
    path = os.path.join(ROOT, "profiles", "r2_dfma_peak.json")
    try:
        with open(path) as f:
            return float(json.load(f)["dfma_tflops"]), "measured (profiles/r2_dfma_peak.json)"
    except Exception:
        return 148 * 64 * 2 * 1.965e9 / 1e12, "nominal"


def bench_density(args):
    import torch
    import multitypetree_b200 as mtt
    from multitypetree_b200 import _lib
    world, rank, local = dist_setup(args.gpus)
    cores = os.cpu_count() or 1
    threads = max(1, cores // world)
    B = args.batch
    tb, m, N = make_cfg2(B, rank, threads=threads)
    ev = tb.n_events()
    alg_bytes = float(12 * ev.sum() + 8 * B)

    L = _lib.lib()
    ctx = mtt.Context(local)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    L.mtt_sc_set_max_tree_events(ctx.handle, int(ev.max()))

    dm, dN = to_dev(m), to_dev(N)
    # change lists in the owner layout (mtt_sc_logp_owner_batch*): what the kernels read
    dev = [to_dev(x) for x in (tb.nodeOff, tb.height, tb.nodeType, tb.chgNode, tb.chgBase,
                               tb.chgTime, tb.chgType)]
    logP = torch.zeros(B, dtype=torch.float64, device="cuda")

    def step(Bs=B, out=logP):
        ctx.check(L.mtt_sc_logp_owner_batch_dev(
            ctx.handle, Bs, 2, 1, dm.data_ptr(), dN.data_ptr(), dev[0].data_ptr(),
            dev[1].data_ptr(), None, dev[2].data_ptr(), dev[3].data_ptr(), dev[4].data_ptr(),
            dev[5].data_ptr(), dev[6].data_ptr(), 0, out.data_ptr(), None))

    for _ in range(args.warmup):
        step()
    barrier(world)
    launches0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
        barrier(world)
    ms = max_over_ranks(e0.elapsed_time(e1), world)
    launches = ctx.launch_count() - launches0
    ms_per_step = ms / args.steps
    value = B * world / (ms_per_step * 1e-3)
    kernel_ms = e0.elapsed_time(e1) / args.steps          # one kernel launch per step
    hbm, hbm_src = peaks()
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9

    # ---- the literal 4096-tree batch of the config, L2 flushed before every launch
    small = None
    if rank == 0:
        flush = torch.empty(192 << 20, dtype=torch.uint8, device="cuda")
        out4k = torch.zeros(4096, dtype=torch.float64, device="cuda")
        ts = []
        for i in range(args.warmup + 10):
            flush.fill_(i & 255)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            step(4096, out4k)
            b.record(stream)
            torch.cuda.synchronize()
            if i >= args.warmup:
                ts.append(a.elapsed_time(b))
        small = {"batch": 4096, "ms": float(np.median(ts)),
                 "value": 4096 / (float(np.median(ts)) * 1e-3), "l2": "flushed before each launch"}
        del flush

    # ---- end to end through the C-ABI with pinned host buffers
    host = [np.ascontiguousarray(x) for x in (m, N, tb.nodeOff, tb.height, tb.nodeType,
                                              tb.chgNode, tb.chgBase, tb.chgTime, tb.chgType)]
    out_h = np.empty(B, np.float64)
    pinned = []
    for a in host + [out_h]:
        if L.mtt_host_register(a.ctypes.data, a.nbytes) == 0:
            pinned.append(a)
    h2d = int(sum(a.nbytes for a in host))
    d2h = int(out_h.nbytes)

    def e2e_step():
        ctx.check(L.mtt_sc_logp_owner_batch(
            ctx.handle, B, 2, 1, _lib.ptr(host[0]), _lib.ptr(host[1]), _lib.ptr(host[2]),
            _lib.ptr(host[3]), None, _lib.ptr(host[4]), _lib.ptr(host[5]), _lib.ptr(host[6]),
            _lib.ptr(host[7]), _lib.ptr(host[8]), 0, _lib.ptr(out_h), None))

    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        e2e_step()
    barrier(world)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier(world)
    e2e_s = max_over_ranks(time.perf_counter() - t0, world)
    e2e_value = B * world * e2e_steps / e2e_s
    # device-resident and end-to-end paths must agree bit for bit
    assert np.array_equal(out_h, logP.cpu().numpy()), "e2e and device-resident results differ"

    # ---- the same trees through the fixed-tips entry: all trees of the workload are over one
    #      taxon set (as all trees of a BEAST analysis are), so the tips are installed once
    #      outside the timed region (like the alignment of the peeling handle) and a step moves
    #      only internal nodes and change lists.  Reported beside `e2e`, not instead of it.
    nL = 100
    nN = 2 * nL - 1
    H2 = tb.height.reshape(B, nN)
    T2 = tb.nodeType.reshape(B, nN)
    assert np.all(H2[:, :nL] == H2[0, :nL]) and np.all(T2[:, :nL] == T2[0, :nL])
    tipH, tipT = np.ascontiguousarray(H2[0, :nL]), np.ascontiguousarray(T2[0, :nL])
    ctx.check(L.mtt_sc_set_tips(ctx.handle, nL, _lib.ptr(tipH), _lib.ptr(tipT)))
    intH, intT = np.ascontiguousarray(H2[:, nL:]), np.ascontiguousarray(T2[:, nL:])
    out_t = np.empty(B, np.float64)
    pinned_t = [a for a in (intH, intT, out_t) if L.mtt_host_register(a.ctypes.data, a.nbytes) == 0]
    h2d_t = int(intH.nbytes + intT.nbytes + sum(host[i].nbytes for i in (0, 1, 5, 6, 7, 8)))

    def tips_step():
        ctx.check(L.mtt_sc_logp_tips_batch(
            ctx.handle, B, 2, 1, _lib.ptr(host[0]), _lib.ptr(host[1]), _lib.ptr(intH), None,
            _lib.ptr(intT), _lib.ptr(host[5]), _lib.ptr(host[6]), _lib.ptr(host[7]),
            _lib.ptr(host[8]), 0, _lib.ptr(out_t), None))

    for _ in range(2):
        tips_step()
    barrier(world)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        tips_step()
    barrier(world)
    tips_s = max_over_ranks(time.perf_counter() - t0, world)
    tips_value = B * world * e2e_steps / tips_s
    assert np.array_equal(out_t, out_h), "fixed-tips and general entries differ"
    for a in pinned + pinned_t:
        L.mtt_host_unregister(a.ctypes.data)

    if rank != 0:
        return
    sample = min(B, 1 << 19)
    cpu_v, cpu_dt = cpu_density(tb, m, N, sample, cores)
    # parity spot-check of the timed results against the oracle (checker only)
    from oracle import oracle as orc
    want = orc.sc_logp_batch(tb.slice(0, 4096), m, N, variant=orc.SORTED, nThreads=cores)
    rel = float(np.max(np.abs(out_h[:4096] - want) / np.abs(want)))
    assert rel < 1e-9, "parity check failed: %g" % rel

    line = {
        "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (structured-coalescent Gillespie simulator, SURVEY 8d; seeds 2000+)",
        "config": cfg2_config(tb, B, world),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s",
                     "frac": achieved / hbm,
                     # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture
                     # (profiles/r2_k1_density_ncu.txt: 153.69 MB for 65536 of these trees),
                     # scaled to this launch's tree count
                     "traffic": 153.69e6 / 65536 * B, "peak_source": hbm_src,
                     "kernel": "k1_density_warp", "algorithmic_bytes_per_launch": alg_bytes,
                     "kernel_ms": kernel_ms},
        "cpu_baseline": {"value": cpu_v, "unit": "evals/s", "cores": cores, "kind": "port",
                         "sample": "%d trees of the same batch (repeated until ~20 CPU-s), oracle "
                                   "faithful restatement of the reference algorithm (not a JVM), "
                                   "%.1f s wall on %d threads" % (sample, cpu_dt, cores)},
        "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "pinned": len(pinned) == len(host) + 1},
        "e2e_fixed_tips": {"value": tips_value, "unit": "evals/s", "h2d_bytes_per_step": h2d_t,
                           "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                           "note": "mtt_sc_set_tips once + mtt_sc_logp_tips_batch per step: the "
                                   "trees share their 100 tips, only internal nodes and change "
                                   "lists cross the link; results bit-identical to e2e"},
        "gpu_launches": int(launches),
        "clocks": clk.summary(),
        "cfg2_b4096": small,
        "parity_max_rel_err_vs_oracle": rel,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------ reference arm

def bench_reference(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # the GPU arm's batch (same generator, seed and size); each step times a bounded sample of it
    B = args.batch
    full, m, N = make_cfg2(B, 0, threads=cores)
    sample = min(B, 1 << 17)
    tb = full.slice(0, sample)
    from oracle import oracle as orc
    for _ in range(max(1, min(args.warmup, 3))):
        orc.sc_logp_batch(tb, m, N, variant=orc.FAITHFUL, nThreads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.sc_logp_batch(tb, m, N, variant=orc.FAITHFUL, nThreads=cores)
    dt = time.perf_counter() - t0
    v = sample * args.steps / dt
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "evals/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (structured-coalescent Gillespie simulator, SURVEY 8d; seeds 2000+)",
        "config": cfg2_config(full, B, world),
        "cpu_baseline": {"value": v, "unit": "evals/s", "cores": cores, "kind": "port",
                         "sample": "the first %d trees of the batch per step; C restatement of "
                                   "StructuredCoalescentTreeDensity (the reference is Java and no "
                                   "JVM exists on this box)" % sample},
        "e2e": {"value": v, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------ other pieces

def bench_peel(args):
    """cfg3: 4-deme 1000-tip trees, 10k-site alignment, density + peeling per evaluation."""
    import torch
    import multitypetree_b200 as mtt
    from multitypetree_b200 import _lib
    from multitypetree_b200.likelihood import (Alignment, HKY, SiteModel, TreeLikelihood,
                                               simulate_alignment)
    world, rank, local = dist_setup(args.gpus)
    cores = os.cpu_count() or 1
    n, d, sites, B = 1000, 4, 10000, args.batch if args.batch != (1 << 19) else 512
    m, N = mtt.synthetic_model(d)
    tb = mtt.simulate_trees(B, n, d, m, N, seed=3000 + rank * B, nThreads=max(1, cores // world))
    nN = 2 * n - 1
    par, hgt = tb.parent.reshape(B, nN), tb.height.reshape(B, nN)
    f = np.array([.25] * 4)
    # one alignment for the whole job: every rank evolves the same sites down the same reference
    # tree at the same clock rate, so the ranks do equal work (weak scaling over trees)
    ref = mtt.simulate_trees(1, n, d, m, N, seed=3000)
    clock = 0.2 / float(ref.height.max())
    al = Alignment(states=simulate_alignment(ref.parent, ref.height, sites, 2.0, f, clock, seed=3000))
    P = al.getPatternCount()
    ctx = mtt.Context(local)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    L = _lib.lib()
    tl = TreeLikelihood(data=al, siteModel=SiteModel(HKY(2.0, f), mutationRate=clock), ctx=ctx)
    dpar, dhgt = to_dev(par), to_dev(hgt)
    out = torch.zeros(B, dtype=torch.float64, device="cuda")
    dm, dN = to_dev(m), to_dev(N)
    dev = [to_dev(x) for x in (tb.nodeOff, tb.height, tb.nodeType, tb.chgNode, tb.chgBase,
                               tb.chgTime, tb.chgType)]
    dens = torch.zeros(B, dtype=torch.float64, device="cuda")
    L.mtt_sc_set_max_tree_events(ctx.handle, int(tb.n_events().max()))

    def step():
        ctx.check(L.mtt_peel_logl_batch_dev(tl.handle, B, dpar.data_ptr(), dhgt.data_ptr(),
                                            out.data_ptr()))
        ctx.check(L.mtt_sc_logp_owner_batch_dev(
            ctx.handle, B, d, 1, dm.data_ptr(), dN.data_ptr(), dev[0].data_ptr(),
            dev[1].data_ptr(), None, dev[2].data_ptr(), dev[3].data_ptr(), dev[4].data_ptr(),
            dev[5].data_ptr(), dev[6].data_ptr(), 0, dens.data_ptr(), None))

    for _ in range(args.warmup):
        step()
    barrier(world)
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
        barrier(world)
    own_ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        print("rank %d: %.3f ms/step" % (rank, own_ms), file=sys.stderr, flush=True)
    ms = max_over_ranks(e0.elapsed_time(e1), world) / args.steps
    launches = ctx.launch_count() - l0
    # e2e: host buffers through the C-ABI (page-locked, as the contract asks: par / hgt are views
    # of tb.parent / tb.height)
    pinned = [a for a in (tb.parent, tb.height, tb.nodeType, tb.chgNode, tb.chgBase, tb.chgTime,
                          tb.chgType, tb.nodeOff)
              if L.mtt_host_register(a.ctypes.data, a.nbytes) == 0]
    lp = tl.calculateLogPBatch(par, hgt)
    ld = mtt.sc_logp_batch(tb, m, N, ctx=ctx)
    barrier(world)
    t0 = time.perf_counter()
    for _ in range(3):
        lp = tl.calculateLogPBatch(par, hgt)
        ld = mtt.sc_logp_batch(tb, m, N, ctx=ctx)
    barrier(world)
    e2e = B * world * 3 / max_over_ranks(time.perf_counter() - t0, world)
    for a in pinned:
        L.mtt_host_unregister(a.ctypes.data)
    if rank != 0:
        return
    from oracle import oracle as orc
    k = min(B, 64)          # ~20 CPU-seconds of the restated BeerLikelihoodCore
    t0 = time.perf_counter()
    want = orc.tree_likelihood_batch(al.patterns, al.patternWeight, par[:k], hgt[:k], [clock], [1.0],
                                     f, 2.0, nThreads=min(cores, k))
    cpu_dt = time.perf_counter() - t0
    rel = float(np.max(np.abs(lp[:k] - want) / np.abs(want)))
    S, C = 4, 1
    alg = 8.0 * S * C * P * (2 * n - 3) + 4.0 * n * P + 8.0 * P + 8.0 * S * S * C * (2 * n - 2)
    flop = (n - 1) * P * C * (4 * S * S + S)
    hbm, src = peaks()
    print(json.dumps({
        "metric": "cfg3 (density + peeling) evaluations/s", "value": B * world / (ms * 1e-3),
        "unit": "evals/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg3: 4-deme 1000-tip trees x 10k-site alignment "
                               "(%d patterns), density + peeling" % P, "batch_per_gpu": B},
        "roofline": {"bound": "hbm", "achieved": alg * B / (ms * 1e-3) / 1e9, "peak": hbm,
                     "unit": "GB/s", "frac": alg * B / (ms * 1e-3) / 1e9 / hbm, "traffic": None,
                     "peak_source": src, "kernel": "k2_peel",
                     "note": "algorithmic bytes of the materialised-partials model (SURVEY 8d); "
                             "the kernel keeps partials on chip, so frac may exceed 1",
                     "fp64_tflops": flop * B / (ms * 1e-3) / 1e12,
                     "fp64_peak_tflops": fp64_peak()[0], "fp64_peak_source": fp64_peak()[1],
                     "fp64_frac": flop * B / (ms * 1e-3) / 1e12 / fp64_peak()[0]},
        "cpu_baseline": {"value": k / cpu_dt, "unit": "evals/s", "cores": min(cores, k),
                         "kind": "port", "sample": "%d trees, peeling only" % k},
        "e2e": {"value": e2e, "unit": "evals/s",
                "h2d_bytes_per_step": int(par.nbytes + hgt.nbytes + tb.nbytes()),
                "d2h_bytes_per_step": int(16 * B), "pinned": len(pinned) == 8},
        "gpu_launches": int(launches), "clocks": clk.summary(),
        "parity_max_rel_err_vs_oracle": rel}))


def bench_peel_chains(args):
    """cfg3 shape, MCMC regime: 64 resident chains, one node-height proposal per chain and step
    (MultiTypeUniform) given as the changed node (mtt_peel_chains_update_sparse), incremental
    peeling + accept / reject, against a
    full re-evaluation of the same proposals (mtt_peel_logl_batch).  Host-buffer API on both
    sides: the planner's host time is part of the number."""
    import torch
    import multitypetree_b200 as mtt
    from multitypetree_b200 import _lib
    from multitypetree_b200.likelihood import (Alignment, HKY, SiteModel, TreeLikelihood,
                                               simulate_alignment)
    world, rank, local = dist_setup(args.gpus)
    cores = os.cpu_count() or 1
    n = args.tips or 1000
    d, sites = 4, args.sites or 10000
    K = args.batch if args.batch != (1 << 19) else 64
    m, N = mtt.synthetic_model(d)
    tb = mtt.simulate_trees(K, n, d, m, N, seed=3000 + rank * K, nThreads=max(1, cores // world))
    nN = 2 * n - 1
    par = np.ascontiguousarray(tb.parent.reshape(K, nN), np.int32).copy()
    hgt = np.ascontiguousarray(tb.height.reshape(K, nN), np.float64).copy()
    f = np.array([.25] * 4)
    ref = mtt.simulate_trees(1, n, d, m, N, seed=3000)
    clock = 0.2 / float(ref.height.max())
    al = Alignment(states=simulate_alignment(ref.parent, ref.height, sites, 2.0, f, clock, seed=3000))
    P = al.getPatternCount()
    ctx = mtt.Context(local)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    tl = TreeLikelihood(data=al, siteModel=SiteModel(HKY(2.0, f), mutationRate=clock), ctx=ctx)
    ch = tl.chains(K)
    # children of every node (the topology does not change in this workload)
    kids = np.full((K, nN, 2), -1, np.int64)
    for k in range(K):
        order = np.argsort(par[k], kind="stable")
        order = order[par[k][order] >= 0]
        kids[k, par[k][order][0::2], 0] = order[0::2]
        kids[k, par[k][order][1::2], 1] = order[1::2]
    rows = np.arange(K)
    rng = np.random.default_rng(3000 + rank)

    def propose():
        v = rng.integers(n, nN, K)
        lo = np.maximum(hgt[rows, kids[rows, v, 0]], hgt[rows, kids[rows, v, 1]])
        pp = par[rows, v]
        hi = np.where(pp >= 0, hgt[rows, np.maximum(pp, 0)], lo + 1.0)
        ph = hgt.copy()
        ph[rows, v] = lo + (hi - lo) * rng.uniform(0.05, 0.95, K)
        # the proposal in sparse form: one changed node per chain, with its parent and height
        node = v.astype(np.int32)
        return ph, (node, np.ascontiguousarray(par[rows, v]), np.ascontiguousarray(ph[rows, v]))

    ch.update(par, hgt)
    ch.accept()
    L = _lib.lib()
    got = np.empty(K, np.float64)
    off = np.arange(K + 1, dtype=np.int64)
    off_p, got_p = _lib.ptr(off), _lib.ptr(got)
    t_upd = t_full = dev_ms = 0.0
    entries = dev_entries = 0
    err = 0.0
    launches = 0
    # pass 1 (warm-up + timed steps): throughput through the host API; pass 2 (as many steps
    # again): the same with the library's CUDA events around its launch, for the device time --
    # reading the events back costs host time, so it is not part of pass 1
    with ClockSampler(local) as clk:
        for it in range(args.warmup + 2 * args.steps):
            ph, (node, npar, nhgt) = propose()
            timed = args.warmup <= it < args.warmup + args.steps
            if it == args.warmup + args.steps:
                ch.set_timing(True)
            acc = (rng.random(K) < 0.5).astype(np.uint8)
            lf0 = ctx.launch_count()
            # the C-ABI calls a Java shim would make, on the caller's buffers
            t0 = time.perf_counter()
            ctx.check(L.mtt_peel_chains_update_sparse(ch._s, off_p, _lib.ptr(node), _lib.ptr(npar),
                                                      _lib.ptr(nhgt), None, None, None, got_p))
            ctx.check(L.mtt_peel_chains_accept(ch._s, _lib.ptr(acc)))
            dt = time.perf_counter() - t0
            acc = acc != 0
            if timed:
                t_upd += dt
                entries += ch.last_entries
                launches += ctx.launch_count() - lf0        # (not the comparison's below)
            elif it >= args.warmup + args.steps:
                dev_ms += ch.last_device_ms
                dev_entries += ch.last_entries
            t0 = time.perf_counter()
            want = tl.calculateLogPBatch(par, ph)
            if timed:
                t_full += time.perf_counter() - t0
            err = max(err, float(np.max(np.abs(got - want) / np.abs(want))))
            hgt[acc] = ph[acc]
    barrier(world)
    t_upd = max_over_ranks(t_upd * 1e3, world) / 1e3
    if rank != 0:
        return
    steps = args.steps
    C = 1
    stride = (P + 255) // 256 * 256
    alg = dev_entries * stride * 2.0 * (32 * C + 8)     # one vector read + one written per entry
    hbm, src = peaks()
    print(json.dumps({
        "metric": "MCMC regime: proposals evaluated/s (incremental peeling, resident chains)",
        "value": K * world * steps / t_upd, "unit": "chain-steps/s", "n_gpus": world,
        "steps": steps, "warmup": args.warmup, "ms_per_step": t_upd / steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "%s: %d chains x %d-tip trees x %d patterns, one node-height "
                               "proposal per chain and step, accept with probability 1/2"
                               % ("cfg3 shape" if n == 1000 else "resident chains", K, n, P),
                   "chains_per_gpu": K,
                   "nodes_recomputed_per_proposal": entries / (K * steps),
                   "resident_bytes": int(K) * 2 * (n - 1) * (4 * C + 1) * stride * 8},
        "roofline": {"bound": "hbm", "achieved": alg / (dev_ms * 1e-3) / 1e9 if dev_ms else None,
                     "peak": hbm, "unit": "GB/s",
                     "frac": alg / (dev_ms * 1e-3) / 1e9 / hbm if dev_ms else None,
                     "traffic": None, "peak_source": src, "kernel": "k2_peel_update",
                     "device_ms_per_step": dev_ms / steps,
                     "note": "CUDA events around the launch of every update of a second pass of "
                             "%d steps; algorithmic bytes = (32 C + 8) read + written per pattern "
                             "and recomputed node" % steps},
        "full_reevaluation": {"value": K * steps / t_full, "unit": "chain-steps/s",
                              "note": "mtt_peel_logl_batch on the same proposals, host buffers"},
        "e2e": {"value": K * world * steps / t_upd, "unit": "chain-steps/s",
                "h2d_bytes_per_step": int(entries / steps * 80 + 4 * (2 * K + 1) + 8 * K),
                "d2h_bytes_per_step": int(8 * K),
                "note": "the API takes host buffers: value is already end to end"},
        "gpu_launches": int(launches), "clocks": clk.summary(),
        "parity_max_rel_err_vs_full": err}))


def bench_peel_sharded(args):
    """cfg4 shape: 8-deme 5000-tip trees, site patterns sharded across ranks (125k patterns per
    GPU = 1M over 8 GPUs), per-rank partial logL combined with one NCCL all-reduce."""
    import torch
    import torch.distributed as dist
    import multitypetree_b200 as mtt
    from multitypetree_b200 import _lib
    from multitypetree_b200.likelihood import (Alignment, HKY, SiteModel, TreeLikelihood,
                                               simulate_alignment)
    world, rank, local = dist_setup(args.gpus)
    n, d = args.tips or 5000, 8
    sites = args.sites or 125000
    B = args.batch if args.batch != (1 << 19) else 8
    m, N = mtt.synthetic_model(d)
    tb = mtt.simulate_trees(B, n, d, m, N, seed=4000, nThreads=max(1, (os.cpu_count() or 1) // world))
    nN = 2 * n - 1
    par, hgt = tb.parent.reshape(B, nN), tb.height.reshape(B, nN)
    f = np.array([.25] * 4)
    clock = 0.2 / float(hgt[0].max())
    # this rank's shard of the alignment: its own sites, evolved on the same reference tree
    al = Alignment(states=simulate_alignment(par[0], hgt[0], sites, 2.0, f, clock, seed=4000 + rank),
                   compress=False)
    ctx = mtt.Context(local)
    stream = torch.cuda.current_stream()
    side = torch.cuda.Stream()
    torch.cuda.set_stream(side)
    ctx.set_stream(side.cuda_stream)
    L = _lib.lib()
    tl = TreeLikelihood(data=al, siteModel=SiteModel(HKY(2.0, f), mutationRate=clock), ctx=ctx)
    dpar, dhgt = to_dev(par), to_dev(hgt)
    out = torch.zeros(B, dtype=torch.float64, device="cuda")

    def step():
        ctx.check(L.mtt_peel_logl_batch_dev(tl.handle, B, dpar.data_ptr(), dhgt.data_ptr(),
                                            out.data_ptr()))
        if world > 1:
            dist.all_reduce(out, op=dist.ReduceOp.SUM)      # 8 B per tree over NVLink

    for _ in range(args.warmup):
        step()
    barrier(world)
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record(side)
        for _ in range(args.steps):
            step()
        e1.record(side)
        barrier(world)
    ms = max_over_ranks(e0.elapsed_time(e1), world) / args.steps
    total = out.cpu().numpy().copy()
    # parity: this rank's partial against the oracle on a few patterns-worth of work
    from oracle import oracle as orc
    k = 1
    sub = Alignment(states=al.patterns[:, :2000], compress=False)
    tl2 = TreeLikelihood(data=sub, siteModel=SiteModel(HKY(2.0, f), mutationRate=clock), ctx=ctx)
    got = tl2.calculateLogPBatch(par[:k], hgt[:k])
    want = orc.tree_likelihood_batch(sub.patterns, sub.patternWeight, par[:k], hgt[:k], [clock],
                                     [1.0], f, 2.0)
    rel = float(np.max(np.abs(got - want) / np.abs(want)))
    assert rel < 1e-9, rel
    if rank != 0:
        return
    P, S, C = sites, 4, 1
    alg = 8.0 * S * C * P * (2 * n - 3) + 4.0 * n * P + 8.0 * P + 8.0 * S * S * C * (2 * n - 2)
    flop = (n - 1) * P * C * (4 * S * S + S)
    hbm, src = peaks()
    print(json.dumps({
        "metric": "cfg4 pattern-sharded tree-likelihood evaluations/s (whole alignment)",
        "value": B / (ms * 1e-3), "unit": "evals/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg4: 8-deme %d-tip trees, %d site patterns per GPU x %d GPUs, "
                               "NCCL all-reduce of the per-shard logL" % (n, sites, world),
                   "trees_per_step": B},
        "roofline": {"bound": "hbm", "achieved": alg * B / (ms * 1e-3) / 1e9, "peak": hbm,
                     "unit": "GB/s", "frac": alg * B / (ms * 1e-3) / 1e9 / hbm, "traffic": None,
                     "peak_source": src, "kernel": "k2_peel",
                     "note": "per-GPU algorithmic bytes of the materialised-partials model; "
                             "partials stay on chip", "fp64_tflops_per_gpu": flop * B / (ms * 1e-3) / 1e12},
        "logL_first_tree": float(total[0]), "gpu_launches": int(ctx.launch_count() - l0),
        "clocks": clk.summary(), "parity_max_rel_err_vs_oracle": rel}))


def bench_retype(args):
    """cfg5: NodeRetype-style uniformization sweep, chains x 10-deme models."""
    import multitypetree_b200 as mtt
    from multitypetree_b200.uniformization import UniformizationModels
    world, rank, local = dist_setup(args.gpus)
    d, chains = 10, 64 // world if world <= 64 else 1
    jobs_per_chain = args.batch if args.batch != (1 << 19) else 4096
    rng = np.random.default_rng(5000 + rank)
    ms_ = np.stack([mtt.synthetic_model(d)[0] * rng.uniform(0.5, 2.0) for _ in range(chains)])
    ctx = mtt.Context(local)
    um = UniformizationModels(ms_, ctx=ctx)
    nJobs = chains * jobs_per_chain
    brOff = np.arange(nJobs + 1) * 3
    nBr = 3 * nJobs
    t0_ = rng.uniform(0, 5, nBr)
    t1_ = t0_ + rng.exponential(1.5, nBr)
    a = rng.integers(0, d, nBr)
    b = rng.integers(0, d, nBr)
    model = np.repeat(np.arange(chains), 3 * jobs_per_chain)
    # Uniform windows: a job rarely needs more than a few dozen uniforms (1 + 2 nVirt per branch),
    # so the first pass ships 96 per job; jobs that report MTT_ST_NEED_MORE_UNIFORMS are
    # resubmitted with 1024, inside the timed region; likewise 16 returned changes per branch first,
    # 64 for the branches that had more.
    W, W2 = 96, 1024
    u = rng.random(nJobs * W)
    u2 = rng.random(32768 * W2)
    uOff = np.arange(nJobs + 1) * W
    pinned = mtt.lib().mtt_host_register(u.ctypes.data, u.nbytes) == 0
    mtt.lib().mtt_host_register(u2.ctypes.data, u2.nbytes)
    br3 = np.arange(nBr).reshape(nJobs, 3)

    CAP, CAP2 = 16, 64       # changes returned per branch: first pass / resubmission
    # inputs as the C-ABI wants them, once; page-locked outputs reused by every step
    brOff = np.ascontiguousarray(brOff, np.int64)
    uOff = np.ascontiguousarray(uOff, np.int64)
    a, b = a.astype(np.uint8), b.astype(np.uint8)
    model = np.ascontiguousarray(model, np.int32)
    for arr in (t0_, t1_, a, b, model, brOff, uOff):
        mtt.lib().mtt_host_register(arr.ctypes.data, arr.nbytes)
    outbuf = um.retype_buffers(nJobs, nBr, CAP, pin=True)

    t_call = [0.0]

    def step():
        tc = time.perf_counter()
        r = um.retype(brOff, t0_, t1_, a, b, u, uOff, outCap=CAP, model=model, out=outbuf)
        t_call[0] += time.perf_counter() - tc
        bad = r["status"] == 5
        bad |= r["nChg"] > CAP
        short = np.unique(np.flatnonzero(bad) // 3)
        if len(short):
            if len(short) > 32768:
                raise RuntimeError("too many short windows: %d" % len(short))
            idx = br3[short].ravel()
            r2 = um.retype(np.arange(len(short) + 1) * 3, t0_[idx], t1_[idx], a[idx], b[idx],
                           u2[:len(short) * W2], np.arange(len(short) + 1) * W2, outCap=CAP2,
                           model=model[idx])
            r["status"][idx] = r2["status"]
        return r, len(short)

    for _ in range(args.warmup):
        r, nShort = step()
    barrier(world)
    l0 = ctx.launch_count()
    t_call[0] = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r, nShort = step()
    barrier(world)
    dt = max_over_ranks(time.perf_counter() - t0, world)
    if rank != 0:
        return
    print(json.dumps({
        "metric": "cfg5 uniformization branch retypes/s (end to end, host buffers)",
        "value": nBr * world * args.steps / dt, "unit": "branches/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "cfg5: %d chains x %d NodeRetype-shaped jobs (3 branches), d=10"
                               % (chains * world, jobs_per_chain), "uniform_window": W,
                   "resubmitted_jobs_per_step": int(nShort), "resubmit_window": W2,
                   "pinned": bool(pinned)},
        "main_call_ms_per_step": t_call[0] / args.steps * 1e3,
        "ok_fraction": float((r["status"] == 0).mean()),
        "gpu_launches": int(ctx.launch_count() - l0)}))


def bench_untyped(args):
    """untyped-tree particle density, trees x particles in one call
    (--workload untyped_sample: the same particles returned as typed trees)."""
    import multitypetree_b200 as mtt
    from multitypetree_b200.uniformization import UniformizationModels
    from multitypetree_b200.untyped import untyped_logp_batch, untyped_sample_batch
    as_trees = args.workload == "untyped_sample"
    if as_trees:
        def untyped_logp_batch(um, N, trees, P, orders, u, uOff, ctx=None):      # noqa: F811
            return untyped_sample_batch(um, N, trees, P, orders, u, uOff, chgCap=256, ctx=ctx)
    world, rank, local = dist_setup(args.gpus)
    d, n = 2, (args.tips or 100)
    B = args.batch if args.batch != (1 << 19) else 64
    P, W = 256, 1024
    m, N = mtt.synthetic_model(d)
    tb = mtt.simulate_trees(B, n, d, m, N, seed=6000 + rank * B)
    trees = [tb.tree(b) for b in range(B)]
    rng = np.random.default_rng(6000 + rank)
    orders = np.concatenate([rng.permutation(n).astype(np.int32) for _ in range(B * P)])
    u = rng.random(B * P * W)
    uOff = np.arange(B * P + 1, dtype=np.int64) * W
    ctx = mtt.Context(local)
    um = UniformizationModels(m, ctx=ctx)
    L = mtt.lib()
    pinned = [a for a in (u, orders) if L.mtt_host_register(a.ctypes.data, a.nbytes) == 0]
    for _ in range(args.warmup):
        r = untyped_logp_batch(um, N, trees, P, orders, u, uOff, ctx=ctx)
    barrier(world)
    l0 = ctx.launch_count()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = untyped_logp_batch(um, N, trees, P, orders, u, uOff, ctx=ctx)
    barrier(world)
    dt = max_over_ranks(time.perf_counter() - t0, world)
    if rank != 0:
        return
    # CPU baseline: the oracle restatement, one thread, a bounded sample of the same particles
    from oracle import oracle as orc
    o = orc.UnifModel(m)
    sample = min(B * P, 2048)
    c0 = time.perf_counter()
    worst = 0.0
    for q in range(sample):
        t = trees[q // P]
        if as_trees:
            ro = orc.untyped_particle_tree(o, N, t.parent, t.height, t.nodeType,
                                           orders[q * n:(q + 1) * n], u[uOff[q]:uOff[q + 1]])
            b, p_, c = q // P, q % P, ro["nChg"]
            same = (r["nChg"][b, p_] == c and c <= 256
                    and np.array_equal(r["nodeType"][b][p_], ro["nodeType"].astype(np.uint8))
                    and np.array_equal(r["chgNode"][b, p_, :c], ro["chgNode"])
                    and np.all(np.abs(r["chgTime"][b, p_, :c] - ro["chgTime"])
                               <= 4 * np.spacing(ro["chgTime"])))
            if not same:
                raise SystemExit("typed tree of particle %d differs from the oracle's" % q)
        else:
            ro = orc.untyped_particle(o, N, t.parent, t.height, t.nodeType,
                                      orders[q * n:(q + 1) * n], u[uOff[q]:uOff[q + 1]])
        worst = max(worst, abs(ro["logW"] - r["logW"].ravel()[q]) / max(1.0, abs(ro["logW"])))
    cdt = time.perf_counter() - c0
    print(json.dumps({
        "metric": ("typed trees sampled from untyped ones: typings/s (end to end, host buffers; "
                   "identical to the oracle on the sample, times to the last place)" if as_trees else
                   "untyped-tree density: particle evaluations/s (end to end, host buffers)"),
        "value": B * P * world * args.steps / dt, "unit": "particles/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "%s: %d trees x %d particles per GPU, %d tips, d=%d"
                               % ("typed trees from untyped ones (room for 256 changes each)"
                                  if as_trees else "untyped density", B, P, n, d),
                   "tree_evals_per_s": B * world * args.steps / dt,
                   "uniform_window": W, "pinned": len(pinned) == 2},
        "cpu_baseline": {"value": sample / cdt, "unit": "particles/s", "cores": 1, "kind": "port",
                         "sample": "%d particles of the same batch, oracle/orc_untyped.c" % sample},
        "parity_max_rel_err_vs_oracle": worst,
        "gpu_launches": int(ctx.launch_count() - l0)}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="density", choices=["density", "peel", "peel_chains", "peel_sharded", "retype", "untyped", "untyped_sample"])
    ap.add_argument("--tips", type=int, default=0)
    ap.add_argument("--sites", type=int, default=0)
    ap.add_argument("--batch", type=int, default=1 << 19, help="trees per GPU per step")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return bench_reference(args)
    if args.workload == "peel":
        return bench_peel(args)
    if args.workload == "peel_chains":
        return bench_peel_chains(args)
    if args.workload == "peel_sharded":
        return bench_peel_sharded(args)
    if args.workload == "retype":
        return bench_retype(args)
    if args.workload in ("untyped", "untyped_sample"):
        return bench_untyped(args)
    return bench_density(args)


if __name__ == "__main__":
    main()
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass
