#!/usr/bin/env python3
"""bench.py -- tree log-likelihood evaluations/sec on B200 (BASELINE.json metric).

A "step" is one full tree log-likelihood evaluation (all nodes dirty): the BEAGLE call sequence
BeagleDataLikelihoodDelegate.calculateLikelihood issues (BDLD:812-937) --
setEigenDecomposition, setCategoryRates/Weights, setStateFrequencies, updateTransitionMatrices
(2N-2 branches), updatePartials (N-1 operations), calculateRootLogLikelihoods -- through the C ABI
of libhmsbeagle.so, with BEAST's double-buffer index flipping between steps.

  value : steps enqueued back to back on the instance stream, tip data / partials resident in HBM,
          result left on the device (one D2H at the end); CUDA-event timed, max over ranks.
  e2e   : the same sequence through the synchronous reference-facing calls with HOST buffers:
          every step uploads the eigen system, rates, frequencies, branch lengths and op list and
          lands the 8-byte log-likelihood on the host (and, for N > 1, sums it across ranks).
  N > 1 : weak scaling -- every rank owns one 10,000-pattern shard of an (N x 10,000)-pattern
          alignment (BEAST's -beagle_instances pattern split, Patterns.java:142-169) and the
          per-shard log-likelihoods are summed with ONE NCCL all-reduce of a single double.
          value = shard evaluations/sec over all ranks = N x joint evaluations/sec.
  --impl reference : the CPU restatement of the reference path (oracle/beagle_cpu.c; the real
          BEAGLE-CPU is un-vendored and cannot be built here) on all host cores, rank 0 only.
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

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import beast_mcmc_b200  # noqa: E402,F401
from harness import evomodel as em  # noqa: E402
from harness import treedatalikelihood as tdl  # noqa: E402

WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on at N=1
    "gtr_g4_1000x10k": dict(taxa=1000, patterns=10000, states=4, categories=4, rootHeight=0.1, treeSeed=20240924),
    # the same with a scale buffer written by every op ("also run with scaleWrite on every op")
    "gtr_g4_1000x10k_rescaled": dict(taxa=1000, patterns=10000, states=4, categories=4, rootHeight=0.1, treeSeed=20240924,
                                     scaling=True, data="gtr_g4_1000x10k"),
    # configs[2]: codon model on the dense-contraction path
    "codon_mg94_500x5k": dict(taxa=500, patterns=5000, states=61, categories=1, rootHeight=0.1, treeSeed=2),
    # amino-acid shape (20 states, G4): the tensor path's NT=3 instance, on the memory side of the roofline
    "aa20_g4_500x5k": dict(taxa=500, patterns=5000, states=20, categories=4, rootHeight=0.3, treeSeed=4),
    "codon_mg94_500x5k_g4": dict(taxa=500, patterns=5000, states=61, categories=4, rootHeight=0.1, treeSeed=2),
    # configs[0]-like latency case (benchmark1.xml shape: 1441 taxa, 593 patterns, HKY, no gamma)
    "hky_1441x593": dict(taxa=1441, patterns=593, states=4, categories=1, rootHeight=0.1, treeSeed=1441),
    # configs[0] as shipped: the reference's own benchmark alignments (tests/golden/benchmark{1,2}_patterns.npz, extracted
    # from examples/Benchmarks/benchmark{1,2}.xml), seeded coalescent start tree as the XMLs draw a random one
    "benchmark1_xml": dict(taxa=1441, patterns=593, states=4, categories=1, rootHeight=0.05, treeSeed=666, fixture="benchmark1",
                           scaling=True),      # underflows unscaled: evaluated the way BEAST does after its first underflow
    "benchmark2_xml": dict(taxa=62, patterns=5565, states=4, categories=4, rootHeight=0.3, treeSeed=666, fixture="benchmark2"),
    # configs[3]-like: Makona-shaped synthetic (data absent from the reference tree)
    "makona_like_1610x6k": dict(taxa=1610, patterns=6000, states=4, categories=4, rootHeight=0.0025, treeSeed=3),
}


# ------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------
def build_workload(name, shard_index, overrides):
    w = dict(WORKLOADS[name])
    w.update({k: v for k, v in overrides.items() if v is not None})
    tree = em.Tree.coalescent(w["taxa"], w["rootHeight"], w["treeSeed"])
    if name.startswith("makona"):
        rng = np.random.default_rng(3)
        tree.branchRate = np.exp(rng.normal(0.0, 0.3, tree.nodeCount))     # relaxed clock folded into lengths
    if w["states"] == 4:
        if name.startswith("hky") or name == "benchmark1_xml":
            model = em.HKY(2.0, np.full(4, 0.25))                  # benchmark1.xml: <HKYModel> kappa 2, uniform frequencies
        elif name == "benchmark2_xml":
            model = em.GTR(1.0, 1.0, 1.0, 1.0, 1.0, 1.0, np.full(4, 0.25))   # benchmark2.xml:702-725 start values
        else:
            model = em.GTR(1.0, 4.0, 0.7, 1.2, 5.0, 1.0, np.array([0.30, 0.22, 0.24, 0.24]))
    elif w["states"] == 61:
        model = em.MG94HKYCodonModel(1.0, 0.3, 2.0)
    else:
        rng = np.random.default_rng(7)
        S = w["states"]
        model = em.SubstitutionModel(rng.uniform(0.2, 3.0, S * (S - 1) // 2), rng.dirichlet(np.full(S, 5.0)))
    site = em.GammaSiteRateModel(shape=0.5, gammaCategoryCount=w["categories"]) if w["categories"] > 1 \
        else em.GammaSiteRateModel()
    # the simulated alignment is cached per box (sweeps re-use it); it is regenerated when absent
    cache = os.path.join(os.environ.get("B200_BENCH_CACHE", "/tmp/b200_bench_cache"),
                         f"{w.get('data', name)}_{w['taxa']}_{w['patterns']}_{w['states']}_{w['categories']}_{shard_index}.npz")
    if w.get("fixture"):
        z = np.load(os.path.join(ROOT, "tests", "golden", w["fixture"] + "_patterns.npz"))
        pats = em.Patterns(z["states"].astype(np.int32), z["weights"], 4)
        assert pats.taxonCount == w["taxa"] and pats.patternCount == w["patterns"]
    elif os.path.exists(cache):
        z = np.load(cache)
        pats = em.Patterns(z["states"], z["weights"], w["states"])
    else:
        pats = em.synthetic_patterns(tree, model, site, w["patterns"], seed=1 + 1000003 * shard_index)
        try:
            os.makedirs(os.path.dirname(cache), exist_ok=True)
            np.savez(cache + f".{os.getpid()}.tmp.npz", states=pats.states, weights=pats.weights)
            os.replace(cache + f".{os.getpid()}.tmp.npz", cache)
        except OSError:
            pass
    return w, tree, pats, model, site


class Evaluation:
    """Pre-built call arguments of one full evaluation (what the Java side hands to JNI), in the two
    buffer-index parities BEAST's BufferIndexHelper alternates between."""

    def __init__(self, tree, pats, model, site, traversal, scaling=False):
        self.tree, self.pats, self.model, self.site = tree, pats, model, site
        self.scaling = scaling
        N, n = tree.tipCount, tree.nodeCount
        self.N, self.n = N, n
        like = tdl.TreeDataLikelihood.__new__(tdl.TreeDataLikelihood)
        like.tree, like.traversalType, like.updateNode = tree, traversal, np.ones(n, dtype=bool)
        like._dispatch()
        self.branchNodes = np.array([b for b, _ in like.branchOperations], dtype=np.int32)
        self.lengths = np.array([t for _, t in like.branchOperations], dtype=np.float64)
        self.nodeOps = like.nodeOperations
        internal = n - N
        self.ops, self.probIdx, self.rootIdx, self.scaleIdx, self.cumIdx = [], [], [], [], []
        for parity in (0, 1):
            pidx = lambda k: k if k < N else k + parity * internal       # BufferIndexHelper.getOffsetIndex
            midx = lambda k: k + parity * n
            sidx = lambda k: (k - N) + parity * (internal + 1)           # scale buffers: BDLD:203,626-628,868-881
            ops = np.empty(len(self.nodeOps) * 7, dtype=np.int32)
            for q, (node, c1, c2) in enumerate(self.nodeOps):
                ops[7 * q: 7 * q + 7] = (pidx(node), sidx(node) if scaling else -1, -1, pidx(c1), midx(c1), pidx(c2), midx(c2))
            self.scaleIdx.append(np.array([sidx(node) for node, _, _ in self.nodeOps], dtype=np.int32))
            self.cumIdx.append(internal + parity * (internal + 1))
            self.ops.append(ops)
            self.probIdx.append((self.branchNodes + parity * n).astype(np.int32))
            self.rootIdx.append(pidx(tree.root))
        self.eig = model.getEigenDecomposition()
        # op mix for the algorithmic byte/flop count
        self.mix = {"pp": 0, "sp": 0, "ss": 0}
        for _, c1, c2 in self.nodeOps:
            k = (c1 < N) + (c2 < N)
            self.mix["pp" if k == 0 else ("sp" if k == 1 else "ss")] += 1

    def algorithmic(self, S, C, P):
        """bytes and flops of ONE updatePartials launch (whole op list), SURVEY.md 8(d) figures."""
        pp, sp, ss = self.mix["pp"], self.mix["sp"], self.mix["ss"]
        byt = pp * (3 * C * P * S * 8 + 2 * C * S * S * 8) + sp * (2 * C * P * S * 8 + 4 * P + 2 * C * S * S * 8) + \
            ss * (C * P * S * 8 + 8 * P + 2 * C * S * S * 8)
        flo = pp * C * P * S * (4 * S + 1) + sp * C * P * S * (2 * S + 1) + ss * C * P * S
        return byt, flo

    def h2d_bytes(self, S, C):
        return int(self.ops[0].nbytes + self.probIdx[0].nbytes + self.lengths.nbytes +
                   (2 * S * S + S) * 8 + 2 * C * 8 + S * 8)


def create_instance(factory, ev, S, C, P, resource):
    N, n = ev.N, ev.n
    inst = factory(N, 2 * (n - N) + N, N, S, P, 2, 2 * n, C, 2 * (n - N + 1), resource, 0, 0)
    for t in range(N):
        inst.setTipStates(t, np.ascontiguousarray(ev.pats.states[t], dtype=np.int32))
    inst.setPatternWeights(np.ascontiguousarray(ev.pats.weights))
    return inst


def issue_sync(inst, ev, parity, out):
    """One evaluation through the synchronous reference-facing calls, host buffers in / host double out."""
    inst.setEigenDecomposition(parity, ev.eig.Evec, ev.eig.Ievc, ev.eig.Eval)
    inst.setCategoryRates(ev.site.getCategoryRates())
    inst.setCategoryWeights(0, ev.site.getCategoryProportions())
    inst.setStateFrequencies(0, ev.model.getFrequencies())
    inst.updateTransitionMatrices(parity, ev.probIdx[parity], None, None, ev.lengths, len(ev.lengths))
    inst.updatePartials(ev.ops[parity], len(ev.nodeOps), -1)
    cum = MINUS1
    if ev.scaling:             # BDLD:915-926
        inst.resetScaleFactors(ev.cumIdx[parity])
        inst.accumulateScaleFactors(ev.scaleIdx[parity], len(ev.nodeOps), ev.cumIdx[parity])
        cum = np.array([ev.cumIdx[parity]], dtype=np.int32)
    inst.calculateRootLogLikelihoods(np.array([ev.rootIdx[parity]], dtype=np.int32), ZERO, ZERO, cum, 1, out)
    return out[0]


ZERO = np.zeros(1, dtype=np.int32)
MINUS1 = np.full(1, -1, dtype=np.int32)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi in loop mode (-lms) for the duration of the timed regions (B200_PROFILING.md clocks line)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown," \
        "clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu_index, self.proc = gpu_index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu_index), "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        rows = []
        if self.proc is not None:
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
                out, _ = self.proc.communicate()
            rows = [[c.strip() for c in line.split(",")] for line in out.splitlines() if line.count(",") >= 8]
        sm = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        pw = [float(r[3]) for r in rows if r[3].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(rows)}


# ------------------------------------------------------------------------------------------------
# CPU arm (oracle port): cpu_baseline and --impl reference
# ------------------------------------------------------------------------------------------------
def cpu_pick_threads(ev, S, C, P, cores):
    """The port's pthread pool does not scale to every host (128-way barriers on a shared box): try a few
    thread counts on two evaluations each and keep the fastest -- the baseline gets its best configuration."""
    from oracle import cpu
    best, best_t = None, cores
    tried = []
    for th in sorted({cores, max(1, cores // 2), max(1, cores // 4), min(cores, 32), min(cores, 16), min(cores, 8)}, reverse=True):
        inst = create_instance(cpu.factory(threads=th), ev, S, C, P, None)
        out = np.zeros(1)
        issue_sync(inst, ev, 0, out)
        t0 = time.perf_counter()
        issue_sync(inst, ev, 1, out)
        issue_sync(inst, ev, 0, out)
        dt = (time.perf_counter() - t0) / 2
        inst.finalize()
        tried.append((th, dt))
        if best is None or dt < best:
            best, best_t = dt, th
    return best_t, tried


def cpu_time_evaluations(ev, S, C, P, threads, min_evals, budget_s):
    from oracle import cpu                      # checker / baseline only, never the product path
    inst = create_instance(cpu.factory(threads=threads), ev, S, C, P, None)
    out = np.zeros(1)
    issue_sync(inst, ev, 0, out)                # warm-up (allocations)
    times, val = [], 0.0
    t_all = time.perf_counter()
    k = 0
    while k < min_evals or (time.perf_counter() - t_all < budget_s and k < 1000):
        t0 = time.perf_counter()
        val = issue_sync(inst, ev, (k + 1) & 1, out)
        times.append(time.perf_counter() - t0)
        k += 1
    inst.finalize()
    return times, val


def run_reference_arm(args, meta_base):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from beast_mcmc_b200 import build
    build.build_oracle()
    w, tree, pats, model, site = build_workload(args.workload, 0, vars(args))
    S, C, P = w["states"], site.getCategoryCount(), pats.patternCount
    ev = Evaluation(tree, pats, model, site, "POST_ORDER", scaling=bool(w.get("scaling")))
    cores = os.cpu_count() or 1
    from oracle import cpu
    threads, tried = cpu_pick_threads(ev, S, C, P, cores)
    inst = create_instance(cpu.factory(threads=threads), ev, S, C, P, None)
    out = np.zeros(1)
    for k in range(args.warmup):
        issue_sync(inst, ev, k & 1, out)
    t0 = time.perf_counter()
    for k in range(args.steps):
        issue_sync(inst, ev, k & 1, out)
    dt = time.perf_counter() - t0
    inst.finalize()
    value = args.steps / dt
    line = dict(meta_base)
    line.update({
        "impl": "reference", "value": value, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "vs_baseline": None, "logL": float(out[0]),
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": threads, "kind": "port", "host_cores": cores,
                         "sample": f"{args.steps} full evaluations of the same workload (oracle/beagle_cpu.c, "
                                   f"{threads} pthreads over pattern blocks = fastest of {[t for t, _ in tried]})"},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# main
# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="gtr_g4_1000x10k", choices=list(WORKLOADS))
    ap.add_argument("--taxa", type=int)
    ap.add_argument("--patterns", type=int)
    ap.add_argument("--categories", type=int)
    ap.add_argument("--states", type=int)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=12.0)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    w0 = dict(WORKLOADS[args.workload])
    w0.update({k: v for k, v in vars(args).items() if k in w0 and v is not None})
    meta_base = {
        "metric": "tree log-likelihood evals/sec", "unit": "evals/s", "higher_is_better": True,
        "scaling": "weak", "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "taxa": w0["taxa"], "patterns_per_gpu": w0["patterns"],
                   "states": w0["states"], "categories": w0["categories"],
                   "step": "full evaluation, all nodes dirty: eigen+rates+freqs upload, 2N-2 matrices, N-1 partials ops, root",
                   "sharding": f"{world} x {w0['patterns']}-pattern shards, one per GPU, NCCL sum of 1 double",
                   "l2": "inputs larger than L2: each step writes N-1 partials buffers (1.28 GB at the default "
                         "workload) into the alternate buffer parity"},
    }
    if args.impl == "reference":
        run_reference_arm(args, meta_base)
        return

    import torch
    import torch.distributed as dist
    from beast_mcmc_b200 import beagle

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    w, tree, pats, model, site = build_workload(args.workload, rank, vars(args))
    S, C, P = w["states"], site.getCategoryCount(), pats.patternCount
    scaling = bool(w.get("scaling"))
    ev = Evaluation(tree, pats, model, site, "REVERSE_LEVEL_ORDER", scaling=scaling)     # what BEAST sends a non-CPU instance
    inst = create_instance(beagle.BeagleFactory.loadBeagleInstance, ev, S, C, P, [local_rank + 1, 0])
    lib = beagle.load_library()

    import ctypes as Cc
    devp, strm = Cc.c_void_p(), Cc.c_void_p()
    out = np.zeros(1)
    logL = issue_sync(inst, ev, 0, out)
    rc = lib.b200RootLogLikelihoodDevice(inst.instance, ev.rootIdx[0], 0, 0, ev.cumIdx[0] if scaling else -1,
                                         Cc.byref(devp), Cc.byref(strm))
    assert rc == 0
    stream = torch.cuda.ExternalStream(strm.value, device=torch.device("cuda", local_rank))

    class _Dev:      # zero-copy torch view of the engine's device-resident result
        __cuda_array_interface__ = {"shape": (1,), "typestr": "<f8", "data": (devp.value, False), "version": 3}
    dres = torch.as_tensor(_Dev(), device=torch.device("cuda", local_rank))

    def reduce_async():
        if world > 1:
            with torch.cuda.stream(stream):
                dist.all_reduce(dres, op=dist.ReduceOp.SUM)

    def step_async(k):
        p = k & 1
        # eigen system, rates and frequencies are resident (slot 0); buffers flip like BEAST's do
        inst.updateTransitionMatrices(0, ev.probIdx[p], None, None, ev.lengths, len(ev.lengths))
        inst.updatePartials(ev.ops[p], len(ev.nodeOps), -1)
        cum = -1
        if scaling:
            inst.resetScaleFactors(ev.cumIdx[p])
            inst.accumulateScaleFactors(ev.scaleIdx[p], len(ev.nodeOps), ev.cumIdx[p])
            cum = ev.cumIdx[p]
        lib.b200RootLogLikelihoodDevice(inst.instance, ev.rootIdx[p], 0, 0, cum, None, None)
        reduce_async()

    def step_e2e(k):
        p = k & 1
        inst.setEigenDecomposition(p, ev.eig.Evec, ev.eig.Ievc, ev.eig.Eval)
        inst.setCategoryRates(ev.site.getCategoryRates())
        inst.setCategoryWeights(0, ev.site.getCategoryProportions())
        inst.setStateFrequencies(0, ev.model.getFrequencies())
        inst.updateTransitionMatrices(p, ev.probIdx[p], None, None, ev.lengths, len(ev.lengths))
        inst.updatePartials(ev.ops[p], len(ev.nodeOps), -1)
        cum = -1
        if scaling:
            inst.resetScaleFactors(ev.cumIdx[p])
            inst.accumulateScaleFactors(ev.scaleIdx[p], len(ev.nodeOps), ev.cumIdx[p])
            cum = ev.cumIdx[p]
        if world == 1:
            inst.calculateRootLogLikelihoods(np.array([ev.rootIdx[p]], dtype=np.int32), ZERO, ZERO,
                                             np.array([cum], dtype=np.int32), 1, out)
            return out[0]
        lib.b200RootLogLikelihoodDevice(inst.instance, ev.rootIdx[p], 0, 0, cum, None, None)
        reduce_async()
        with torch.cuda.stream(stream):
            return float(dres.item())          # 8-byte D2H of the joint log-likelihood

    def bracket():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput ("value") + live kernel timing for the roofline ----------------
    for k in range(args.warmup):
        step_async(k)
    bracket()
    inst.setKernelTiming(True)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    bracket()
    t0 = time.perf_counter()
    e0.record(stream)
    for k in range(args.steps):
        step_async(k)
    e1.record(stream)
    bracket()
    wall = time.perf_counter() - t0
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    wall = max_over_ranks(wall)
    k_ms, k_n = inst.getKernelTiming(0)
    m_ms, m_n = inst.getKernelTiming(1)
    r_ms, r_n = inst.getKernelTiming(2)
    inst.setKernelTiming(False)
    joint = float(dres.cpu()[0])

    # ---- end to end through the synchronous public calls ("e2e") ------------------------------------
    for k in range(3):
        step_e2e(k)
    bracket()
    t0 = time.perf_counter()
    per_call = []
    for k in range(args.steps):
        tc = time.perf_counter()
        last = step_e2e(k)
        per_call.append(time.perf_counter() - tc)
    bracket()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    per_call.sort()
    e2e_dist = {q: 1e3 * per_call[min(len(per_call) - 1, int(f * len(per_call)))] for q, f in
                (("p10_ms", 0.10), ("median_ms", 0.50), ("p90_ms", 0.90))}
    clocks = sampler.stop() if sampler else None

    # ---- secondary: the incremental evaluation MCMC mostly issues (one tip-to-root path dirty) ------------
    inc = None
    if world == 1 and not scaling:
        issue_sync(inst, ev, 0, out)                       # parity-0 buffers hold the current state
        rng = np.random.default_rng(5)
        N, n, internal = ev.N, ev.n, ev.n - ev.N
        paths = []
        for _ in range(64):
            node, path = int(rng.integers(0, N)), []
            while tree.parent[node] >= 0:
                par = int(tree.parent[node])
                sib = int(tree.child[par][0]) if int(tree.child[par][1]) == node else int(tree.child[par][1])
                path.append((par, node, sib))
                node = par
            ops = np.empty(7 * len(path), dtype=np.int32)
            for q, (par, child, sib) in enumerate(path):
                cidx = child if (child < N or q == 0) else child + internal        # freshly written -> parity 1
                ops[7 * q: 7 * q + 7] = (par + internal, -1, -1, cidx, child + (n if q == 0 else 0), sib, sib)
            first = path[0][1]
            paths.append((ops, len(path), np.array([first + n], dtype=np.int32),
                          np.array([tree.branchLength(first) * 1.01]), np.array([tree.root + internal], dtype=np.int32)))
        def run_incremental(target, rounds):
            for ops, cnt, pidx, blen, rootIdx in paths[:8]:
                target.updateTransitionMatrices(0, pidx, None, None, blen, 1)
                target.updatePartials(ops, cnt, -1)
                target.calculateRootLogLikelihoods(rootIdx, ZERO, ZERO, MINUS1, 1, out)
            t0 = time.perf_counter()
            reps = 0
            for _ in range(rounds):
                for ops, cnt, pidx, blen, rootIdx in paths:
                    target.updateTransitionMatrices(0, pidx, None, None, blen, 1)
                    target.updatePartials(ops, cnt, -1)
                    target.calculateRootLogLikelihoods(rootIdx, ZERO, ZERO, MINUS1, 1, out)
                    reps += 1
            return reps, time.perf_counter() - t0

        reps, dt = run_incremental(inst, max(1, min(args.steps, 2000) // 64 + 1))
        inc = {"evals_per_s": reps / dt, "us_per_eval": 1e6 * dt / reps,
               "mean_ops_per_eval": float(np.mean([c for _, c, _, _, _ in paths])),
               "what": "one branch length changed: 1 matrix, tip-to-root path of partials ops, root; host buffers, synchronous"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = world * args.steps / (dev_ms * 1e-3)
    byt, flo = ev.algorithmic(S, C, P)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (copy, burst)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    # updatePartials is one launch per phase of independent subtrees; the roofline unit is the whole
    # operation list (all its launches) of one step
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath)).get(args.workload)
        if tj and all(getattr(args, k) is None for k in ("taxa", "patterns", "categories", "states")):
            traffic = tj["dram_bytes_read_per_step"] + tj["dram_bytes_write_per_step"]
            traffic_src = tj["source"]
    k_avg_ms = k_ms / args.steps
    achieved = byt / (k_avg_ms * 1e-3) / 1e9
    line = dict(meta_base)
    line.update({
        "value": value, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "wall_ms_per_step": 1e3 * wall / args.steps,
        "vs_baseline": None, "logL": joint, "joint_evals_per_s": args.steps / (dev_ms * 1e-3),
        "roofline": {"bound": "hbm", "kernel": "k_walk4 (updatePartials, whole op list per launch)"
                     if S <= 4 else "k_walk_mma (updatePartials on the fp64 tensor pipe; math-bound for S=61: see gflops)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     # DRAM-side view of the same launches: measured bytes (ncu) over the live time.  frac counts ALGORITHMIC
                     # bytes, of which children re-read from L2 never reach HBM -- it can exceed 1; dram_frac cannot.
                     "dram_achieved": (traffic / (k_avg_ms * 1e-3) / 1e9) if traffic else None,
                     "dram_frac": (traffic / (k_avg_ms * 1e-3) / 1e9 / peak) if traffic else None,
                     "algorithmic_bytes_per_step": byt, "algorithmic_flops_per_step": flo,
                     "gflops": flo / (k_avg_ms * 1e-3) / 1e9, "partials_ms_per_step": k_avg_ms,
                     "launches_per_step": k_n / args.steps,
                     "op_mix": ev.mix, "share_of_step": k_ms / dev_ms,
                     "other_kernels_ms_per_step": {"transition_matrices": m_ms / args.steps,
                                                   "root": r_ms / args.steps}},
        "e2e": {"value": world * args.steps / e2e_s, "unit": "evals/s", "ms_per_step": 1e3 * e2e_s / args.steps,
                "h2d_bytes_per_step": ev.h2d_bytes(S, C), "d2h_bytes_per_step": 8, "logL": float(last),
                "per_call_rank0": e2e_dist},
        "gpu_launches": int(k_n + m_n + r_n),
        "clocks": clocks,
        "incremental": inc,
    })
    if not args.no_cpu_baseline:
        from beast_mcmc_b200 import build
        build.build_oracle()
        cores = os.cpu_count() or 1
        evc = Evaluation(tree, pats, model, site, "POST_ORDER", scaling=scaling)
        threads, tried = cpu_pick_threads(evc, S, C, P, cores)
        times, cval = cpu_time_evaluations(evc, S, C, P, threads, 3, args.cpu_budget)
        line["cpu_baseline"] = {"value": 1.0 / statistics.median(times), "unit": "evals/s", "cores": threads,
                                "host_cores": cores, "kind": "port",
                                "sample": f"{len(times)} full evaluations of this rank-0 shard (median), "
                                          f"oracle/beagle_cpu.c with {threads} pthreads (fastest of "
                                          f"{[t for t, _ in tried]})",
                                "logL": cval, "rel_diff_vs_gpu": abs(cval - logL) / abs(cval)}
        t1, _ = cpu_time_evaluations(evc, S, C, P, 1, 1, 0.0)
        line["cpu_baseline"]["single_thread"] = 1.0 / statistics.median(t1)
        if inc is not None:
            from oracle import cpu
            cinst = create_instance(cpu.factory(threads=threads), evc, S, C, P, None)
            issue_sync(cinst, evc, 0, out)
            creps, cdt = run_incremental(cinst, 2)
            cinst.finalize()
            inc["cpu_port_evals_per_s"] = creps / cdt
    print(json.dumps(line), flush=True)
    inst.finalize()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
