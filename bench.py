#!/usr/bin/env python3
"""bench.py -- the reference's headline metric on B200 (BASELINE.json).

  python bench.py --gpus N --steps K --warmup W              (ours; torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K --warmup W   (CPU reference arm)

Workload (config.workload = "eri_class_sweep", BASELINE.json configs[1]): one step = one pass
over the 22 canonical classes (ss|ss)..(dd|dd) (la>=lb, lc>=ld, la+lb<=lc+ld, l<=2), 10^7 random
primitive shell quartets per class, every Cartesian integral materialised in HBM.
`value` = shell quartets per second with tasks and outputs resident in HBM; `e2e` = the same
call with HOST buffers (tasks from pinned host memory, integrals copied back).  Beside it the
`fock` object times the direct Fock build that consumes the integrals (configs[2],
(H2O)_64 / def2-TZVP, Schwarz-screened; at N > 1 the quartets are sharded and the partial G's
all-reduced over NCCL), which is the "Fock-build s at 1/2/4/8 B200" half of the metric.
Multi-GPU: the sweep is weak-scaled (each rank processes its own 10^7 quartets per class, no
collective: quartets are independent); the Fock build is strong-scaled.

The CPU oracle (oracle/) is executed only by the cpu_baseline leg and by --impl reference.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "eri_shell_quartets_per_s"
UNIT = "shell quartets/s"


def nc(l):
    return (l + 1) * (l + 2) // 2


def sweep_classes(lmax=2):
    from libint_b200.flops import canonical_classes
    return canonical_classes(lmax)


def class_table(cl, npairs, seed):
    """4*npairs primitive shells: group g holds the shells of index position g of the class.
    Distribution: centres U(-2,2)^3 bohr, exponents 10^U(-1,1.5), unit coefficients --
    T = rho*|PQ|^2 spans the Boys interpolation table and the asymptotic branch."""
    rng = np.random.default_rng(20240607 + seed)
    n = 4 * npairs
    l = np.repeat(np.array(cl, dtype=np.int32), npairs)
    O = rng.uniform(-2.0, 2.0, (n, 3))
    al = 10.0 ** rng.uniform(-1.0, 1.5, n)
    co = np.ones(n)
    return l, np.zeros(n, dtype=np.int32), np.ones(n, dtype=np.int32), O, al, co


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms.  Started BEFORE the warm-up steps (nvidia-smi
    needs several hundred ms to deliver its first sample; a three-step timed region is shorter than that), the
    samples are then cut to the timed window by their timestamps; if the window is too short to hold one, the
    samples of the warm-up steps -- the same kernels back to back -- are reported and `window` says so."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.t0 = self.t1 = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "50"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        time.sleep(0.12)   # let the sample that covers the end of the window arrive
        self.p.terminate()
        try:
            self.p.wait(5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        rows = []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().splitlines():
            t = [x.strip() for x in ln.split(",")]
            if len(t) < 9:
                continue
            try:
                ts = datetime.datetime.strptime(t[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(t[1]), float(t[2]), float(t[3]),
                             [nm for nm, v in zip(names, t[5:9]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        os.unlink(self.f.name)
        window = "timed region"
        sel = [r for r in rows if self.t0 is not None and self.t1 is not None and self.t0 <= r[0] <= self.t1 + 0.06]
        if not sel:   # region shorter than the sampling latency: the warm-up steps ran the same kernels
            sel = [r for r in rows if self.t1 is None or r[0] <= self.t1 + 0.06]
            window = "warm-up + timed region (timed region shorter than one sampling interval)"
        if sel:
            # under load = the GPU is drawing power for the kernels; idle samples before the first launch would
            # drag the median to the idle clock
            busy = [r for r in sel if r[3] >= 0.5 * max(x[3] for x in sel)] or sel
            out = {"sm_mhz": float(np.median([r[1] for r in busy])), "sm_max_mhz": float(max(r[2] for r in sel)),
                   "reasons": sorted({nm for r in sel for nm in r[4]}), "power_w_max": max(r[3] for r in sel),
                   "samples": len(busy), "window": window}
        return out


# --------------------------------------------------------------------------------------
# CPU reference legs (oracle = reference Engine on restated kernels; see oracle/)
# --------------------------------------------------------------------------------------
def cpu_sweep(classes, npairs, budget_s, nthreads, fast=True, use_pairs=True):
    """times the reference Engine (one per thread, round-robin) on a bounded sample of every
    class; returns the equal-count-mix throughput, per-class rates and the sample text.
    fast / use_pairs: the -O3 -march=x86-64-v3 build of the oracle and precomputed ShellPairs handed to
    compute2 (as hartree-fock++.cc:1697 does) -- the honest baseline; False/False reproduces round 1's
    (-O2 x86-64-v2, ShellPair::init inside every compute2 call)."""
    from oracle import pyoracle as po
    from libint_b200.flops import quartet_flops
    per = {}
    tot_q, tot_t = 0, 0.0
    tbudget = budget_s / len(classes)
    for ci, cl in enumerate(classes):
        tab = class_table(cl, npairs, ci)
        sh = po.Shells(*tab, raw=False)
        rng = np.random.default_rng(ci)
        # size the sample from the flop model (~1.5 GFLOP/s/thread guess), then time it
        nq = int(min(2_000_000, max(2000, tbudget * nthreads * 1.0e9 / quartet_flops(*cl))))
        b = rng.integers(0, npairs, nq)
        k = rng.integers(0, npairs, nq)
        q4 = np.stack([b, npairs + b, 2 * npairs + k, 3 * npairs + k], axis=1).astype(np.int32)
        t, _ = po.time_quartets(sh, q4, nthreads, use_pairs=use_pairs, fast=fast)
        per["".join(map(str, cl))] = nq / t
        tot_q += nq
        tot_t += t
    # equal count per class, as the GPU step: time for one quartet of each class
    mix = len(classes) / sum(1.0 / r for r in per.values())
    return mix, per, ("equal-count mix over %d classes, %d quartets timed in %.1f s; oracle build %s, ShellPairs %s"
                      % (len(classes), tot_q, tot_t,
                         "-O3 -march=x86-64-v3" if (fast and po.fast_available()) else "-O2 -march=x86-64-v2",
                         "precomputed" if use_pairs else "rebuilt per quartet"))


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    ncores = os.cpu_count() or 1
    classes = sweep_classes()
    vals = []
    for _ in range(args.warmup):
        cpu_sweep(classes, args.npairs, 2.0, ncores)
    # a step = one bounded sample of the sweep; the whole run stays within ~2 minutes of CPU work
    per_step = min(args.cpu_seconds, 120.0 / max(1, args.steps))
    t0 = time.time()
    for _ in range(args.steps):
        mix, per, sample = cpu_sweep(classes, args.npairs, per_step, ncores)
        vals.append(mix)
    wall = time.time() - t0
    v = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(1, args.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "eri_class_sweep", "classes": len(classes), "lmax": 2,
                       "quartets_per_class": args.quartets, "contraction": 1},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": ncores, "kind": "port",
                             "sample": sample + " (reference libint2::Engine compiled from its own headers "
                             "on the restated build_eri kernels; the generated library cannot be built here)"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "per_class_quartets_per_s": per}
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--quartets", type=int, default=10_000_000, help="quartets per class per step")
    ap.add_argument("--npairs", type=int, default=4096, help="bra / ket shell pairs per class")
    ap.add_argument("--chunk", type=int, default=1 << 20, help="quartets per launch")
    ap.add_argument("--e2e-quartets", type=int, default=10_000_000,
                    help="quartets per class in the host-buffer leg (default: the full configs[1] workload)")
    ap.add_argument("--cpu-seconds", type=float, default=20.0)
    ap.add_argument("--parity-quartets", type=int, default=20000, help="quartets per class checked against the arbiter")
    ap.add_argument("--no-fock", action="store_true")
    ap.add_argument("--fock-waters", default="4,4,4")
    ap.add_argument("--fock-basis", default="def2-tzvp")
    ap.add_argument("--fock-precision", type=float, default=1e-10)
    ap.add_argument("--fock256", choices=["auto", "on", "off"], default="auto",
                    help="configs[4], (H2O)_256 / cc-pVTZ: auto = when running on >= 2 GPUs")
    ap.add_argument("--no-grad", action="store_true")
    ap.add_argument("--grad-waters", default="3,3,3")
    ap.add_argument("--grad-basis", default="cc-pvdz")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-df3c", action="store_true")
    ap.add_argument("--df3c-carbons", type=int, default=40)
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    from libint_b200 import capi
    from libint_b200.flops import quartet_flops
    from libint_b200.fock import allreduce_sum_, init_distributed
    rank, world, local = init_distributed()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    ctx = capi.Context(local)
    stream = torch.cuda.Stream(dev)
    ctx.set_stream(stream.cuda_stream)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    # FP64 roofline denominator, measured here (MEASURED_PEAKS.json carries HBM and bf16 only)
    fp64_peak = max(capi.fp64_peak_probe(ctx, 4096)[0] for _ in range(3))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))

    classes = sweep_classes()
    nq, chunk = args.quartets, min(args.chunk, args.quartets)
    work = []
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    max_blk = 0
    for ci, cl in enumerate(classes):
        tab = class_table(cl, args.npairs, ci)
        bs = capi.Basis(ctx, *tab)
        i = np.arange(args.npairs, dtype=np.int32)
        bra = capi.Pairs(ctx, bs, bs, i, args.npairs + i)
        ket = capi.Pairs(ctx, bs, bs, 2 * args.npairs + i, 3 * args.npairs + i)
        tasks = torch.randint(0, args.npairs, (nq, 2), dtype=torch.int32, device=dev, generator=gen)
        blk = capi.eri_block_size(bra, ket)
        max_blk = max(max_blk, blk)
        work.append({"cl": cl, "bs": bs, "bra": bra, "ket": ket, "tasks": tasks, "blk": blk,
                     "flops": quartet_flops(*cl), "tab": tab})
    out = torch.empty(chunk * max_blk, dtype=torch.float64, device=dev)  # > L2 for every class but the smallest
    nchunks = (nq + chunk - 1) // chunk

    def sweep_step(events=None):
        with torch.cuda.stream(stream):
            for w in work:
                if events is not None:
                    e0 = torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                for c in range(nchunks):
                    t = w["tasks"][c * chunk:(c + 1) * chunk]
                    capi.eri_batch(ctx, w["bra"], w["ket"], t, out=out)
                if events is not None:
                    e1 = torch.cuda.Event(enable_timing=True)
                    e1.record(stream)
                    events.append((e0, e1))

    # timing rule: at least three untimed warm-up steps, whatever was asked for
    args.warmup = max(3, args.warmup)
    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(args.warmup):
        sweep_step()
    barrier()
    if sampler:
        sampler.mark_start()
    l0 = ctx.launch_count
    ev_all = []
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record(stream)
    for _ in range(args.steps):
        ev = []
        sweep_step(ev)
        ev_all.append(ev)
    t_end.record(stream)
    barrier()
    if sampler:
        sampler.mark_end()
    launches = ctx.launch_count - l0
    ms_total = max_over_ranks(t_start.elapsed_time(t_end))
    clocks = sampler.stop() if sampler else None
    ms_step = ms_total / args.steps
    value = world * nq * len(classes) / (ms_step * 1e-3)

    # per-class table and the roofline of the dominant kernel
    per = {}
    for ci, w in enumerate(work):
        ms = float(np.mean([ev[ci][0].elapsed_time(ev[ci][1]) for ev in ev_all]))
        qps = nq / (ms * 1e-3)
        per["".join(map(str, w["cl"]))] = {
            "ms": ms, "quartets_per_s": qps, "tflops": qps * w["flops"] / 1e12,
            "fp64_frac": qps * w["flops"] / 1e12 / fp64_peak,
            "hbm_gbs": qps * (8 * w["blk"] + 8) / 1e9, "hbm_frac": qps * (8 * w["blk"] + 8) / 1e9 / hbm_peak,
            "flops_per_quartet": w["flops"], "bytes_per_quartet": 8 * w["blk"] + 8}
    dom = max(per, key=lambda k: per[k]["ms"])
    d = per[dom]
    # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture
    # (profiles/traffic.json, written by scripts/ncu_traffic.py on the GPU box); null if the
    # capture was taken with another launch size
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        t = tj.get(dom)
        if t and int(t.get("launch_quartets", 0)) == chunk:
            traffic = float(t["dram_bytes"])
    except Exception:
        pass
    # Which roof binds the dominant kernel: its arithmetic intensity (model flops per algorithmic
    # byte; store mode writes every integral) against the machine balance FP64 peak / HBM
    # bandwidth.  (dd|dd): 28762 flops / 10376 B = 2.8 flop/B < 34 TF / 6.5 TB/s = 5.3 flop/B ->
    # the HBM roof is the lower one; the FP64-pipe fraction is reported beside it.
    bytes_q = 8 * work[list(per).index(dom)]["blk"] + 8
    intensity = d["flops_per_quartet"] / bytes_q
    balance = fp64_peak * 1e12 / (hbm_peak * 1e9)
    hbm_side = {"achieved": d["hbm_gbs"], "peak": hbm_peak, "unit": "GB/s", "frac": d["hbm_frac"],
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy bandwidth)" if peaks else
                               "fallback 6650 GB/s (MEASURED_PEAKS.json absent)"}
    fp64_side = {"achieved": d["tflops"], "peak": fp64_peak, "unit": "TFLOP/s", "frac": d["fp64_frac"],
                 "peak_source": "FP64 FMA probe (lb200_fp64_peak_probe) measured in this run; "
                                "MEASURED_PEAKS.json has no FP64 figure"}
    hbm_bound = intensity < balance
    roofline = {"bound": "hbm" if hbm_bound else "fp64",
                "kernel": "eri_rowreg_prim_kernel<%s> (%s|%s), store mode" % (dom, dom[:2], dom[2:])}
    roofline.update(hbm_side if hbm_bound else fp64_side)
    roofline.update({"traffic": traffic, "algorithmic_bytes": float(bytes_q * chunk),
                     "flops_per_byte": intensity, "machine_balance_flops_per_byte": balance,
                     "launch_ms": d["ms"] * chunk / nq,   # CUDA-event time of the class scaled to one full-size launch
                     "share_of_step": d["ms"] / ms_step,
                     "fp64": fp64_side, "hbm": hbm_side})

    # ---- e2e: host buffers through the C ABI (pinned tasks in, integrals out) ----------
    ne = min(args.e2e_quartets, nq)
    host_tasks = [w["tasks"][:ne].cpu().pin_memory() for w in work]
    # one pinned result buffer of bounded size (1 GiB per rank: eight ranks share the host's RAM),
    # refilled call after call; a class whose results exceed it takes several calls
    host_cap = min(ne * max_blk, 1 << 27)
    host_out = torch.empty(host_cap, dtype=torch.float64).pin_memory()
    h2d = sum(8 * ne for _ in work)
    d2h = sum(8 * w["blk"] * ne for w in work)

    def e2e_step():
        for w, ht in zip(work, host_tasks):
            per = max(1, host_cap // w["blk"])
            tasks_np, out_np = ht.numpy(), host_out.numpy()
            for t0 in range(0, ne, per):
                m = min(per, ne - t0)
                capi.eri_batch(ctx, w["bra"], w["ket"], tasks_np[t0:t0 + m],
                               out=out_np[:m * w["blk"]].reshape(m, w["blk"]))

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    nrep = max(1, min(args.steps, 2))
    for _ in range(nrep):
        e2e_step()
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / nrep)
    e2e = {"value": world * ne * len(classes) / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "quartets_per_class": ne,
           "note": "host-buffer lb200_eri_batch (pinned tasks in, integrals out into a 1 GiB pinned buffer, "
                   "several calls for the large classes); PCIe-bound by the materialised integrals"}

    # ---- the consumer: direct Fock build (configs[2]) -----------------------------------
    fock = None
    if not args.no_fock:
        try:
            fock = run_fock(args, ctx, dev, stream, rank, world, barrier, max_over_ranks, allreduce_sum_, fp64_peak)
        except capi.Lb200Error as e:
            fock = {"error": str(e)}

    # ---- configs[4]: (H2O)_256 / cc-pVTZ direct J/K build sharded over the ranks + Fock all-reduce.  On by
    # default for N >= 2 (one rank needs ~100 s per build; --fock256 forces it at N = 1).
    fock256 = None
    if args.fock256 == "on" or (args.fock256 == "auto" and world >= 2 and not args.no_fock):
        try:
            fock256 = run_fock(args, ctx, dev, stream, rank, world, barrier, max_over_ranks, allreduce_sum_, fp64_peak,
                               waters="8,8,4", basis="cc-pvtz", warm=False, e2e=False, profile=world >= 4,
                               cpu_leg=False)
        except capi.Lb200Error as e:
            fock256 = {"error": str(e)}

    # ---- 3-centre (P|mu nu) class sweep (configs[3]): C40H82, def2-TZVP / def2-universal-JKFIT -
    df3c = None
    if not args.no_df3c:
        try:
            df3c = run_df3c(args, ctx, dev, stream, out, barrier, max_over_ranks, world, rank, hbm_peak)
        except capi.Lb200Error as e:
            df3c = {"error": str(e)}

    # ---- two-body forces (first derivatives) of a water cluster -------------------------
    grad = None
    if not args.no_grad:
        try:
            grad = run_grad(args, ctx, dev, stream, rank, world, barrier, max_over_ranks)
        except Exception as e:   # a "next" row must never take the headline metric down with it
            grad = {"error": "%s: %s" % (type(e).__name__, e)}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "eri_class_sweep", "classes": len(classes), "lmax": 2,
                       "quartets_per_class": nq, "contraction": 1, "pairs_per_side": args.npairs,
                       "launch_quartets": chunk,
                       "l2_policy": "outputs (%.1f GB per class) exceed L2; pair tables are L2-resident by design"
                                    % (8 * max_blk * chunk / 1e9)},
            "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "per_class": per, "fock": fock, "fock256": fock256, "df3c": df3c, "grad": grad,
            # the strong-scaled half of the metric as top-level numbers: seconds per Fock build at this N
            "fock_build_seconds": fock.get("seconds") if isinstance(fock, dict) else None,
            "fock256_build_seconds": fock256.get("seconds") if isinstance(fock256, dict) else None}

    if rank == 0 and not args.no_cpu_baseline:
        line["parity"] = sweep_parity(ctx, work, args.npairs, args.parity_quartets)
        ncores = os.cpu_count() or 1
        mix, cper, sample = cpu_sweep(classes, args.npairs, args.cpu_seconds, ncores)
        # round 1's softer baseline once, so the effect of the two changes is visible
        mix_r1, _, sample_r1 = cpu_sweep(classes, args.npairs, min(5.0, args.cpu_seconds), ncores, fast=False,
                                         use_pairs=False)
        line["cpu_baseline"] = {"value": mix, "unit": UNIT, "cores": ncores, "kind": "port",
                                "sample": sample, "per_class_quartets_per_s": cper,
                                "round1_method": {"value": mix_r1, "sample": sample_r1},
                                "note": "reference libint2::Engine (its own headers) on run-time-loop restated "
                                        "kernels: stock generated libint (unrolled, CSE'd) is faster per quartet; "
                                        "the generator cannot be built in this image (DESIGN.md section 2)"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()
    return 0


def sweep_parity(ctx, work, npairs, per_class=20000):
    """BASELINE configs[1]'s tolerance check on the sweep's own inputs, with an arbiter: per class,
    `per_class` quartets of the bench geometry through lb200_eri_batch (GPU), the reference Engine
    (oracle, CPU) and the extended-precision truth of oracle/truth.cc (long double, all host threads).
    Reported per class and in total: elements outside the literal 1e-12 rel / 1e-14 abs tolerance
    AGAINST THE TRUTH for the GPU and for the reference, their max / rms errors, and the shell sets on
    which the GPU is further from the truth than both the tolerance and the reference
    (oracle.pyoracle.parity_stats).  The reference itself misses the literal tolerance where the HRR
    cancels (tests/eri/test.cc:77-83), so the criterion is gpu_vs_truth <= oracle_vs_truth."""
    from libint_b200 import capi
    from oracle import pyoracle as po
    nthr = os.cpu_count() or 1
    per = {}
    tot = {"integrals": 0, "gpu_outside": 0, "oracle_outside": 0, "worse_sets": 0, "shell_sets": 0}
    gmax = omax = 0.0
    t0 = time.perf_counter()
    for w in work:
        t = w["tasks"][:per_class].cpu().numpy()
        got = capi.eri_batch(ctx, w["bra"], w["ket"], t)
        sh = po.Shells(*w["tab"], raw=False)
        q4 = np.stack([t[:, 0], npairs + t[:, 0], 2 * npairs + t[:, 1], 3 * npairs + t[:, 1]], axis=1).astype(np.int32)
        orc = po.compute_batch(sh, q4, nthreads=nthr)
        hi, lo = po.truth_batch(sh, q4, nthreads=nthr)
        st = po.parity_stats(got, orc, hi, lo)
        per["".join(map(str, w["cl"]))] = {k: st[k] for k in (
            "gpu_outside", "oracle_outside", "gpu_max_abs", "oracle_max_abs", "gpu_rms", "oracle_rms",
            "gpu_max_scaled", "oracle_max_scaled", "worse_sets", "max_ratio_nonliteral")}
        for k in tot:
            tot[k] += st[k]
        gmax, omax = max(gmax, st["gpu_max_scaled"]), max(omax, st["oracle_max_scaled"])
    return {"against": "extended-precision truth (oracle/truth.cc, long double); reference = libint2::Engine (oracle)",
            "tolerance": "1e-12 rel + 1e-14 abs, literal", "quartets_per_class": per_class,
            "integrals_checked": tot["integrals"], "gpu_vs_truth_outside": tot["gpu_outside"],
            "oracle_vs_truth_outside": tot["oracle_outside"],
            "gpu_vs_truth_max_scaled": gmax, "oracle_vs_truth_max_scaled": omax,
            "sets_gpu_worse_than_tolerance_and_reference": tot["worse_sets"], "shell_sets": tot["shell_sets"],
            "classes_gpu_rms_le_oracle_rms": int(sum(1 for v in per.values() if v["gpu_rms"] <= v["oracle_rms"])),
            "classes": len(per), "seconds": time.perf_counter() - t0, "per_class": per}


def run_df3c(args, ctx, dev, stream, out, barrier, max_over_ranks, world, rank=0, hbm_peak=6469.9):
    """configs[3]: every (P|mu nu) shell triplet of an all-trans alkane once, as implicit (DF shell) x (orbital
    pair) products per class (lb200_eri_product, Cartesian, materialised in HBM).  Sharded by DF shell over the
    ranks with no collective: each rank sweeps its contiguous share of every bra block.  Beside
    it: the reference Engine's xs_xx loop on the host cores (bounded sample), the HBM roofline of the write,
    and the density-fitted Fock build that consumes the integrals (libint_b200.dfjk)."""
    import torch
    from libint_b200.basis import BasisSet, alkane
    from libint_b200.df3c import ThreeCenter
    atoms = alkane(args.df3c_carbons)
    obs, dfbs = BasisSet("def2-tzvp", atoms), BasisSet("def2-tzvp-jk", atoms)
    t0 = time.perf_counter()
    tc = ThreeCenter(ctx, obs, dfbs)
    setup_s = time.perf_counter() - t0
    with torch.cuda.stream(stream):
        tc.sweep(out, rank=rank, nranks=world)
        barrier()
        ev = []
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        n_mine = tc.sweep(out, events=ev, rank=rank, nranks=world)
        e1.record(stream)
        barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    n = tc.ntriplets()
    nints_mine = sum(nn * blk for _, nn, blk, _, _ in ev)
    agg = {}
    for c, nn, blk, a, b in ev:   # merge the contraction buckets of a class
        e = agg.setdefault(c, [0, 0.0, blk])
        e[0] += nn
        e[1] += a.elapsed_time(b)
    top = sorted(agg.items(), key=lambda kv: -kv[1][1])[:6]
    dom_c, (dom_n, dom_ms, dom_blk) = top[0]
    res = {"workload": "C%dH%d (P|mu nu), obs def2-tzvp (%d shells, %d bf), dfbs def2-tzvp-jk (%d shells, %d bf, "
                       "max l %d), %d significant orbital pairs" % (args.df3c_carbons, 2 * args.df3c_carbons + 2,
                                                                     len(obs), obs.nbf, len(dfbs), dfbs.nbf,
                                                                     dfbs.max_l, tc.npairs),
           "shell_triplets": n, "classes": len(tc.classes()), "launch_groups": len(tc.blocks()), "seconds": ms * 1e-3,
           "triplets_per_s": n / (ms * 1e-3), "n_gpus": world, "scaling": "strong",
           "sharding": "by DF shell (bra rows of every class block), no collective",
           "cartesian_integrals_this_rank": nints_mine,
           "hbm_write_gbs": nints_mine * 8 / (ms * 1e-3) / 1e9, "setup_seconds": setup_s,
           "roofline": {"bound": "hbm", "kernel": "store-mode class kernel (%d s|%d %d)" % dom_c,
                        "achieved": dom_n * dom_blk * 8 / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0,
                        "peak": hbm_peak, "unit": "GB/s",
                        "frac": dom_n * dom_blk * 8 / (dom_ms * 1e-3) / 1e9 / hbm_peak if dom_ms > 0 else 0.0,
                        "whole_sweep_frac": nints_mine * 8 / (ms * 1e-3) / 1e9 / hbm_peak,
                        "note": "algorithmic bytes = 8 x Cartesian integrals written; the sweep is 100+ small "
                                "launches, the dominant class is timed with CUDA events around its launches"},
           "slowest_classes": [{"class": "(%d s|%d %d)" % c, "triplets": nn,
                                "ns_per_triplet": 1e6 * t / nn} for c, (nn, t, _) in top]}
    if rank == 0 and not args.no_cpu_baseline:
        # the reference's DF set-up loop on the host cores: a bounded random sample of the same triplets
        from oracle import pyoracle as po
        ncores = os.cpu_count() or 1
        rng = np.random.default_rng(11)
        ns = 4000000
        k = rng.integers(0, tc.npairs, ns)
        t3 = np.stack([rng.integers(0, len(dfbs), ns), tc.pair_a[k], tc.pair_b[k]], axis=1).astype(np.int32)
        sec, _ = po.time_triplets(po.Shells(*dfbs.flat(), raw=False), po.Shells(*obs.flat(), raw=False), t3, ncores)
        res["cpu_baseline"] = {"value": ns / sec, "unit": "shell triplets/s", "cores": ncores, "kind": "port",
                               "sample": "%d random (DF shell, significant orbital pair) triplets of the same workload, "
                                         "reference Engine xs_xx, %.1f s" % (ns, sec)}
    # the consumer: density-fitted G = 2J - K with streamed slabs (one rank; needs the memory of W)
    if rank == 0 and args.df3c_carbons >= 8:
        try:
            from libint_b200.dfjk import DFFockBuilder
            with torch.cuda.stream(stream):
                t0 = time.perf_counter()
                fb = DFFockBuilder(obs, dfbs, ctx=ctx, slab_bytes=8 << 30, threshold=0.0)
                torch.cuda.synchronize(dev)
                dsetup = time.perf_counter() - t0
                nocc = sum(a.atomic_number for a in atoms) // 2
                C = torch.linalg.qr(torch.randn((obs.nbf, nocc), dtype=torch.float64, device=dev,
                                                generator=torch.Generator(device=dev).manual_seed(3)))[0]
                fb(C)
                torch.cuda.synchronize(dev)
                d0 = torch.cuda.Event(enable_timing=True)
                d1 = torch.cuda.Event(enable_timing=True)
                d0.record(stream)
                G = fb(C)
                d1.record(stream)
                torch.cuda.synchronize(dev)
            res["dfjk"] = {"workload": "density-fitted G = 2J - K, nocc %d, ndf %d, %d slabs x 2 sweeps" % (nocc, dfbs.nbf, len(fb.slabs)),
                           "seconds": d0.elapsed_time(d1) * 1e-3, "setup_seconds_metric_cholesky": dsetup,
                           "checksum": float(G.abs().sum().item()), "symmetric_err": float((G - G.T).abs().max().item())}
            del fb, G, C
            torch.cuda.empty_cache()
        except Exception as e:   # noqa: BLE001 -- report, do not lose the bench line
            res["dfjk"] = {"error": repr(e)[:300]}
    return res


def fock_roofline(f, fp64_peak, build_seconds, pure_basis=True):
    """Per-class FP64 roofline of the last profiled Fock build (lb200_fock_get_profile): model flops
    (libint_b200/flops.py, SURVEY 8d: K_eff * F_prim + F_hrr with K_eff = surviving primitive quartets
    counted in the kernel) over the CUDA-event time of that class's launches, plus the digestion's
    12 flops per integral (hartree-fock++.cc:1721-1743) reported beside it."""
    from libint_b200.flops import hrr_flops, prim_flops, canonical
    rows = f.profile()
    agg = {}
    for la, lb, lc, ld, bb, bk, ms, nq, npq in rows:
        cl = canonical(int(la), int(lb), int(lc), int(ld))
        e = agg.setdefault(cl, {"ms": 0.0, "quartets": 0.0, "prim_quartets": 0.0, "launch_groups": 0})
        e["ms"] += ms
        e["quartets"] += nq
        e["prim_quartets"] += npq
        e["launch_groups"] += 1
    out, tot_flops, tot_dig, tot_ms = [], 0.0, 0.0, 0.0
    for cl, e in agg.items():
        fl = e["prim_quartets"] * prim_flops(*cl) + e["quartets"] * hrr_flops(*cl)
        nfun = 1
        for l in cl:
            nfun *= (2 * l + 1) if (pure_basis and l >= 2) else (l + 1) * (l + 2) // 2
        dig = 12.0 * nfun * e["quartets"]
        tot_flops += fl
        tot_dig += dig
        tot_ms += e["ms"]
        out.append({"class": "(%d%d|%d%d)" % cl, "quartets": e["quartets"],
                    "k_eff": e["prim_quartets"] / max(1.0, e["quartets"]), "ms": e["ms"],
                    "model_gflop": fl / 1e9, "digest_gflop": dig / 1e9,
                    "tflops": fl / (e["ms"] * 1e-3) / 1e12 if e["ms"] > 0 else 0.0,
                    "fp64_frac": fl / (e["ms"] * 1e-3) / 1e12 / fp64_peak if e["ms"] > 0 else 0.0})
    out.sort(key=lambda r: -r["ms"])
    return {"bound": "fp64", "peak": fp64_peak, "unit": "TFLOP/s",
            "peak_source": "FP64 FMA probe of this run (148 SM x 64 DFMA/clk x 2 x 1.965 GHz = 37.2 TF nominal)",
            "model_tflop_total": tot_flops / 1e12, "digest_tflop_total": tot_dig / 1e12,
            "class_kernel_ms_profiled": tot_ms,
            "achieved_in_class_kernels": tot_flops / (tot_ms * 1e-3) / 1e12 if tot_ms > 0 else 0.0,
            "frac_in_class_kernels": tot_flops / (tot_ms * 1e-3) / 1e12 / fp64_peak if tot_ms > 0 else 0.0,
            "achieved": tot_flops / build_seconds / 1e12, "frac": tot_flops / build_seconds / 1e12 / fp64_peak,
            "frac_with_digestion": (tot_flops + tot_dig) / build_seconds / 1e12 / fp64_peak,
            "note": "achieved/frac: this rank's model flops over the timed (unprofiled) build; per-class rows from "
                    "a separate profiled build (one stream sync per launch)",
            "classes": len(out), "top_classes": out[:12], "all_classes": out}


def run_grad(args, ctx, dev, stream, rank, world, barrier, max_over_ranks):
    """Two-body forces F2 (compute_2body_fock_deriv<1> traced with D, hartree-fock++.cc:642-656,1775-2055) of a
    water cluster, quartets sharded over the ranks like the Fock build, 3 * natoms partial sums all-reduced."""
    import torch
    from libint_b200.basis import BasisSet, water_cluster
    from libint_b200.fock import FockBuilder
    nx, ny, nz = [int(x) for x in args.grad_waters.split(",")]
    atoms = water_cluster(nx, ny, nz)
    obs = BasisSet(args.grad_basis, atoms)
    fb = FockBuilder(obs, ctx=ctx, rank=rank, nranks=world)
    n = obs.nbf
    rng = np.random.default_rng(7)
    C = rng.standard_normal((n, max(1, n // 8))) / np.sqrt(n)
    Dd = torch.from_numpy(C @ C.T).to(dev)
    with torch.cuda.stream(stream):
        fb.forces_2body(Dd, precision=args.fock_precision)   # warm-up (builds the shifted pair-block twins)
    barrier()
    t0 = time.perf_counter()
    with torch.cuda.stream(stream):
        g, st = fb.forces_2body(Dd, precision=args.fock_precision, stats=True)
    barrier()
    sec = max_over_ranks(time.perf_counter() - t0)
    nq = st["nquartets"]
    if world > 1:
        t = torch.tensor([nq], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t)
        nq = float(t.item())
    # the other GPU half of the force block: one-body + Pulay sums (lb200_onebody_forces, hartree-fock++.cc:601-627);
    # set-up sized, every rank evaluates all of it; its own guard so that it can never cost the two-body numbers
    try:
        Dn = C @ C.T
        W = (C * rng.uniform(-1.0, -0.1, C.shape[1])) @ C.T   # like C_occ eps_occ C_occ^T
        fb.forces_1body(Dn, W, atoms)
        t1 = time.perf_counter()
        F1, FP = fb.forces_1body(Dn, W, atoms)
        onebody = {"seconds": time.perf_counter() - t1,   # this rank's own clock: no collective inside the guard
                   "max_net_force": float(max(np.abs(F1.sum(axis=0)).max(), np.abs(FP.sum(axis=0)).max())),
                   "checksum": float(np.abs(F1).sum() + np.abs(FP).sum()),
                   "note": "host D, W in, 2 x 3 natoms sums out (uploads and Cartesian-isation included)"}
    except Exception as e:
        onebody = {"error": "%s: %s" % (type(e).__name__, e)}
    return {"workload": "(H2O)_%d / %s two-body forces, Schwarz x density screened at %g"
                        % (nx * ny * nz, args.grad_basis, args.fock_precision),
            "onebody_forces": onebody,
            "nbf": n, "natoms": len(atoms), "shell_quartets": nq, "seconds": sec,
            "derivative_shell_sets_per_s": 12 * nq / sec, "device_ms_this_rank": st["ms"],
            "launches_this_rank": st["launches"], "n_gpus": world, "scaling": "strong",
            "max_net_force": float(np.abs(g.sum(axis=0)).max()), "checksum": float(np.abs(g).sum()),
            "cpu_baseline": None,
            "note": "six shifted-class store launches + one contraction per (bra class, ket class) chunk; no CPU "
                    "baseline: the reference's eri1 kernels are generated code that cannot be built here, and the "
                    "closed-form derivative oracle is a checker, not an implementation to time"}


def run_fock(args, ctx, dev, stream, rank, world, barrier, max_over_ranks, allreduce_sum_, fp64_peak=None,
             waters=None, basis=None, warm=True, e2e=True, profile=True, cpu_leg=True):
    """One strong-scaled direct Fock build (quartets sharded over the ranks by bra row, partial G's summed
    by lb200_fock_allreduce inside the timed region): configs[2] by default, configs[4] with
    waters="8,8,4", basis="cc-pvtz"."""
    import torch
    from libint_b200 import capi
    from libint_b200.basis import BasisSet, water_cluster
    basis = basis or args.fock_basis
    nx, ny, nz = [int(x) for x in (waters or args.fock_waters).split(",")]
    atoms = water_cluster(nx, ny, nz)
    obs = BasisSet(basis, atoms)
    t0 = time.perf_counter()
    B = capi.Basis(ctx, *obs.flat())
    f = capi.Fock(ctx, B)
    setup_s = time.perf_counter() - t0
    n = obs.nbf
    rng = np.random.default_rng(7)
    C = rng.standard_normal((n, max(1, n // 8))) / np.sqrt(n)
    D = C @ C.T  # symmetric PSD, like C_occ C_occ^T
    Dh = torch.from_numpy(D).pin_memory()
    Dd = Dh.to(dev)
    G = torch.empty((n, n), dtype=torch.float64, device=dev)
    Gh = torch.empty((n, n), dtype=torch.float64).pin_memory()

    from libint_b200.fock import make_comm
    comm = make_comm(ctx)   # lb200_comm_create: NCCL behind the C ABI (None on one rank)

    def reduce_(G):
        if comm is not None:
            comm.allreduce_(G)   # lb200_fock_allreduce on the context's stream (= `stream`)
        else:
            allreduce_sum_(G)

    def build():
        with torch.cuda.stream(stream):
            f.build(Dd, args.fock_precision, rank=rank, nranks=world, out=G)
            reduce_(G)

    if warm:
        build()  # warm-up
    barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    build()
    e1.record(stream)
    barrier()
    sec = max_over_ranks(e0.elapsed_time(e1) * 1e-3)
    # end to end from host D to host G
    e2e_sec = None
    if e2e:
        barrier()
        t0 = time.perf_counter()
        with torch.cuda.stream(stream):
            Dd.copy_(Dh, non_blocking=True)
            f.build(Dd, args.fock_precision, rank=rank, nranks=world, out=G)
            reduce_(G)
            Gh.copy_(G, non_blocking=True)
        barrier()
        e2e_sec = max_over_ranks(time.perf_counter() - t0)
    else:
        with torch.cuda.stream(stream):
            Gh.copy_(G, non_blocking=True)
        barrier()
    roof = None
    if profile:
        f.set_profile(True)
        _, st = f.build(Dd, args.fock_precision, rank=rank, nranks=world, out=G, stats=True)
        f.set_profile(False)
        roof = fock_roofline(f, fp64_peak, sec) if fp64_peak else None
        nquart = st["nquartets"]
    else:
        nquart = float("nan")
    if world > 1 and profile:
        t = torch.tensor([nquart], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t)
        nquart = float(t.item())
    # CPU side of the same metric: the reference's multithreaded compute_2body_fock pattern
    # (oracle: reference Engine per thread on the restated kernels) on the box's host cores.  Its
    # task enumeration alone is O(npair^2) per thread, so it is timed on a bounded sample of the
    # workload -- an (H2O)_8 sub-cluster of the same lattice and basis -- and the GPU is timed on
    # that sub-cluster as well.
    cpu = None
    if rank == 0 and not args.no_cpu_baseline and cpu_leg:
        from oracle import pyoracle as po
        obs8 = BasisSet(basis, water_cluster(2, 2, 2))
        B8 = capi.Basis(ctx, *obs8.flat())
        f8 = capi.Fock(ctx, B8)
        C8 = rng.standard_normal((obs8.nbf, max(1, obs8.nbf // 8))) / np.sqrt(obs8.nbf)
        D8 = C8 @ C8.T
        f8.build(D8, args.fock_precision)
        t0 = time.perf_counter()
        G8, st8 = f8.build(D8, args.fock_precision, stats=True)
        gpu8_s = st8["ms"] * 1e-3
        ncores = os.cpu_count() or 1
        # timing build of the oracle (-O3, AVX2+FMA) with the reference's precomputed SchwarzInf ShellPairs
        of = po.Fock(po.Shells(*obs8.flat(), raw=False), f8.pair_s1, f8.pair_s2, nthreads=ncores, fast=True)
        Gc, stc = of.build(D8, args.fock_precision)
        err = float(np.max(np.abs(G8 - Gc) - 1e-12 * np.abs(Gc)))
        cpu = {"value": stc["nquartets"] / stc["seconds"], "unit": "shell quartets/s", "cores": ncores,
               "kind": "port", "seconds": stc["seconds"], "build": "-O3 -march=x86-64-v3" if po.fast_available() else "-O2 -march=x86-64-v2",
               "sample": "(H2O)_8 / %s sub-cluster, full build, %d shell quartets" % (basis, int(stc["nquartets"])),
               "gpu_seconds_same_sample": gpu8_s, "gpu_quartets_per_s_same_sample": st8["nquartets"] / gpu8_s,
               "max_abs_err_beyond_1e-12_rel": err}
    return {"cpu_baseline": cpu,
            "workload": "(H2O)_%d / %s direct Fock (J - K/2), Schwarz x density screened at %g"
                        % (nx * ny * nz, basis, args.fock_precision),
            "nshell": len(obs), "nbf": n, "significant_pairs": int(len(f.pair_s1)),
            "shell_quartets": nquart, "seconds": sec, "e2e_seconds": e2e_sec,
            "quartets_per_s": nquart / sec, "setup_seconds": setup_s, "n_gpus": world, "scaling": "strong",
            "timed_builds": 1, "warmup_builds": 1 if warm else 0,
            "allreduce": "lb200_fock_allreduce: ncclAllReduce(sum, f64, nbf^2) on the build stream" if world > 1 else None, "roofline": roof,
            "checksum": float(Gh.abs().sum().item())}


if __name__ == "__main__":
    sys.exit(main())
