"""GPU vs reference Engine (oracle) vs extended-precision truth on BASELINE configs[1]'s own inputs.

  python scripts/parity_truth.py [--quartets 100000] [--lmax 2] [--out profiles/r02_parity_truth.json]
                                 [--lib-suffix _nofma]

For every canonical class of the sweep: N random quartets of the bench geometry through
lb200_eri_batch (GPU), libint2::Engine::compute2 (oracle, CPU) and oracle/truth.cc (long double;
a 256-quartet subsample is re-checked in __float128), then oracle.pyoracle.parity_stats."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quartets", type=int, default=100000)
    ap.add_argument("--npairs", type=int, default=4096)
    ap.add_argument("--lmax", type=int, default=2)
    ap.add_argument("--out", default="")
    ap.add_argument("--lib-suffix", default=None)
    args = ap.parse_args()
    if args.lib_suffix is not None:
        os.environ["LB200_LIB_SUFFIX"] = args.lib_suffix
    from bench import class_table, sweep_classes
    from libint_b200 import capi
    from oracle import pyoracle as po
    ctx = capi.Context(0)
    nthr = os.cpu_count() or 1
    res = {"quartets_per_class": args.quartets, "threads": nthr, "lib": capi.LIB_PATH, "classes": {}}
    tot = {"gpu_outside": 0, "oracle_outside": 0, "integrals": 0, "worse_sets": 0}
    for ci, cl in enumerate(sweep_classes(args.lmax)):
        tab = class_table(cl, args.npairs, ci)
        n = args.npairs
        bs = capi.Basis(ctx, *tab)
        i = np.arange(n, dtype=np.int32)
        bra = capi.Pairs(ctx, bs, bs, i, n + i)
        ket = capi.Pairs(ctx, bs, bs, 2 * n + i, 3 * n + i)
        rng = np.random.default_rng(777 + ci)
        t = rng.integers(0, n, (args.quartets, 2)).astype(np.int32)
        q4 = np.stack([t[:, 0], n + t[:, 0], 2 * n + t[:, 1], 3 * n + t[:, 1]], axis=1).astype(np.int32)
        t0 = time.time()
        got = capi.eri_batch(ctx, bra, ket, t)
        sh = po.Shells(*tab, raw=False)
        t1 = time.time()
        orc = po.compute_batch(sh, q4, nthreads=nthr)
        t2 = time.time()
        hi, lo = po.truth_batch(sh, q4, nthreads=nthr)
        t3 = time.time()
        st = po.parity_stats(got, orc, hi, lo)
        nq = min(256, args.quartets)
        hq, lq = po.truth_batch(sh, q4[:nq], nthreads=nthr, quad=True)
        st["long_double_vs_quad_max_abs"] = float(np.abs((hi[:nq] - hq) + (lo[:nq] - lq)).max())
        st["seconds"] = {"gpu": t1 - t0, "oracle": t2 - t1, "truth": t3 - t2}
        name = "".join(map(str, cl))
        res["classes"][name] = st
        for k in tot:
            tot[k] += st[k]
        print("(%s|%s) gpu: out %6d max %.2e rms %.2e | oracle: out %6d max %.2e rms %.2e | gpu-orc %.2e worse_sets %d closer g/o %d/%d"
              % (name[:2], name[2:], st["gpu_outside"], st["gpu_max_abs"], st["gpu_rms"], st["oracle_outside"],
                 st["oracle_max_abs"], st["oracle_rms"], st["gpu_vs_oracle_max_abs"], st["worse_sets"],
                 st["sets_gpu_closer"], st["sets_oracle_closer"]), flush=True)
    res["total"] = tot
    print(json.dumps(tot))
    if args.out:
        json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
