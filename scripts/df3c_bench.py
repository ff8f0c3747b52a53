"""3-centre (P|mu nu) sweep for an alkane (BASELINE configs[3]): python scripts/df3c_bench.py [ncarbon] [check]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
from libint_b200 import capi
from libint_b200.basis import BasisSet, alkane
from libint_b200.df3c import ThreeCenter

nC = int(sys.argv[1]) if len(sys.argv) > 1 else 40
dev = torch.device("cuda", 0)
ctx = capi.Context(0)
stream = torch.cuda.Stream(dev)
ctx.set_stream(stream.cuda_stream)
atoms = alkane(nC)
obs, dfbs = BasisSet("def2-tzvp", atoms), BasisSet("def2-tzvp-jk", atoms)
t0 = time.time()
tc = ThreeCenter(ctx, obs, dfbs)
print("C%dH%d: obs %d shells / %d bf, dfbs %d shells / %d bf (max l %d), %d significant pairs, %d classes, "
      "%.3e shell triplets; setup %.2f s" % (nC, 2 * nC + 2, len(obs), obs.nbf, len(dfbs), dfbs.nbf, dfbs.max_l,
                                              tc.npairs, len(tc.classes()), tc.ntriplets(), time.time() - t0))
out = torch.empty(1 << 27, dtype=torch.float64, device=dev)  # 1 GiB chunk buffer
with torch.cuda.stream(stream):
    tc.sweep(out)  # warm-up
    torch.cuda.synchronize()
    ev = []
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    n = tc.sweep(out, events=ev)
    e1.record(stream)
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
nints = sum(nn * blk for _, nn, blk, _, _ in ev)
print("sweep: %d triplets, %.3e Cartesian integrals (%.1f GB) in %.1f ms -> %.3e triplets/s, %.1f GB/s written"
      % (n, nints, nints * 8 / 1e9, ms, n / ms * 1e3, nints * 8 / ms / 1e6))
for cls, nn, blk, a, b in sorted(ev, key=lambda x: -x[3].elapsed_time(x[4]))[:12]:
    t = a.elapsed_time(b)
    print("  (%d s|%d %d) %10d triplets %8.2f ms %8.2f ns/triplet" % (cls + (nn, t, 1e6 * t / nn)))
if len(sys.argv) > 2:
    # spot check against the Engine mirror (single-triplet path) on a few random triplets
    from libint_b200.engine import Engine, Operator, BraKet
    rng = np.random.default_rng(0)
    eng = Engine(Operator.coulomb, max(obs.max_nprim, dfbs.max_nprim), max(obs.max_l, dfbs.max_l), ctx=ctx)
    eng.set(BraKet.xs_xx)
    print("engine mirror ok")
