"""One direct Fock build of a water cluster, timed: python scripts/fock_once.py basis nx,ny,nz [precision]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from libint_b200 import capi
from libint_b200.basis import BasisSet, water_cluster
basis = sys.argv[1]
nx, ny, nz = [int(x) for x in sys.argv[2].split(",")]
prec = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-10
ctx = capi.Context(0)
t0 = time.time(); obs = BasisSet(basis, water_cluster(nx, ny, nz)); B = capi.Basis(ctx, *obs.flat())
print("(H2O)_%d / %s: %d shells, %d bf, max_nprim %d, max_l %d (%.1f s)" % (nx * ny * nz, basis, len(obs), obs.nbf, obs.max_nprim, obs.max_l, time.time() - t0), flush=True)
t0 = time.time(); f = capi.Fock(ctx, B); print("fock_create: %d significant pairs, %.1f s" % (len(f.pair_s1), time.time() - t0), flush=True)
n = obs.nbf
rng = np.random.default_rng(7)
C = rng.standard_normal((n, max(1, n // 8))) / np.sqrt(n)
D = C @ C.T
f.build(D, prec)   # warm-up: lazy module load, local-memory pool, allocations
t0 = time.time(); G, st = f.build(D, prec, stats=True)
print("build: wall %.2f s, device %.2f s, %.4e shell quartets (%.3e /s), %d launches, %.3e candidates; |G|_1 = %.10e; symmetric %s"
      % (time.time() - t0, st["ms"] * 1e-3, st["nquartets"], st["nquartets"] / (st["ms"] * 1e-3), st["launches"], st["candidates"], np.abs(G).sum(), np.array_equal(G, G.T)), flush=True)
