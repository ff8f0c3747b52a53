"""One-body + Pulay forces of a water cluster on the GPU, timed: python scripts/onebody_forces_once.py basis nx,ny,nz"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from libint_b200 import capi
from libint_b200.basis import BasisSet, water_cluster
basis = sys.argv[1]
nx, ny, nz = [int(x) for x in sys.argv[2].split(",")]
atoms = water_cluster(nx, ny, nz)
obs = BasisSet(basis, atoms)
ctx = capi.Context(0)
B = capi.Basis(ctx, *obs.flat())
n = obs.nbf
rng = np.random.default_rng(7)
C = rng.standard_normal((n, max(1, n // 8))) / np.sqrt(n)
D = C @ C.T
W = (C * rng.uniform(-1.0, -0.1, C.shape[1])) @ C.T
charges = [(float(a.atomic_number), a.xyz) for a in atoms]
for rep in range(2):
    t0 = time.time()
    F1, FP = capi.onebody_forces(ctx, B, charges, obs.shell2atom, D, W)
    dt = time.time() - t0
    print("(H2O)_%d / %s one-body + Pulay forces: %.3f s wall (%d shells, %d atoms); net |sum F1| %.2e, |sum FP| %.2e"
          % (nx * ny * nz, basis, dt, len(obs), len(atoms), np.abs(F1.sum(axis=0)).max(), np.abs(FP.sum(axis=0)).max()), flush=True)
