"""Two-body forces of a water cluster, timed: python scripts/forces_once.py basis nx,ny,nz [precision]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from libint_b200.basis import BasisSet, water_cluster
from libint_b200.fock import FockBuilder
basis = sys.argv[1]
nx, ny, nz = [int(x) for x in sys.argv[2].split(",")]
prec = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-10
obs = BasisSet(basis, water_cluster(nx, ny, nz))
fb = FockBuilder(obs, rank=0, nranks=1)
n = obs.nbf
rng = np.random.default_rng(7)
C = rng.standard_normal((n, max(1, n // 8))) / np.sqrt(n)
D = C @ C.T
fb.forces_2body(D, precision=prec)   # warm-up: builds the shifted twins of the pair blocks
t0 = time.time(); g, st = fb.forces_2body(D, precision=prec, stats=True)
print("(H2O)_%d / %s forces: wall %.3f s, device %.3f s, %.4e shell quartets (%.3e derivative shell sets/s), %d launches; "
      "|F2|_1 = %.10e, net force %.2e" % (nx * ny * nz, basis, time.time() - t0, st["ms"] * 1e-3, st["nquartets"],
                                         12 * st["nquartets"] / (st["ms"] * 1e-3), st["launches"], np.abs(g).sum(),
                                         np.abs(g.sum(axis=0)).max()), flush=True)
