"""Per class-pair time of one Fock build: LB200_FOCK_PROFILE=1 python scripts/fock_profile.py [basis] [nx,ny,nz]"""
import os, sys, time
import numpy as np
sys.path.insert(0, ".")
os.environ["LB200_FOCK_PROFILE"] = "1"
from libint_b200 import capi
from libint_b200.basis import BasisSet, water_cluster
basis = sys.argv[1] if len(sys.argv) > 1 else "def2-tzvp"
nx, ny, nz = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "4,4,4").split(",")]
ctx = capi.Context(0)
obs = BasisSet(basis, water_cluster(nx, ny, nz))
B = capi.Basis(ctx, *obs.flat())
t0 = time.time(); f = capi.Fock(ctx, B); print("setup %.2f s, pairs %d" % (time.time() - t0, len(f.pair_s1)))
n = obs.nbf
rng = np.random.default_rng(7)
C = rng.standard_normal((n, max(1, n // 8))) / np.sqrt(n)
D = C @ C.T
G = f.build(D, 1e-10)
t0 = time.time(); G, st = f.build(D, 1e-10, stats=True); print("build %.3f s" % (time.time() - t0), st)
if len(sys.argv) > 3:   # one rank's share of an N-rank build (strong-scaling estimate on one GPU)
    os.environ.pop("LB200_FOCK_PROFILE", None)
    for N in [int(x) for x in sys.argv[3].split(",")]:
        ts = []
        for r in range(min(N, 2)):
            f.build(D, 1e-10, rank=r, nranks=N)
            t0 = time.time(); f.build(D, 1e-10, rank=r, nranks=N); ts.append(time.time() - t0)
        print("nranks %d: rank shares %s s" % (N, ", ".join("%.3f" % t for t in ts)))
