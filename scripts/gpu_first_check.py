"""First GPU sanity run: every built class against the CPU oracle on random contracted shells."""
import itertools, sys, time
import numpy as np
sys.path.insert(0, ".")
from libint_b200 import capi
from oracle import pyoracle as po

rng = np.random.default_rng(7)
ctx = capi.Context(0)
unit = capi.Basis.unit(ctx)

def rand_basis(ls, K):
    n = len(ls)
    O = rng.uniform(-1.0, 1.0, (n, 3))
    al = rng.uniform(0.2, 3.0, n * K)
    co = rng.uniform(0.2, 1.5, n * K)
    return ls, [0] * n, [K] * n, O, al, co

worst = 0
pcs = [(a, b) for a in range(4) for b in range(a + 1)] + [(4, 0)]
t0 = time.time()
fails = []
for (la, lb), (lc, ld) in itertools.product(pcs, pcs):
    K = 2 if la + lb + lc + ld <= 8 else 1
    l, pure, nprim, O, al, co = rand_basis([la, lb, lc, ld], K)
    bs = capi.Basis(ctx, l, pure, nprim, O, al, co)
    bra = capi.Pairs(ctx, bs, bs, [0], [1])
    ket = capi.Pairs(ctx, bs, bs, [2], [3])
    got = capi.eri_batch(ctx, bra, ket, np.array([[0, 0]], dtype=np.int32))[0]
    ref = po.compute2(po.Shells(l, pure, nprim, O, al, co, raw=False), precision=0.0).ravel()
    err = np.abs(got - ref).max()
    rel = (np.abs(got - ref) / (np.abs(ref) + 1e-300)).max()
    tol_ok = np.all(np.abs(got - ref) <= 1e-14 + 1e-12 * np.abs(ref))
    worst = max(worst, err)
    if not tol_ok:
        fails.append((la, lb, lc, ld, err, rel))
    print("(%d%d|%d%d) maxabs %.2e maxrel %.2e max|ref| %.2e %s" % (la, lb, lc, ld, err, rel, np.abs(ref).max(), "ok" if tol_ok else "FAIL"), flush=True)
print("worst abs", worst, "fails", fails, "time", time.time() - t0)
