import sys, numpy as np
sys.path.insert(0, ".")
from libint_b200 import capi
from oracle import pyoracle as po
la, lb, lc, ld = [int(x) for x in sys.argv[1:5]]
K = int(sys.argv[5]) if len(sys.argv) > 5 else 1
rng = np.random.default_rng(7)
ctx = capi.Context(0)
n = 4
O = rng.uniform(-1.0, 1.0, (n, 3)); al = rng.uniform(0.2, 3.0, n * K); co = rng.uniform(0.2, 1.5, n * K)
l = [la, lb, lc, ld]
bs = capi.Basis(ctx, l, [0] * 4, [K] * 4, O, al, co)
bra = capi.Pairs(ctx, bs, bs, [0], [1]); ket = capi.Pairs(ctx, bs, bs, [2], [3])
got = capi.eri_batch(ctx, bra, ket, np.array([[0, 0]], dtype=np.int32))[0]
ref = po.compute2(po.Shells(l, [0] * 4, [K] * 4, O, al, co, raw=False), precision=0.0).ravel()
bad = np.nonzero(np.abs(got - ref) > 1e-14 + 1e-12 * np.abs(ref))[0]
print("class", l, "nbad", len(bad), "of", len(ref), "first bad idx", bad[:20])
nc = lambda l: (l + 1) * (l + 2) // 2
if len(bad):
    idx = np.array(np.unravel_index(bad, (nc(la), nc(lb), nc(lc), nc(ld)))).T
    print("bad a:", sorted(set(idx[:, 0])), "\nbad b:", sorted(set(idx[:, 1])), "\nbad c:", sorted(set(idx[:, 2])), "\nbad d:", sorted(set(idx[:, 3])))
