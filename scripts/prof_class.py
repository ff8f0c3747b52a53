"""Run a few launches of one class kernel (for ncu): python scripts/prof_class.py la lb lc ld [nquartets] [reps]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import torch  # noqa: E402

from bench import class_table  # noqa: E402
from libint_b200 import capi  # noqa: E402

cl = tuple(int(x) for x in sys.argv[1:5])
nq = int(sys.argv[5]) if len(sys.argv) > 5 else 1 << 20
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 3
npairs = 4096
ctx = capi.Context(0)
tab = class_table(cl, npairs, 0)
bs = capi.Basis(ctx, *tab)
i = np.arange(npairs, dtype=np.int32)
bra = capi.Pairs(ctx, bs, bs, i, npairs + i)
ket = capi.Pairs(ctx, bs, bs, 2 * npairs + i, 3 * npairs + i)
g = torch.Generator(device="cuda")
g.manual_seed(1)
tasks = torch.randint(0, npairs, (nq, 2), dtype=torch.int32, device="cuda", generator=g)
if len(sys.argv) > 7 and sys.argv[7] == "sorted":
    key = tasks[:, 0].long() * npairs + tasks[:, 1].long()
    tasks = tasks[torch.argsort(key)].contiguous()
blk = capi.eri_block_size(bra, ket)
out = torch.empty((nq, blk), dtype=torch.float64, device="cuda")
for r in range(reps):
    ctx.synchronize()
    t0 = time.perf_counter()
    capi.eri_batch(ctx, bra, ket, tasks, out=out)
    ctx.synchronize()
    dt = time.perf_counter() - t0
    print("class %s: %d quartets in %.3f ms -> %.3e q/s" % (cl, nq, dt * 1e3, nq / dt))
