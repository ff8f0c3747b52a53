"""Time store-mode launches of a few classes with several experiment libraries (scripts/build_variant.py) in one
process: python scripts/variants_once.py "_v0,_v1,..." "2222,2122" [nquartets] [reps].  Prints the best time per
(variant, class) and the largest deviation of each variant's integrals from the first variant's."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import torch  # noqa: E402

from bench import class_table  # noqa: E402
from libint_b200 import capi  # noqa: E402

sufs = sys.argv[1].split(",")
classes = [tuple(int(x) for x in c) for c in sys.argv[2].split(",")]
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 20
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
npairs = 4096
ref = {}
base_path = capi.LIB_PATH
for suf in sufs:
    capi._lib = None
    capi.LIB_PATH = base_path.replace("liblibint_b200.so", "liblibint_b200%s.so" % suf)
    ctx = capi.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    for cl in classes:
        tab = class_table(cl, npairs, 0)
        bs = capi.Basis(ctx, *tab)
        i = np.arange(npairs, dtype=np.int32)
        bra = capi.Pairs(ctx, bs, bs, i, npairs + i)
        ket = capi.Pairs(ctx, bs, bs, 2 * npairs + i, 3 * npairs + i)
        g = torch.Generator(device="cuda")
        g.manual_seed(1)
        tasks = torch.randint(0, npairs, (nq, 2), dtype=torch.int32, device="cuda", generator=g)
        blk = capi.eri_block_size(bra, ket)
        out = torch.empty((nq, blk), dtype=torch.float64, device="cuda")
        best = 1e30
        for r in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            capi.eri_batch(ctx, bra, ket, tasks, out=out)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        chk = out[:65536].clone()
        dev = 0.0 if cl not in ref else float((chk - ref[cl]).abs().max())
        ref.setdefault(cl, chk)
        print("variant %-4s class %s: %.3f ms / %d quartets  (max |x - x_first| = %.2e)" % (suf, "".join(map(str, cl)), best, nq, dev),
              flush=True)
        del out, tasks, bra, ket, bs
    ctx.close()
