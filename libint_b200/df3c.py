"""Three-centre Coulomb integrals (P|mu nu) for density fitting, batched by class.

Reference: the DF set-up of the direct-SCF driver computes Zxy[P][mu][nu] with one
Engine(Operator::coulomb, BraKet::xs_xx).compute2(dfbs[s1], Shell::unit(), obs[s2], obs[s3]) per
shell triplet on a thread pool (tests/hartree-fock/hartree-fock++.cc:2215-2262;
engine.impl.h:1836-1873 for the xs_xx canonicalisation).  Here the triplets are grouped by
class (L s|lc ld): the bra block holds every DF shell of angular momentum L paired with the
unit shell, the ket block every significant orbital shell pair of class (lc ld)
(compute_shellpairs, hartree-fock++.cc:1305-1381), and one lb200_eri_batch call evaluates a
chunk of the Cartesian product.  The dense tensor (107 GB for C40H82 / def2-TZVP /
def2-universal-JKFIT) is never materialised: the caller consumes each chunk (`sink`) while it
is resident in HBM -- exactly what the DF Fock builder of SURVEY 8(f)2 will do.
"""
import numpy as np

from . import capi


class ThreeCenter:
    def __init__(self, ctx, obs, dfbs, pair_threshold=1e-12):
        self.ctx, self.obs, self.dfbs = ctx, obs, dfbs
        self.B = capi.Basis(ctx, *obs.flat())
        self.Bdf = capi.Basis(ctx, *dfbs.flat())
        self.unit = capi.Basis.unit(ctx)
        s1, s2 = capi.significant_pairs(self.B, pair_threshold)
        lo = np.array([s.l for s in obs])
        # ket blocks: class (lc ld), first shell = higher AM (Pairs requires la >= lb)
        a, b = np.array(s1), np.array(s2)
        sw = lo[a] < lo[b]
        a, b = np.where(sw, b, a), np.where(sw, a, b)
        # blocks are split by contraction as well (uncontracted | contracted), like the Fock
        # builder's buckets: the quartets of a launch run their primitive loops in lockstep, and
        # all-uncontracted launches take the pipelined kernel
        npo = np.array([s.nprim for s in obs])
        cb = (npo[a] * npo[b] > 1).astype(int)
        self.kets = {}
        for key in sorted(set(zip(lo[a].tolist(), lo[b].tolist(), cb.tolist()))):
            m = (lo[a] == key[0]) & (lo[b] == key[1]) & (cb == key[2])
            self.kets[key] = capi.Pairs(ctx, self.B, self.B, a[m], b[m])
        ldf = np.array([s.l for s in dfbs])
        cdf = (np.array([s.nprim for s in dfbs]) > 1).astype(int)
        self.bras = {}
        for key in sorted(set(zip(ldf.tolist(), cdf.tolist()))):
            idx = np.nonzero((ldf == key[0]) & (cdf == key[1]))[0].astype(np.int32)
            self.bras[key] = capi.Pairs(ctx, self.Bdf, self.unit, idx, np.zeros_like(idx))
        self.npairs = int(len(a))
        self.pair_a, self.pair_b = a.astype(np.int32), b.astype(np.int32)   # first shell = higher AM

    def blocks(self):
        """(bra key, ket key) of every launch group; bra key = (L, contracted), ket key =
        (lc, ld, contracted)."""
        return [(kb, kk) for kb in self.bras for kk in self.kets]

    def classes(self):
        """distinct angular-momentum classes (L s|lc ld)."""
        return sorted(set((kb[0], kk[0], kk[1]) for kb, kk in self.blocks()))

    def ntriplets(self):
        return sum(self.bras[kb].npair * self.kets[kk].npair for kb, kk in self.blocks())

    def sweep(self, out, chunk_bytes=1 << 30, sink=None, events=None, rank=0, nranks=1):
        """Every (P|mu nu) shell triplet once, class by class, Cartesian, into the device
        buffer `out` (torch CUDA float64, reused chunk after chunk).  `sink((bra key, ket key), t0, n,
        view)` sees each finished chunk.  With nranks > 1 this rank takes its contiguous share of the DF
        shells of every bra block (the tensor shards by P, no collective).  Returns the number of shell
        triplets this rank computed."""
        import torch
        dev = out.device
        # the kernels run on the context's stream: make it torch's current one, so that `out` is not
        # overwritten while a sink on the torch stream still reads the previous chunk
        self.ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
        total = 0
        for kb, kk in self.blocks():
            cls = (kb[0], kk[0], kk[1])
            bra, ket = self.bras[kb], self.kets[kk]
            blk = capi.eri_block_size(bra, ket)
            r_lo, r_hi = bra.npair * rank // nranks, bra.npair * (rank + 1) // nranks
            n = (r_hi - r_lo) * ket.npair
            per = max(1, min(max(n, 1), min(out.numel(), chunk_bytes // 8) // blk))
            if events is not None:
                e0 = torch.cuda.Event(enable_timing=True)
                e0.record(torch.cuda.current_stream(dev))
            # whole bra rows per launch: an implicit (DF shells) x (orbital pairs) product, no task list
            rows = max(1, per // ket.npair)
            for b0 in range(r_lo, r_hi, rows):
                nb = min(rows, r_hi - b0)
                m = nb * ket.npair
                if m * blk > out.numel():   # one row does not fit: split it over ket ranges
                    kper = max(1, out.numel() // blk)
                    for k0 in range(0, ket.npair, kper):
                        nk = min(kper, ket.npair - k0)
                        capi.eri_product(self.ctx, bra, ket, b0, 1, k0, nk, out)
                        if sink is not None:
                            sink((kb, kk), b0 * ket.npair + k0, nk, out[:nk * blk].view(nk, blk))
                    continue
                capi.eri_product(self.ctx, bra, ket, b0, nb, 0, ket.npair, out)
                if sink is not None:
                    sink((kb, kk), b0 * ket.npair, m, out[:m * blk].view(m, blk))
            if events is not None:
                e1 = torch.cuda.Event(enable_timing=True)
                e1.record(torch.cuda.current_stream(dev))
                events.append((cls, n, blk, e0, e1))
            total += n
        return total
