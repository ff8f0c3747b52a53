"""Algorithmic FP64 flop model of one (la lb|lc ld) shell quartet.

Head-Gordon-Pople scheme as the reference's generated kernels use it
(LIBINT_ERI_STRATEGY 1): per primitive quartet
    F_prim = 90 (prerequisites, engine.impl.h:1331-1701) + 16(L+1)+6 (Boys, boys.h:434-442)
             + (L+1) (scaling by pfac, engine.impl.h:1510-1512) + VRR + N_tgt (contraction adds)
and per contracted quartet F_hrr.  VRR enumerates every Cartesian component of every
needed (e0|f0)^(m) with the per-term costs of src/bin/libint/vrr_11_twoprep_11.h (3 flops for
the two leading terms, +5 when the a-2 / c-2 term exists, +3 when the cross term exists),
building on C when f > 0 else on A, along the first nonzero direction
(src/lib/libint/OSVRR_xs_xs.h:74-78).  HRR = 2 flops per produced Cartesian integral
(src/bin/libint/hrr.h:246,324), ket side first, then bra.  The same table is used for the GPU
and the CPU numbers; it is a model of the algorithm, not a count of executed instructions.
"""
import functools


def nc(l):
    return (l + 1) * (l + 2) // 2


def _components(l):
    return [(x, y, l - x - y) for x in range(l, -1, -1) for y in range(l - x, -1, -1)]


@functools.lru_cache(maxsize=None)
def vrr_flops(la, lb, lc, ld):
    emax, fmax = la + lb, lc + ld
    need = {}  # (qa, qc) -> set of m for which [qa 0|qc 0]^(m) is needed (component-level closure)
    stack = []

    def require(qa, qc, m):
        s = need.setdefault((qa, qc), set())
        if m not in s:
            s.add(m)
            stack.append((qa, qc, m))

    def dec(q, d, k=1):
        q = list(q)
        q[d] -= k
        return tuple(q)

    for e in range(la, emax + 1):
        for f in range(lc, fmax + 1):
            for qa in _components(e):
                for qc in _components(f):
                    require(qa, qc, 0)
    flops = 0
    while stack:
        qa, qc, m = stack.pop()
        if sum(qc) > 0:      # build on C (vrr_11_twoprep_11.h:305-383)
            d = 0 if qc[0] else (1 if qc[1] else 2)
            c = 3
            require(qa, dec(qc, d), m)
            require(qa, dec(qc, d), m + 1)
            if qc[d] > 1:
                c += 5
                require(qa, dec(qc, d, 2), m)
                require(qa, dec(qc, d, 2), m + 1)
            if qa[d] > 0:
                c += 3
                require(dec(qa, d), dec(qc, d), m + 1)
            flops += c
        elif sum(qa) > 0:    # build on A (vrr_11_twoprep_11.h:154-222)
            d = 0 if qa[0] else (1 if qa[1] else 2)
            c = 3
            require(dec(qa, d), qc, m)
            require(dec(qa, d), qc, m + 1)
            if qa[d] > 1:
                c += 5
                require(dec(qa, d, 2), qc, m)
                require(dec(qa, d, 2), qc, m + 1)
            flops += c
    return flops


def n_targets(la, lb, lc, ld):
    return sum(nc(e) for e in range(la, la + lb + 1)) * sum(nc(f) for f in range(lc, lc + ld + 1))


@functools.lru_cache(maxsize=None)
def hrr_flops(la, lb, lc, ld):
    ne = sum(nc(e) for e in range(la, la + lb + 1))
    flops = 0
    # ket: level y = 1..ld produces (e0|x y) for x in [lc, lc+ld-y], for all bra target rows
    for y in range(1, ld + 1):
        flops += 2 * ne * sum(nc(x) * nc(y) for x in range(lc, lc + ld - y + 1))
    ncd = nc(lc) * nc(ld)
    for y in range(1, lb + 1):
        flops += 2 * ncd * sum(nc(x) * nc(y) for x in range(la, la + lb - y + 1))
    return flops


def prim_flops(la, lb, lc, ld):
    L = la + lb + lc + ld
    return 90 + 16 * (L + 1) + 6 + (L + 1) + vrr_flops(la, lb, lc, ld) + n_targets(la, lb, lc, ld)


def canonical(la, lb, lc, ld):
    """class as the reference builds it: la>=lb, lc>=ld, la+lb<=lc+ld (build_libint.cc:78-83)."""
    if la < lb:
        la, lb = lb, la
    if lc < ld:
        lc, ld = ld, lc
    if la + lb > lc + ld:
        la, lb, lc, ld = lc, ld, la, lb
    return la, lb, lc, ld


def quartet_flops(la, lb, lc, ld, nprim_quartets=1):
    """algorithmic flops of one contracted shell quartet with `nprim_quartets` surviving
    primitive quartets (class canonicalised first: the flop model is orientation-specific)."""
    la, lb, lc, ld = canonical(la, lb, lc, ld)
    return nprim_quartets * prim_flops(la, lb, lc, ld) + hrr_flops(la, lb, lc, ld)


def canonical_classes(lmax):
    out = []
    for la in range(lmax + 1):
        for lb in range(la + 1):
            for lc in range(lmax + 1):
                for ld in range(lc + 1):
                    if la + lb <= lc + ld and not (la + lb == lc + ld and (la, lb) > (lc, ld) and False):
                        out.append((la, lb, lc, ld))
    return out


if __name__ == "__main__":
    for cl in [(0, 0, 0, 0), (0, 0, 1, 0), (1, 0, 1, 0), (1, 1, 1, 1), (2, 0, 2, 0), (2, 1, 2, 1),
               (2, 2, 2, 2), (3, 0, 3, 0), (3, 1, 3, 1), (3, 2, 3, 2), (3, 3, 3, 3)]:
        print(cl, "VRR", vrr_flops(*cl), "Ntgt", n_targets(*cl), "HRR", hrr_flops(*cl), "total K=1",
              quartet_flops(*cl))
