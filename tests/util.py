"""Shared helpers of the parity tests."""
import itertools

import numpy as np

# tolerance BASELINE.json's north_star states for integrals and Fock matrices
RTOL, ATOL = 1e-12, 1e-14

PAIR_CLASSES = [(a, b) for a in range(4) for b in range(a + 1)] + [(4, 0)]


def pair_key(c):
    return (c[0] + c[1]) * 100 + c[0] * 10 + c[1]


def all_classes(max_l=3, with_g=True):
    pcs = [pc for pc in PAIR_CLASSES if pc[0] <= max_l or (with_g and pc == (4, 0))]
    return [x + y for x, y in itertools.product(pcs, pcs)]


def nc(l):
    return (l + 1) * (l + 2) // 2


def random_shell_table(rng, ls, K, pure=None, spread=1.0, amin=0.2, amax=3.0):
    """(l, pure, nprim, O, alpha, coeff): random contracted shells in the spirit of the
    reference's RandomShellSet (src/bin/test_eri/prep_libint2.h:35-69)."""
    n = len(ls)
    nprim = [K] * n if np.isscalar(K) else list(K)
    O = rng.uniform(-spread, spread, (n, 3))
    al = rng.uniform(amin, amax, sum(nprim))
    co = rng.uniform(0.2, 1.5, sum(nprim))
    return list(ls), list(pure) if pure is not None else [0] * n, nprim, O, al, co


def assert_parity(got, ref, what="", rtol=RTOL, atol=ATOL):
    got = np.asarray(got, dtype=np.float64).ravel()
    ref = np.asarray(ref, dtype=np.float64).ravel()
    assert got.shape == ref.shape, "%s: shape %s vs %s" % (what, got.shape, ref.shape)
    err = np.abs(got - ref)
    bad = err > atol + rtol * np.abs(ref)
    if bad.any():
        i = int(np.argmax(err - (atol + rtol * np.abs(ref))))
        raise AssertionError("%s: %d of %d elements outside %g rel / %g abs; worst at %d: got %.17g "
                             "ref %.17g (abs err %.3g)" % (what, bad.sum(), len(ref), rtol, atol, i,
                                                           got[i], ref[i], err[i]))


def hrr_amplification(l, O):
    """Conditioning of the horizontal recurrence for one shell set: the HGP scheme forms
    (a b| = sum_k C(lb,k) AB^k (a+lb-k 0| (src/bin/libint/hrr.h:246,324), so rounding noise of
    the contracted (e0|f0) intermediates is amplified by up to (1+|AB|)^lb (1+|CD|)^ld in the
    final integrals -- for the reference's generated code as for ours (tests/eri/test.cc:77-83
    notes the same loss for (dp|dd), (dd|dd)).  The absolute tolerance of a parity check between
    two different operation orders is scaled by this factor."""
    O = np.asarray(O, dtype=np.float64).reshape(-1, 3)
    ab = np.linalg.norm(O[0] - O[1])
    amp = (1.0 + ab) ** min(l[0], l[1])
    if len(l) == 4:
        cd = np.linalg.norm(O[2] - O[3])
        amp *= (1.0 + cd) ** min(l[2], l[3])
    return amp
