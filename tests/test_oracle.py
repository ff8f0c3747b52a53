"""Pins the CPU oracle (reference Engine on restated kernels) against the reference's own
known answers for the Coulomb path.  CPU only."""
import itertools
import os

import numpy as np
import pytest

from util import all_classes, nc, random_shell_table

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _s_shell(po, ls):
    n = len(ls)
    return po.Shells(ls, [0] * n, [1] * n, np.zeros((n, 3)), [1.0] * n, [10.0] * n, raw=True)


def test_python_goldens(oracle):
    """python/tests/test_libint2.py:32-49: norms for s = Shell(0,[(1,10)]), p = Shell(1,[(1,10)])."""
    po = oracle
    assert np.linalg.norm(po.compute2(_s_shell(po, [1, 1, 0, 0]), precision=0.0)) == pytest.approx(
        1.62867503968, abs=5e-11)
    assert np.linalg.norm(po.compute2(_s_shell(po, [0, 0, 0]), braket=1, precision=0.0)) == pytest.approx(
        3.6563211198, abs=5e-11)
    ls = [0, 1, 0, 1]
    tot = sum(np.sum(po.compute2(_s_shell(po, [ls[i] for i in q]), precision=0.0) ** 2)
              for q in itertools.product(range(4), repeat=4))
    assert np.sqrt(tot) == pytest.approx(14.7036075402, abs=5e-10)


def test_shell_normalization_constant(oracle):
    """tests/unit/test-core.cc:52-55: Shell{{1},{{2,false,{1}}}} -> coeff 1.64592278064949."""
    c, _ = oracle.shell_renorm(2, [1.0], [1.0])
    assert c[0] == pytest.approx(1.64592278064949, abs=1e-14)


@pytest.mark.parametrize("cl", [c for c in all_classes(max_l=3, with_g=False) if sum(c) <= 8][::3])
def test_oracle_vs_closed_form(oracle, cl):
    """tests/eri/test.cc pattern: every Cartesian integral of a random contracted quartet vs
    the independent closed form eri() (eri.h:121-380); tolerance of tests/eri/test.cc:77-88
    (abs 5e-14 or rel 1e-9)."""
    po = oracle
    rng = np.random.default_rng(hash(cl) % 2 ** 31)
    K = 2
    l, pure, nprim, O, al, co = random_shell_table(rng, cl, K)
    got = po.compute2(po.Shells(l, pure, nprim, O, al, co, raw=False), precision=0.0)
    off = np.arange(4) * K
    # sample of components
    idx = [tuple(rng.integers(0, nc(x)) for x in cl) for _ in range(12)]
    for (ia, ib, ic, id_) in idx:
        def xyz(lx, i):
            k = 0
            for x in range(lx, -1, -1):
                for y in range(lx - x, -1, -1):
                    if k == i:
                        return [x, y, lx - x - y]
                    k += 1
        lmn = xyz(cl[0], ia) + xyz(cl[1], ib) + xyz(cl[2], ic) + xyz(cl[3], id_)
        ref = 0.0
        for p in itertools.product(range(K), repeat=4):
            a = [al[off[s] + p[s]] for s in range(4)]
            c = np.prod([co[off[s] + p[s]] for s in range(4)])
            ref += c * po.eri_closed(lmn, a, O.ravel(), 0)
        v = got[ia, ib, ic, id_]
        assert abs(v - ref) < 5e-14 or abs(v - ref) < 1e-9 * abs(ref), (cl, lmn, v, ref)


def test_oracle_matches_committed_goldens(oracle):
    """the committed fixtures are what this oracle produces (guards fixture drift)."""
    po = oracle
    d = np.load(os.path.join(GOLD, "eri_classes.npz"))
    for ci, cl in enumerate(d["classes"]):
        if ci % 7:
            continue
        K = int(d["c%d_K" % ci])
        sh = po.Shells(list(cl), [0] * 4, [K] * 4, d["c%d_O" % ci], d["c%d_alpha" % ci],
                       d["c%d_coeff" % ci], raw=False)
        np.testing.assert_array_equal(po.compute2(sh, precision=0.0).ravel(), d["c%d_eri" % ci])


def test_closed_form_goldens(oracle):
    d = np.load(os.path.join(GOLD, "closed_form.npz"))["rows"]
    for r in d[:50]:
        v = oracle.eri_closed(r[:12].astype(int), r[12:16], r[16:28], 0)
        assert v == r[28]


def test_boys_cheb7_accuracy(oracle):
    """tests/unit/test-core-ints.cc:78-85 contract (abs <= eps, rel <= 125 eps) against
    FmEval_Reference2 instead of MPFR."""
    d = np.load(os.path.join(GOLD, "boys.npz"))
    F, R = d["cheb7"], d["reference2"]
    eps = np.finfo(float).eps
    assert np.all((np.abs(F - R) <= eps) | (np.abs(F - R) <= 125 * eps * np.abs(R)))
    for i in (0, 100, 250, 405):
        np.testing.assert_array_equal(oracle.boys_cheb7(d["T"][i], 16, 24), F[i])


def test_permutation_symmetry(oracle):
    """tests/unit/test-permute.cc:116 / test-2body.cc:126-189: the 8 index permutations agree."""
    po = oracle
    rng = np.random.default_rng(3)
    cl = (2, 1, 1, 0)
    l, pure, nprim, O, al, co = random_shell_table(rng, cl, 2, pure=[1, 0, 0, 0])
    sh = po.Shells(l, pure, nprim, O, al, co, raw=False)
    base = po.compute2(sh, precision=0.0)
    for perm in [(1, 0, 2, 3), (0, 1, 3, 2), (2, 3, 0, 1), (3, 2, 1, 0), (2, 3, 1, 0)]:
        t = po.compute2(sh.subset(list(perm)), precision=0.0)
        np.testing.assert_allclose(t, base.transpose(perm), rtol=0, atol=1e-13)


def test_screening_bound(oracle):
    """tests/unit/test-precision.cc:82-216: |I_eps - I_0| <= 2 eps with Conservative screening."""
    po = oracle
    rng = np.random.default_rng(9)
    l, pure, nprim, O, al, co = random_shell_table(rng, (1, 0, 1, 0), 3, spread=2.5, amax=8.0)
    sh = po.Shells(l, pure, nprim, O, al, co, raw=False)
    ref = po.compute2(sh, precision=0.0, screening=po.SCREEN_CONSERVATIVE)
    for eps in (1e-8, 1e-10, 1e-12):
        got = po.compute2(sh, precision=eps, screening=po.SCREEN_CONSERVATIVE)
        if got is None:
            got = np.zeros_like(ref)
        assert np.max(np.abs(got - ref)) <= 2 * eps


def test_fock_golden(oracle):
    """oracle Fock build reproduces the committed G for H2O/cc-pVDZ independent of threads."""
    from libint_b200.basis import BasisSet, H2O_XYZ_ANGSTROM, atoms_from_tuples
    po = oracle
    d = np.load(os.path.join(GOLD, "fock_h2o.npz"))
    bs = BasisSet("cc-pvdz", atoms_from_tuples(H2O_XYZ_ANGSTROM))
    ns = len(bs)
    s1, s2 = np.array([(a, b) for a in range(ns) for b in range(a + 1)], dtype=np.int32).T
    f = po.Fock(po.Shells(*bs.flat(), raw=False), s1, s2, nthreads=2)
    G, st = f.build(d["ccpvdz_D"], 1e-12)
    np.testing.assert_allclose(G, d["ccpvdz_G"], rtol=1e-13, atol=1e-14)
    # two task-stride halves sum to the whole (how bench.py samples the CPU baseline)
    G0, _ = f.build(d["ccpvdz_D"], 1e-12, task_stride=2, task_offset=0)
    G1, _ = f.build(d["ccpvdz_D"], 1e-12, task_stride=2, task_offset=1)
    np.testing.assert_allclose(G0 + G1, G, rtol=1e-12, atol=1e-13)
