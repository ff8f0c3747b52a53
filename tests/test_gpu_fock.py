"""Parity of the fused ERI + J/K digestion path with compute_2body_fock of the reference's
direct-SCF driver (tests/hartree-fock/hartree-fock++.cc:1574-1772), through the C ABI."""
import os

import numpy as np
import pytest

from util import assert_parity

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _h2o(name):
    from libint_b200.basis import BasisSet, H2O_XYZ_ANGSTROM, atoms_from_tuples
    return BasisSet(name, atoms_from_tuples(H2O_XYZ_ANGSTROM))


def _sym_density(n, seed=11, scale=0.3):
    rng = np.random.default_rng(seed)
    D = rng.standard_normal((n, n)) * scale
    return 0.5 * (D + D.T)


@pytest.mark.parametrize("name", ["sto-3g", "6-31g", "cc-pvdz"])
def test_fock_h2o_vs_golden(ctx, name):
    """BASELINE config 1's Fock build (H2O, h2o.xyz): G and the Schwarz matrix
    (hartree-fock++.cc:1230-1298) against the committed oracle output."""
    from libint_b200 import capi
    d = np.load(os.path.join(GOLD, "fock_h2o.npz"))
    tag = name.replace("-", "")
    bs = _h2o(name)
    B = capi.Basis(ctx, *bs.flat())
    f = capi.Fock(ctx, B)
    assert len(f.pair_s1) == len(bs) * (len(bs) + 1) // 2  # every H2O pair is significant
    np.testing.assert_allclose(f.schwarz(), d[tag + "_K"], rtol=1e-12, atol=1e-15)
    G, st = f.build(d[tag + "_D"], 1e-12, stats=True)
    assert_parity(G, d[tag + "_G"], "G H2O/" + name, rtol=1e-12, atol=2e-14)
    assert st["nquartets"] == float(d[tag + "_nquartets"])
    assert np.array_equal(G, G.T)


def _cluster_case(ctx, oracle, atoms, name, precision, nthreads=8):
    from libint_b200 import capi
    from libint_b200.basis import BasisSet
    bs = BasisSet(name, atoms)
    B = capi.Basis(ctx, *bs.flat())
    f = capi.Fock(ctx, B)
    of = oracle.Fock(oracle.Shells(*bs.flat(), raw=False), f.pair_s1, f.pair_s2, nthreads=nthreads)
    D = _sym_density(bs.nbf, seed=5, scale=0.1)
    G, st = f.build(D, precision, stats=True)
    Gref, ost = of.build(D, precision)
    return bs, f, of, D, G, Gref, st, ost


def test_fock_water_dimer_screened(ctx, oracle):
    """two waters 6 A apart, cc-pVDZ: significant-pair list, Schwarz x density screening and
    SchwarzInf primitive screening all active; same quartets computed as the reference."""
    from libint_b200.basis import Atom, H2O_XYZ_ANGSTROM, atoms_from_tuples
    a = atoms_from_tuples(H2O_XYZ_ANGSTROM)
    atoms = a + [Atom(x.atomic_number, x.x + 11.3, x.y + 0.4, x.z - 0.2) for x in a]
    bs, f, of, D, G, Gref, st, ost = _cluster_case(ctx, oracle, atoms, "cc-pvdz", 1e-10)
    ns = len(bs)
    assert len(f.pair_s1) < ns * (ns + 1) // 2  # some pairs were dropped by the overlap screen
    # the library evaluates K only for the significant pairs (the only ones the build reads)
    mask = np.zeros((ns, ns), dtype=bool)
    mask[f.pair_s1, f.pair_s2] = mask[f.pair_s2, f.pair_s1] = True
    np.testing.assert_allclose(f.schwarz()[mask], of.schwarz()[mask], rtol=1e-12, atol=1e-15)
    assert np.all(f.schwarz()[~mask] == 0)
    assert st["nquartets"] == ost["nquartets"]
    assert_parity(G, Gref, "G water dimer", rtol=1e-12, atol=2e-14)
    # screening error itself is bounded by the requested precision
    G0 = f.build(D, 1e-14, use_schwarz=False)
    assert np.max(np.abs(G - G0)) < 1e-8


def test_fock_with_f_shells(ctx, oracle):
    """def2-TZVP water (f on O): every class up to (ff|ff) occurs in one build."""
    from libint_b200 import capi
    from libint_b200.basis import H2O_XYZ_ANGSTROM, atoms_from_tuples
    if not capi.eri_class_supported(3, 3, 3, 3):
        pytest.xfail("(ff|ff) kernel not built yet")
    bs, f, of, D, G, Gref, st, ost = _cluster_case(ctx, oracle, atoms_from_tuples(H2O_XYZ_ANGSTROM),
                                                   "def2-tzvp", 1e-12)
    assert st["nquartets"] == ost["nquartets"]
    assert_parity(G, Gref, "G H2O/def2-TZVP", rtol=1e-12, atol=5e-14)


def test_primitive_pair_data_schwarzinf(ctx, oracle):
    """ScreeningMethod::SchwarzInf shell-pair data (hartree-fock++.cc:1383-1431,
    shell.h:1259-1328): the library's GPU-evaluated primitive Schwarz factors give the same
    surviving primitive pairs and ln_scr as the reference's evaluator."""
    from libint_b200 import capi
    bs = _h2o("cc-pvdz")
    B = capi.Basis(ctx, *bs.flat())
    ns = len(bs)
    s1, s2 = np.array([(a, b) for a in range(ns) for b in range(a + 1)], dtype=np.int32).T
    of = oracle.Fock(oracle.Shells(*bs.flat(), raw=False), s1, s2, nthreads=2)
    lnp = np.log(np.finfo(float).eps / 1e10)
    l = np.array([s.l for s in bs])
    for a, b in [(0, 0), (3, 0), (5, 3), (5, 5), (8, 1), (11, 5)]:
        hi, lo = (a, b) if l[a] >= l[b] else (b, a)
        P = capi.Pairs(ctx, B, B, [hi], [lo], screening=capi.SCREEN_SCHWARZ_INF, ln_prec=lnp)
        got = P.get(0)
        ref = of.pairdata(a, b)
        if (hi, lo) != (a, b):  # reference stores (s1,s2) with s1 >= s2 by index; swap p1/p2
            ref = ref[:, [0, 1, 2, 3, 4, 5, 6, 8, 7]]
            ref = ref[np.lexsort((ref[:, 8], ref[:, 7]))]
        assert got.shape == ref.shape, (a, b)
        np.testing.assert_allclose(got[:, :5], ref[:, :5], rtol=2e-15, atol=1e-300)
        np.testing.assert_allclose(got[:, 6], ref[:, 6], rtol=1e-12, atol=1e-12)
        np.testing.assert_array_equal(got[:, 7:], ref[:, 7:])


def test_rank_partition_sums_to_whole(ctx):
    """north_star multi-GPU rule: the partial G's of the N ranks add up to the 1-rank G."""
    from libint_b200 import capi
    bs = _h2o("cc-pvdz")
    B = capi.Basis(ctx, *bs.flat())
    f = capi.Fock(ctx, B)
    D = _sym_density(bs.nbf)
    G = f.build(D, 1e-12)
    for nranks in (2, 3, 8):
        parts = [f.build(D, 1e-12, rank=r, nranks=nranks, stats=True) for r in range(nranks)]
        Gs = sum(p[0] for p in parts)
        assert sum(p[1]["nquartets"] for p in parts) == f.build(D, 1e-12, stats=True)[1]["nquartets"]
        assert_parity(Gs, G, "sum of %d partial G" % nranks, rtol=1e-12, atol=2e-14)


def test_device_buffers_and_linearity(ctx):
    """torch CUDA tensors in/out give the host-buffer result; G is linear in D when
    screening is off (a size-independent property used at full benchmark sizes)."""
    import torch
    from libint_b200 import capi
    from libint_b200.fock import FockBuilder
    bs = _h2o("6-31g")
    fb = FockBuilder(bs, ctx=ctx)
    D1, D2 = _sym_density(bs.nbf, 1), _sym_density(bs.nbf, 2)
    G1 = fb.fock.build(D1, 1e-14, use_schwarz=False)
    G2 = fb.fock.build(D2, 1e-14, use_schwarz=False)
    G12 = fb.fock.build(D1 + 2 * D2, 1e-14, use_schwarz=False)
    assert_parity(G12, G1 + 2 * G2, "linearity", rtol=1e-11, atol=1e-12)
    Gt = fb(torch.from_numpy(D1).cuda(), precision=1e-14, use_schwarz=False)
    assert Gt.is_cuda and Gt.dtype == torch.float64
    assert_parity(Gt.cpu().numpy(), G1, "device buffers", rtol=1e-12, atol=2e-14)


def _purity_case(ctx, oracle, bs, what):
    from libint_b200 import capi
    B = capi.Basis(ctx, *bs.flat())
    assert B.nbf == bs.nbf
    f = capi.Fock(ctx, B)
    of = oracle.Fock(oracle.Shells(*bs.flat(), raw=False), f.pair_s1, f.pair_s2, nthreads=4)
    D = _sym_density(bs.nbf, seed=3, scale=0.2)
    G, st = f.build(D, 1e-12, stats=True)
    Gref, ost = of.build(D, 1e-12)
    assert st["nquartets"] == ost["nquartets"]
    np.testing.assert_allclose(f.schwarz(), of.schwarz(), rtol=1e-12, atol=1e-15)
    assert_parity(G, Gref, what, rtol=1e-12, atol=2e-14)
    assert np.array_equal(G, G.T)


def test_fock_cartesian_d_basis(ctx, oracle):
    """6-31G* keeps Cartesian d shells (Gaussian convention, basis.h.in:368-386): the digestion
    must honour the per-shell purity flag, not assume pure-iff-l>=2 (engine.impl.h:1965-1985)."""
    bs = _h2o("6-31g*")
    assert any(s.l == 2 and not s.pure for s in bs)
    _purity_case(ctx, oracle, bs, "G H2O/6-31G* (Cartesian d)")


def test_fock_set_pure_false(ctx, oracle):
    """BasisSet::set_pure(false) (basis.h.in:165-171) on cc-pVDZ: every shell Cartesian."""
    bs = _h2o("cc-pvdz")
    bs.set_pure(False)
    assert bs.nbf == 25
    _purity_case(ctx, oracle, bs, "G H2O/cc-pVDZ set_pure(false)")


def test_fock_pure_p_shell_and_mixed(ctx, oracle):
    """set_pure(true) makes the p shells solid harmonics too (order y,z,x, solidharmonics.h:114-174);
    then a mixed pattern: one Cartesian d next to pure ones."""
    bs = _h2o("cc-pvdz")
    bs.set_pure(True)
    _purity_case(ctx, oracle, bs, "G H2O/cc-pVDZ set_pure(true)")
    from libint_b200.basis import Atom, H2O_XYZ_ANGSTROM, atoms_from_tuples, BasisSet
    a = atoms_from_tuples(H2O_XYZ_ANGSTROM)
    atoms = a + [Atom(x.atomic_number, x.x + 3.1, x.y + 0.4, x.z - 0.2) for x in a]
    bs = BasisSet("cc-pvdz", atoms)
    flip = [i for i, s in enumerate(bs) if s.l == 2][0]
    bs[flip].pure = False
    first_p = [i for i, s in enumerate(bs) if s.l == 1][0]
    bs[first_p].pure = True
    bs._refresh()
    _purity_case(ctx, oracle, bs, "G (H2O)2/cc-pVDZ mixed purity")


def test_significant_pairs_device_matches_host(ctx):
    """compute_shellpairs (hartree-fock++.cc:1305-1381) on the GPU == the host loop, for a cluster large
    enough that most pairs are dropped, Cartesian and pure shells."""
    from libint_b200 import capi
    from libint_b200.basis import BasisSet, water_cluster
    for name, pure in (("def2-tzvp", None), ("cc-pvdz", False)):
        bs = BasisSet(name, water_cluster(3, 2, 2))
        if pure is not None:
            bs.set_pure(pure)
        B = capi.Basis(ctx, *bs.flat())
        for thr in (1e-12, 1e-8):
            d1, d2 = capi.significant_pairs(B, thr, device=True)
            h1, h2 = capi.significant_pairs(B, thr, device=False)
            assert len(d1) < B.nshell * (B.nshell + 1) // 2
            assert np.array_equal(d1, h1) and np.array_equal(d2, h2)
