"""The reference's own plugin boundary on the GPU library:
liblibint_b200_iface.so exports Libint_t + libint2_build_eri/_3eri/_2eri + libint2_static_init /
libint2_{need_memory,init,cleanup}_* (src/bin/libint/iface.cc:114-185,302-418), and the reference's
UNMODIFIED header-only libint2::Engine / Shell / ShellPair (compiled from /root/reference in the build
container: oracle/_ref/librefengine_b200.so = oracle_capi.cc against include/libint2/util/generated)
runs on top of it.  Compared with the CPU oracle (same Engine on the restated CPU kernels)."""
import os
import subprocess

import numpy as np
import pytest

from util import all_classes, assert_close_to_oracle, random_shell_table

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_api_port_vs_engine(oracle):
    """port of tests/unit/c-api.c + test-c-api.cc: a C program fills Libint_t itself and calls
    libint2_build_eri[a][b][c][d]; results x normalization == Engine results."""
    exe = os.path.join(ROOT, "oracle", "_ref", "c_api_port")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = r.stdout.split("\n")
    assert lines[-2] == "done"
    cen = np.array([[0.0, 1.0, 2.0], [1.0, 2.0, 0.0], [2.0, 0.0, 1.0], [0.0, 1.0, 2.0]])
    al = np.array([[1.1, 0.4], [2.3, 0.7], [3.4, 0.9], [4.8, 0.6]])
    co = np.array([[1.0, 0.5], [1.0, 0.8], [1.0, 0.3], [1.0, 0.6]])
    i, nclass = 0, 0
    while lines[i].startswith("class"):
        t = lines[i].split()
        cl, n = [int(x) for x in t[1:5]], int(t[6])
        got = np.array([float(x) for x in lines[i + 1:i + 1 + n]])
        i += 1 + n
        # the C program applies raw coefficients; the Engine's shells embed the normalization of a
        # primitive of that l and exponent (Shell::renorm without unit normalization) -- give the oracle
        # the same raw numbers as "already normalized" coefficients
        sh = oracle.Shells(cl, [0] * 4, [2] * 4, cen, al.ravel(), co.ravel(), raw=False)
        ref = oracle.compute2(sh, precision=0.0).ravel()
        assert_close_to_oracle(got, ref, sh, [0, 1, 2, 3], "C API class %s" % cl)
        nclass += 1
    assert nclass == 21   # canonical classes l <= 2 without (ss|ss)


@pytest.mark.parametrize("K", [1, 3])
def test_reference_engine_on_gpu_library(oracle, K):
    """Engine::compute2<coulomb, xx_xx> of the reference, linked to the GPU library, for every class
    l <= 3 in canonical and non-canonical shell orders (tests/unit/test-permute.cc:116)."""
    rng = np.random.default_rng(4242 + K)
    n = 0
    for cl in all_classes(max_l=3, with_g=False):
        if K == 3 and sum(cl) > 8:
            continue
        if n % 3 and sum(cl) > 6:   # thin out the big classes: one PCIe round trip per shell set
            n += 1
            continue
        n += 1
        table = random_shell_table(rng, cl, K)
        sh = oracle.Shells(*table, raw=False)
        got = oracle.compute2(sh, precision=0.0, b200=True)
        ref = oracle.compute2(sh, precision=0.0)
        assert_close_to_oracle(got, ref, sh, [0, 1, 2, 3], "refengine/GPU class %s K=%d" % (cl, K))


def test_reference_engine_pure_3center_2center(oracle):
    rng = np.random.default_rng(99)
    table = random_shell_table(rng, (3, 2, 2, 1), 2, pure=[1, 1, 1, 0])
    sh = oracle.Shells(*table, raw=False)
    for perm in [(0, 1, 2, 3), (1, 0, 3, 2), (2, 3, 0, 1), (3, 2, 1, 0)]:
        s = sh.subset(list(perm))
        got = oracle.compute2(s, precision=0.0, b200=True)
        ref = oracle.compute2(s, precision=0.0)
        assert_close_to_oracle(got, ref, s, [0, 1, 2, 3], "pure perm %s" % (perm,))
    t3 = random_shell_table(rng, (4, 3, 2), 2, pure=[1, 1, 1])
    s3 = oracle.Shells(*t3, raw=False)
    got = oracle.compute2(s3, braket=1, precision=0.0, b200=True)
    ref = oracle.compute2(s3, braket=1, precision=0.0)
    np.testing.assert_allclose(got, ref, rtol=1e-11, atol=1e-13)
    t2 = random_shell_table(rng, (4, 3), 2, pure=[1, 0])
    s2 = oracle.Shells(*t2, raw=False)
    got = oracle.compute2(s2, braket=2, precision=0.0, b200=True)
    ref = oracle.compute2(s2, braket=2, precision=0.0)
    np.testing.assert_allclose(got, ref, rtol=1e-11, atol=1e-13)
    # default precision (primitive screening inside the reference Engine) and an s-only quartet
    t0 = random_shell_table(rng, (0, 0, 0, 0), 3)
    s0 = oracle.Shells(*t0, raw=False)
    np.testing.assert_allclose(oracle.compute2(s0, b200=True), oracle.compute2(s0), rtol=1e-13)


def test_reference_engine_first_derivatives_on_gpu_library(oracle):
    """Engine(Operator::coulomb, max_nprim, max_l, deriv_order = 1).compute2<coulomb, xx_xx, 1> of the reference,
    UNMODIFIED, on libint2_build_eri1 of the GPU library: twelve shell sets per quartet in canonical and permuted
    shell orders (the Engine re-maps the derivative index, engine.impl.h:1996-2003), pure and Cartesian shells,
    against the closed-form derivative integrals at the reference's own thresholds (tests/eri/test.cc:77-88,434-437)."""
    from util import _sph_matrix, nc
    rng = np.random.default_rng(515)
    cases = [((0, 0, 0, 0), None), ((1, 0, 0, 0), None), ((0, 1, 1, 0), None), ((1, 1, 1, 1), None),
             ((0, 2, 1, 1), (0, 1, 0, 0)), ((2, 1, 0, 2), (1, 0, 0, 1)), ((1, 2, 2, 0), (0, 0, 1, 0)),
             ((2, 2, 1, 2), (1, 1, 0, 1)), ((2, 2, 2, 2), (1, 0, 1, 0))]
    for ls, pure in cases:
        K = 1 if sum(ls) >= 7 else 2
        l, pu, nprim, O, al, co = random_shell_table(rng, ls, K, pure=pure)
        sh = oracle.Shells(l, pu, nprim, O, al, co, raw=False)
        got = oracle.compute2_deriv1(sh, precision=0.0)
        assert got is not None and got.shape[0] == 12
        ref = oracle.deriv1_closed(oracle.Shells(l, [0] * 4, nprim, O, al, co, raw=False)).reshape([12] + [nc(x) for x in ls])
        for ax in range(4):
            if pu[ax] and ls[ax] > 0:
                M = _sph_matrix(oracle, ls[ax])
                ref = np.moveaxis(np.tensordot(M, ref, axes=([1], [ax + 1])), 0, ax + 1)
        ref = ref.reshape(12, -1)
        err = np.abs(got - ref)
        bad = (err > 1e-9 * np.abs(ref)) & (err > 5e-14)
        assert not bad.any(), (ls, err.max())
        assert err.max() <= 1e-11 * max(1.0, np.abs(ref).max()), (ls, err.max())
    # l = 3 is beyond LIBINT2_MAX_AM_eri1 = 2 of the GPU library: lmax_exceeded, as on a libint built that way
    t = random_shell_table(rng, (3, 0, 0, 0), 1)
    with pytest.raises(RuntimeError):
        oracle.compute2_deriv1(oracle.Shells(*t, raw=False), precision=0.0)


def test_reference_engine_lmax_exceeded(oracle):
    """four-centre g shells are beyond LIBINT2_MAX_AM_eri of the GPU library: the reference Engine
    reports lmax_exceeded (engine.h:893-916) instead of calling a null table entry."""
    rng = np.random.default_rng(3)
    table = random_shell_table(rng, (4, 0, 0, 0), 1)
    sh = oracle.Shells(*table, raw=False)
    with pytest.raises(RuntimeError):
        oracle.compute2(sh, precision=0.0, b200=True)


def test_reference_fock_driver_on_gpu_library(oracle):
    """the reference's direct-SCF consumer (compute_2body_fock pattern, hartree-fock++.cc:1574-1772:
    Engine per thread, precomputed SchwarzInf ShellPairs) on the GPU library == on the CPU kernels."""
    from libint_b200.basis import BasisSet, H2O_XYZ_ANGSTROM, atoms_from_tuples
    bs = BasisSet("6-31g", atoms_from_tuples(H2O_XYZ_ANGSTROM))
    ns = len(bs)
    s1, s2 = np.array([(a, b) for a in range(ns) for b in range(a + 1)], dtype=np.int32).T
    sh = oracle.Shells(*bs.flat(), raw=False)
    rng = np.random.default_rng(8)
    D = rng.standard_normal((bs.nbf, bs.nbf)) * 0.2
    D = 0.5 * (D + D.T)
    fg = oracle.Fock(sh, s1, s2, nthreads=2, b200=True)
    fo = oracle.Fock(sh, s1, s2, nthreads=2)
    np.testing.assert_allclose(fg.schwarz(), fo.schwarz(), rtol=1e-12, atol=1e-15)
    Gg, stg = fg.build(D, 1e-12)
    Go, sto = fo.build(D, 1e-12)
    assert stg["nquartets"] == sto["nquartets"]
    np.testing.assert_allclose(Gg, Go, rtol=1e-12, atol=5e-14)
