"""include/libint_b200.hpp: the header-only C++ mirror of libint2::Engine (engine.h:503-526,
:787-791) and of the direct-SCF Fock builder over the C ABI.  The driver tests/cxx/cxx_api_driver.cc
is compiled with g++ against the header and the in-tree .so; without a GPU it must fail loudly,
with one its numbers are compared with the oracle (reference Engine)."""
import os
import subprocess

import numpy as np
import pytest

from util import assert_parity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "libint_b200", "_lib")


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("cxx") / "cxx_api_driver")
    cmd = ["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cxx", "cxx_api_driver.cc"), "-o", out,
           "-L", LIBDIR, "-llibint_b200", "-Wl,-rpath," + LIBDIR]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return out


def _write_shells(path, table):
    l, pure, nprim, O, al, co = table
    off = np.concatenate([[0], np.cumsum(nprim)])
    with open(path, "w") as f:
        f.write("%d\n" % len(l))
        for s in range(len(l)):
            f.write("%d %d %d %.17g %.17g %.17g\n" % (l[s], pure[s], nprim[s], *np.asarray(O)[s]))
            f.write(" ".join("%.17g" % x for x in al[off[s]:off[s + 1]]) + "\n")
            f.write(" ".join("%.17g" % x for x in co[off[s]:off[s + 1]]) + "\n")


def _table(seed=5):
    rng = np.random.default_rng(seed)
    l = [2, 1, 0, 2, 1, 3]
    pure = [1, 0, 0, 0, 0, 1]
    nprim = [2, 3, 1, 1, 2, 1]
    O = rng.uniform(-1.2, 1.2, (len(l), 3))
    al = rng.uniform(0.3, 2.5, sum(nprim))
    co = rng.uniform(0.3, 1.4, sum(nprim))
    return l, pure, nprim, O, al, co


def test_cxx_header_compiles_and_fails_loudly_without_gpu(driver, tmp_path):
    import torch
    p = str(tmp_path / "shells.txt")
    _write_shells(p, _table())
    r = subprocess.run([driver, p, "lmax"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.split() == ["lmax_exceeded", "4", "5"]
    if torch.cuda.is_available():
        pytest.skip("GPU present: the no-GPU contract is checked on the CPU box")
    r = subprocess.run([driver, p, "eri", "0", "1", "2", "3"], capture_output=True, text=True)
    assert r.returncode == 3, (r.returncode, r.stderr)      # LB200_ERR_CUDA through the C++ API
    assert "libint_b200::error" in r.stderr


@pytest.mark.gpu
def test_cxx_engine_vs_oracle(driver, oracle, tmp_path):
    po = oracle
    tab = _table()
    p = str(tmp_path / "shells.txt")
    _write_shells(p, tab)
    quartets = [(0, 1, 2, 3), (1, 0, 3, 2), (2, 3, 0, 1), (4, 5, 1, 0), (5, 5, 5, 5), (2, 2, 2, 2), (1, 4, 0, 3)]
    args = [str(i) for q in quartets for i in q]
    r = subprocess.run([driver, p, "eri"] + args, capture_output=True, text=True, check=True)
    lines = r.stdout.strip().splitlines()
    assert len(lines) == len(quartets)
    sh = po.Shells(*tab, raw=False)
    for q, ln in zip(quartets, lines):
        ref = po.compute2(sh.subset(list(q)), precision=0.0).ravel()
        got = np.array([float(x) for x in ln.split()])
        assert_parity(got, ref, "C++ Engine::compute %s" % (q,))
    triplets = [(5, 0, 1), (0, 3, 4), (3, 1, 0)]
    r = subprocess.run([driver, p, "eri3"] + [str(i) for t in triplets for i in t],
                       capture_output=True, text=True, check=True)
    for t, ln in zip(triplets, r.stdout.strip().splitlines()):
        ref = po.compute2(sh.subset(list(t)), braket=1, precision=0.0).ravel()
        assert_parity(np.array([float(x) for x in ln.split()]), ref, "C++ Engine xs_xx %s" % (t,))


@pytest.mark.gpu
def test_cxx_fock_builder_vs_oracle(driver, oracle, tmp_path):
    from libint_b200.basis import BasisSet, H2O_XYZ_ANGSTROM, atoms_from_tuples
    po = oracle
    bs = BasisSet("6-31g", atoms_from_tuples(H2O_XYZ_ANGSTROM))
    p = str(tmp_path / "shells.txt")
    _write_shells(p, bs.flat())
    rng = np.random.default_rng(2)
    D = rng.standard_normal((bs.nbf, bs.nbf)) * 0.2
    D = 0.5 * (D + D.T)
    dp = str(tmp_path / "D.txt")
    np.savetxt(dp, D.ravel(), fmt="%.17g")
    r = subprocess.run([driver, p, "fock", dp, "1e-13"], capture_output=True, text=True, check=True)
    G = np.array([float(x) for x in r.stdout.split()]).reshape(bs.nbf, bs.nbf)
    ns = len(bs)
    s1, s2 = np.array([(a, b) for a in range(ns) for b in range(a + 1)], dtype=np.int32).T
    Gref, _ = po.Fock(po.Shells(*bs.flat(), raw=False), s1, s2, nthreads=2).build(D, 1e-13)
    assert_parity(G, Gref, "C++ FockBuilder", rtol=1e-12, atol=2e-14)


@pytest.mark.gpu
def test_cxx_forces_vs_golden(driver, tmp_path):
    """FockBuilder::compute_2body_forces (lb200_fock_grad) from C++ against the committed oracle forces"""
    from libint_b200.basis import BasisSet, H2O_XYZ_ANGSTROM, atoms_from_tuples
    d = np.load(os.path.join(ROOT, "tests", "golden", "grad_h2o.npz"))
    bs = BasisSet("6-31g*", atoms_from_tuples(H2O_XYZ_ANGSTROM))
    p = str(tmp_path / "shells.txt")
    _write_shells(p, bs.flat())
    dp = str(tmp_path / "D.txt")
    np.savetxt(dp, d["631gs_D"].ravel(), fmt="%.17g")
    r = subprocess.run([driver, p, "forces", dp, "1e-16", "3"] + [str(a) for a in bs.shell2atom],
                       capture_output=True, text=True, check=True)
    F2 = np.array([float(x) for x in r.stdout.split()]).reshape(3, 3)
    ref = d["631gs_F2"]
    assert np.abs(F2 - ref).max() <= 1e-11 * max(1.0, np.abs(ref).max())


def test_cxx_basisset_mirror_matches_the_python_mirror(tmp_path):
    """CPU: include/libint_b200_basis.hpp -- libint_b200::Atom / read_dotxyz / BasisSet(name, atoms), the C++ mirror of
    libint2::BasisSet (basis.h.in:89-617) -- gives the same shells, maps and counts as libint_b200.basis.BasisSet (which
    tests/test_host.py checks against the reference's own reader), incl. the Cartesian-d convention of 6-31G*
    (test-basis.cc:24-49), the aug-cc-pVDZ decomposition, set_pure, and the error contracts."""
    from libint_b200 import basis as b
    out = str(tmp_path / "basis_api_driver")
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cxx", "basis_api_driver.cc"), "-o", out,
                    "-L", LIBDIR, "-llibint_b200", "-Wl,-rpath," + LIBDIR], check=True, capture_output=True, text=True)
    xyz = tmp_path / "h2o.xyz"
    xyz.write_text("3\n\n" + "".join("%s %.5f %.5f %.5f\n" % ({8: "O", 1: "H"}[Z], *r) for Z, r in b.H2O_XYZ_ANGSTROM))
    names = ["sto-3g", "6-31g*", "cc-pvdz", "aug-cc-pVDZ", "def2-tzvp", "cc-pvtz"]
    env = dict(os.environ, LIBINT_B200_DATA_PATH=os.path.join(ROOT, "libint_b200", "data"))
    r = subprocess.run([out, str(xyz)] + names, capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    atoms = b.atoms_from_tuples(b.H2O_XYZ_ANGSTROM)
    head = lines[0].split()
    assert head[:6] == ["natoms", "3", "Z", "8", "1", "1"] and float(head[-1]) == atoms[1].x
    for k, name in enumerate(names):
        blk = lines[1 + 8 * k: 9 + 8 * k]
        bs = b.BasisSet(name, atoms)
        f = blk[0].split()
        assert f[0] == name and [int(f[2]), int(f[4]), int(f[6]), int(f[8])] == [len(bs), bs.nbf, bs.max_nprim, bs.max_l]
        assert [int(x) for x in blk[1].split()[1:]] == list(bs.shell2bf)
        assert [int(x) for x in blk[2].split()[1:]] == list(bs.shell2atom)
        assert [int(x) for x in blk[3].split()[1:]] == [list(bs.shell2atom).count(a) for a in range(3)]
        assert [int(x) for x in blk[4].split()[1:]] == [int(s.pure) for s in bs]
        assert float(blk[5].split()[1]) == pytest.approx(sum(float(np.sum(s.coeff)) for s in bs), rel=1e-14)
        bs.set_pure(False)
        assert int(blk[6].split()[-1]) == bs.nbf
        bs.set_pure(True)
        assert int(blk[7].split()[-1]) == bs.nbf
    assert lines[-3:] == ["unknown: ios_base::failure", "quiet omit: 0 shells", "missing: logic_error"]
    env.pop("LIBINT_B200_DATA_PATH")
    r = subprocess.run([out, str(xyz), "sto-3g"], capture_output=True, text=True, env=env)
    assert r.returncode == 4 and "LIBINT_B200_DATA_PATH" in r.stderr
