"""Host-side logic and the C-ABI surface; no GPU needed."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    """liblibint_b200.so loads without a GPU and exports what include/libint_b200.h declares."""
    from libint_b200 import capi
    hdr = open(os.path.join(ROOT, "include", "libint_b200.h")).read()
    declared = set(re.findall(r"\b(lb200_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    L = ctypes.CDLL(capi.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(L, name), "missing export " + name
    assert declared == set(capi.SIGNATURES), declared ^ set(capi.SIGNATURES)


def test_no_gpu_fails_loudly():
    """no CPU fallback: without a device context creation raises (on the GPU box it succeeds)."""
    from libint_b200 import capi
    if capi.load().lb200_device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(capi.Lb200Error):
        capi.Context(0)


def test_shell_renorm_matches_reference(oracle):
    """lb200_shell_renorm == Shell::renorm (shell.h:958-999) bit for bit."""
    from libint_b200 import capi
    rng = np.random.default_rng(0)
    for l in range(5):
        for K in (1, 3, 6):
            al = rng.uniform(0.05, 50.0, K)
            co = rng.uniform(-1.0, 1.0, K)
            c, m = capi.shell_renorm(l, al, co, True)
            rc, rm = oracle.shell_renorm(l, al, co)
            np.testing.assert_array_equal(c, rc)
            np.testing.assert_array_equal(m, rm)
    c, _ = capi.shell_renorm(2, [1.0], [1.0], True)
    assert c[0] == pytest.approx(1.64592278064949, abs=1e-14)  # tests/unit/test-core.cc:52-55


def test_basisset_sizes():
    """shell / function counts of the BASELINE.json configurations."""
    from libint_b200.basis import BasisSet, H2O_XYZ_ANGSTROM, alkane, atoms_from_tuples, water_cluster
    h2o = atoms_from_tuples(H2O_XYZ_ANGSTROM)
    bs = BasisSet("cc-pVDZ", h2o)
    assert (len(bs), bs.nbf, bs.max_nprim, bs.max_l) == (12, 24, 8, 2)
    assert len(BasisSet("6-31g", h2o)) == 9  # python/tests/test_libint2.py:25
    w64 = water_cluster(4, 4, 4)
    assert len(w64) == 192
    bs = BasisSet("def2-tzvp", w64)
    assert (len(bs), bs.nbf, bs.max_l) == (1216, 2752, 3)
    bs = BasisSet("cc-pvtz", water_cluster(2, 2, 1))
    assert (len(bs), bs.nbf) == (22 * 4, 58 * 4)
    c40 = alkane(40)
    assert len(c40) == 122 and sum(a.atomic_number == 6 for a in c40) == 40
    assert (len(BasisSet("def2-tzvp", c40)), BasisSet("def2-tzvp", c40).nbf) == (768, 1732)
    df = BasisSet("def2-tzvp-jk", c40)
    assert (len(df), df.nbf, df.max_l) == (1492, 4476, 4)
    # minimum interatomic distance is chemically sane
    xyz = np.array([a.xyz for a in c40])
    d = np.linalg.norm(xyz[:, None] - xyz[None], axis=-1) + np.eye(len(xyz)) * 9
    assert d.min() > 1.9  # bohr; C-H = 2.06


def test_basisset_matches_reference_reader(oracle):
    """our G94 semantics == BasisSet(name, atoms) of the reference (needs /root/reference)."""
    if not os.path.isdir("/root/reference/lib/basis"):
        pytest.skip("reference tree not present")
    from libint_b200.basis import BasisSet, H2O_XYZ_ANGSTROM, atoms_from_tuples
    atoms = atoms_from_tuples(H2O_XYZ_ANGSTROM)
    for name in ["sto-3g", "6-31g", "6-31g*", "cc-pvdz", "aug-cc-pvdz", "cc-pvtz", "def2-tzvp",
                 "def2-tzvp-jk"]:
        bs = BasisSet(name, atoms)
        ref = oracle.basis_load(name, [a.atomic_number for a in atoms], [a.xyz for a in atoms],
                                "/root/reference/lib")
        l, pure, nprim, O, al, co = bs.flat()
        for x, y in ((l, ref.l), (pure, ref.pure), (nprim, ref.nprim), (O, ref.O), (al, ref.alpha),
                     (co, ref.coeff)):
            np.testing.assert_array_equal(x, y)


def test_basis_cartesian_vs_solid_defaults_known_answer():
    """tests/unit/test-basis.cc:24-49: O2 in 6-31G* has Cartesian d shells, 2 * (1*3 + 3*2 + 6*1) = 30 functions
    (basis.h.in:368-386); the correlation-consistent and def2 sets use solid harmonics for l >= 2."""
    from libint_b200.basis import Atom, BasisSet, ANGSTROM_TO_BOHR
    o2 = [Atom(8, 0., 0., 0.), Atom(8, 0., 0., 1.5 * ANGSTROM_TO_BOHR)]
    assert BasisSet("6-31g*", o2).nbf == 30
    assert BasisSet("6-31g", o2).nbf == 18
    assert BasisSet("cc-pvdz", o2).nbf == 2 * (3 + 3 * 2 + 5)
    assert all(s.pure == (s.l > 1) for s in BasisSet("def2-tzvp", o2))


def test_read_dotxyz(tmp_path):
    from libint_b200.basis import ANGSTROM_TO_BOHR, read_dotxyz
    p = tmp_path / "h2o.xyz"
    p.write_text("3\n\nO 0.0 -0.07579 0.0\nH 0.86681 0.60144 0.0\nH -0.86681 0.60144 0.0\n")
    atoms = read_dotxyz(str(p))
    assert [a.atomic_number for a in atoms] == [8, 1, 1]
    assert atoms[1].x == 0.86681 * ANGSTROM_TO_BOHR
    p.write_text("1\n\nXx 0 0 0\n")
    with pytest.raises(ValueError):
        read_dotxyz(str(p))


def test_task_owner_partitions_quartets():
    """every (bra pair, ket pair) quartet has exactly one owner, ownership goes by the bra row, and
    the shares of a triangular task matrix are even."""
    from libint_b200 import capi
    npair = 4000
    for nranks in (1, 2, 4, 8):
        owner = np.array([capi.task_owner(gi, 0, nranks) for gi in range(npair)])
        assert owner.min() >= 0 and owner.max() < nranks
        assert all(capi.task_owner(gi, gj, nranks) == owner[gi] for gi in (5, 77, 3999) for gj in (0, 3, gi))
        counts = np.bincount(owner, weights=np.arange(1, npair + 1), minlength=nranks)   # row gi: gi+1 kets
        assert counts.sum() == npair * (npair + 1) // 2
        assert counts.max() <= 1.1 * counts.mean()
    with pytest.raises(capi.Lb200Error):
        capi.task_owner(1, 1, 0)


def test_engine_argument_errors():
    """Engine ctor contract (tests/unit/test-core.cc:59-78): unsupported requests raise."""
    from libint_b200 import engine as E
    with pytest.raises(E.lmax_exceeded):
        E.Engine(E.Operator.coulomb, 1, 7, ctx=object())
    with pytest.raises(NotImplementedError):
        E.Engine(E.Operator.coulomb, 1, 1, deriv_order=2, ctx=object())
    with pytest.raises(NotImplementedError):   # first derivatives: four-centre integrals only
        E.Engine(E.Operator.coulomb, 1, 1, deriv_order=1, braket=E.BraKet.xs_xx, ctx=object())
    assert len(E.Engine(E.Operator.coulomb, 1, 1, deriv_order=1, ctx=object()).results()) == 12


def test_iface_library_exports_the_reference_boundary():
    """liblibint_b200_iface.so loads without a GPU and exports every symbol of the reference's generated
    C interface (src/bin/libint/iface.cc:114-185,302-418) that include/libint2/util/generated/libint2_iface.h
    declares; the function tables are dimensioned by the MAX_AM macros of libint2_params.h and canonical
    entries are filled by libint2_static_init (null = no kernel, engine.impl.h:1898)."""
    from libint_b200 import capi
    inc = os.path.join(ROOT, "include", "libint2", "util", "generated")
    hdr = open(os.path.join(inc, "libint2_iface.h")).read()
    names = set(re.findall(r"\b(libint2_[a-z0-9_]+)\s*[\[(]", hdr))
    assert {"libint2_build_eri", "libint2_build_3eri", "libint2_build_2eri", "libint2_build_default",
            "libint2_static_init", "libint2_static_cleanup", "libint2_init_eri", "libint2_need_memory_eri",
            "libint2_cleanup_eri", "libint2_init_default", "libint2_cleanup_default"} <= names
    path = os.path.join(os.path.dirname(capi.LIB_PATH), "liblibint_b200_iface.so")
    L = ctypes.CDLL(path)
    for n in sorted(names):
        assert hasattr(L, n), "missing export " + n
    params = open(os.path.join(inc, "libint2_params.h")).read()
    am = {k: int(v) for k, v in re.findall(r"#define LIBINT2_MAX_AM_(\w+) (\d+)", params)}
    assert am == {"default": 4, "eri": 3, "3eri": 4, "2eri": 4, "eri1": 2}
    L.libint2_need_memory_eri.restype = ctypes.c_size_t
    L.libint2_need_memory_eri.argtypes = [ctypes.c_int]
    assert L.libint2_need_memory_eri(2) >= 6 ** 4
    # static_init fills the canonical classes without touching the GPU
    L.libint2_static_init()
    n4 = am["eri"] + 1
    tab = (ctypes.c_void_p * (n4 ** 4)).in_dll(L, "libint2_build_eri")
    idx = lambda a, b, c, d: ((a * n4 + b) * n4 + c) * n4 + d
    assert tab[idx(0, 0, 0, 0)] is None            # (ss|ss) is done by the Engine itself (engine.impl.h:1890)
    assert tab[idx(1, 0, 1, 0)] and tab[idx(2, 2, 2, 2)] and tab[idx(3, 3, 3, 3)] and tab[idx(1, 1, 3, 0)]
    assert tab[idx(1, 1, 1, 0)] is None            # not canonical: la + lb > lc + ld (build_libint.cc:78-83)
    assert tab[idx(0, 1, 1, 1)] is None            # not canonical: la < lb
    # first derivatives: libint2_build_eri1, every canonical class of l <= 2 (its raised twins go up to (f d|)
    assert {"libint2_build_eri1", "libint2_init_eri1", "libint2_need_memory_eri1", "libint2_cleanup_eri1"} <= names
    assert re.search(r"#define LIBINT2_MAX_DERIV_ORDER 1\b", params)
    n1 = am["eri1"] + 1
    tab1 = (ctypes.c_void_p * (n1 ** 4)).in_dll(L, "libint2_build_eri1")
    idx1 = lambda a, b, c, d: ((a * n1 + b) * n1 + c) * n1 + d
    assert tab1[idx1(0, 0, 0, 0)] and tab1[idx1(1, 0, 2, 1)] and tab1[idx1(2, 2, 2, 2)]
    assert tab1[idx1(1, 1, 1, 0)] is None and tab1[idx1(0, 1, 1, 1)] is None
    L.libint2_need_memory_eri1.restype = ctypes.c_size_t
    L.libint2_need_memory_eri1.argtypes = [ctypes.c_int]
    assert L.libint2_need_memory_eri1(2) >= 12 * 6 ** 4


def test_generated_iface_headers_are_current(tmp_path):
    """include/libint2/util/generated/*.h are what tools/gen_iface_headers.py writes."""
    import importlib
    gen = importlib.import_module("libint_b200.tools.gen_iface_headers")
    inc = os.path.join(ROOT, "include", "libint2", "util", "generated")
    for name, text in (("libint2_params.h", gen.params()), ("libint2_types.h", gen.types()),
                       ("libint2_iface.h", gen.iface())):
        assert open(os.path.join(inc, name)).read() == text, name + " is stale: run tools/gen_iface_headers.py"
