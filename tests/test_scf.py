"""SCF energies: the reference's golden Hartree-Fock energies through libint_b200.scf.

CPU half (oracle G through the same loop) pins the oracle against
tests/hartree-fock/hartree-fock-validate.py:14 (-74.942080057696, STO-3G, tol 1e-11),
tests/hartree-fock/hartree-fock++-validate.py:49 (-76.003354058439, h2o_rotated / aug-cc-pVDZ,
tol 5e-12) and python/tests/test_hf.py:117 (-75.1903033978, 6-31G, 7 places).  GPU half runs the
identical loop on the CUDA Fock build; tolerance 1e-10 Eh (BASELINE.json north_star)."""
import numpy as np
import pytest

ETOL = 1e-10


def _atoms(which):
    from libint_b200 import basis as b
    if which == "h2o_2010":   # hartree-fock.cc converts with the CODATA-2010 bohr
        f = 1.0 / b.BOHR_TO_ANGSTROM_CODATA2010
        return [b.Atom(Z, r[0] * f, r[1] * f, r[2] * f) for Z, r in b.H2O_XYZ_ANGSTROM]
    if which == "h2o_bohr":   # python/tests/test_hf.py passes the numbers as they are
        return b.atoms_from_tuples(b.H2O_XYZ_ANGSTROM, angstrom=False)
    if which == "h2o_rotated":
        return b.atoms_from_tuples(b.H2O_ROTATED_XYZ_ANGSTROM)
    return b.atoms_from_tuples(b.H2O_XYZ_ANGSTROM)


def _oracle_builder(po, bs):
    ns = len(bs)
    s1, s2 = np.array([(a, c) for a in range(ns) for c in range(a + 1)], dtype=np.int32).T
    f = po.Fock(po.Shells(*bs.flat(), raw=False), s1, s2, nthreads=4)
    return lambda D, prec: f.build(D, prec)[0]


def _gpu_builder(ctx, bs):
    from libint_b200.fock import FockBuilder
    fb = FockBuilder(bs, ctx=ctx, rank=0, nranks=1)
    return lambda D, prec: fb.build_partial(np.ascontiguousarray(D), prec)


GOLDEN = [("sto-3g", "h2o_2010", -74.942080057696), ("aug-cc-pvdz", "h2o_rotated", -76.003354058439)]


def _run(name, geom, builder_of):
    from libint_b200.basis import BasisSet
    from libint_b200.scf import RHF
    atoms = _atoms(geom)
    bs = BasisSet(name, atoms)
    scf = RHF(bs, atoms, builder_of(bs))
    e = scf.run()
    assert scf.converged
    return e, scf


@pytest.mark.parametrize("name,geom,eref", GOLDEN)
def test_oracle_scf_golden_energy(oracle, name, geom, eref):
    e, _ = _run(name, geom, lambda bs: _oracle_builder(oracle, bs))
    assert abs(e - eref) < ETOL, "%s: %.12f vs %.12f" % (name, e, eref)


def _python_test_hf(scf, ndocc):
    """The reference python test's own (loosely converged, DIIS-free) iteration,
    python/tests/test_hf.py:60-103, so that its 7-place golden is reproduced as printed."""
    import scipy.linalg
    H, S = scf.H, scf.S

    def dens(F):
        _, C = scipy.linalg.eigh(F, S)
        return C[:, :ndocc] @ C[:, :ndocc].T
    D, ehf = dens(H), 0.0
    for _ in range(30):
        F = np.asarray(scf.fock_builder(D, np.finfo(float).eps)) + H
        last = ehf
        D = dens(F)
        ehf = float(np.sum(D * (H + F)))
        if abs(last - ehf) < 1e-6:
            break
    return ehf + scf.enuc


def test_oracle_scf_python_golden(oracle):
    from libint_b200.basis import BasisSet
    from libint_b200.scf import RHF
    atoms = _atoms("h2o_bohr")
    bs = BasisSet("6-31g", atoms)
    scf = RHF(bs, atoms, _oracle_builder(oracle, bs))
    assert abs(_python_test_hf(scf, 5) - (-75.1903033978)) < 5e-8
    # fully converged energy: below the loosely converged golden by < 1e-6
    assert abs(scf.run() - (-75.1903033978)) < 1e-6


def test_onebody_invariants():
    """S has a unit diagonal (Shell::renorm, test-core.cc:52-55 convention), T is positive
    definite, V negative definite, and all three are rotation invariant in their spectra."""
    from libint_b200.basis import BasisSet
    from libint_b200 import onebody
    ev = []
    for geom in ("h2o", "h2o_rotated"):
        atoms = _atoms(geom)
        bs = BasisSet("cc-pvdz", atoms)
        S, T, V = onebody.compute_1body_ints(bs, atoms)
        assert np.allclose(np.diag(S), 1.0, atol=1e-13)
        assert np.linalg.eigvalsh(T).min() > 0 and np.linalg.eigvalsh(V).max() < 0
        w, U = np.linalg.eigh(S)
        X = U / np.sqrt(w)
        ev.append(np.linalg.eigvalsh(X.T @ (T + V) @ X))
    np.testing.assert_allclose(ev[0], ev[1], rtol=0, atol=2e-6)  # rotated file has 16 digits of a 5-digit geometry


def test_cartesian_standard_normalization_known_answers():
    """tests/unit/test-core.cc:80-135 ("cartesian uniform normalization", the standard-convention half): the
    self-overlap of a Cartesian d / f shell Shell{{1.0}, {{l, false, {1.0}}}, {0, 0, 0}} has the diagonal
    1, 1/3, 1/3, 1, 1/3, 1 (check_std_2) and 1, 1/5, 1/5, 1/5, 1/15, 1/5, 1, 1/5, 1/5, 1 (check_std_3): every
    component carries the normalization of x^l.  Pins the host one-body evaluation (the checker of the GPU
    one-body kernels) and the two-electron convention it shares to the reference's own known answers."""
    from libint_b200 import onebody
    from libint_b200.basis import BasisSet, Shell
    want = {2: [1, 1 / 3, 1 / 3, 1, 1 / 3, 1], 3: [1, 1 / 5, 1 / 5, 1 / 5, 1 / 15, 1 / 5, 1, 1 / 5, 1 / 5, 1]}
    for l, diag in want.items():
        bs = BasisSet(shells=[Shell(l, [(1.0, 1.0)], pure=False)])
        S, _, _ = onebody.compute_1body_ints(bs, [])
        np.testing.assert_allclose(np.diag(S), diag, rtol=1e-13, atol=0)


def test_nuclear_attraction_golden_vectors_of_the_reference():
    """tests/unit/test-1body.cc:45-110 ("electrostatic potential"): Engine(Operator::nuclear) shell sets of two
    contracted pure d shells in the field of the DefaultFixture's four point charges (fixture.h:38-41), printed there
    to 16 digits in standard solid-harmonic order -- (0|V|0) (the reference prescales by 2.3 and divides it out) and (0|V|1).  Pins the host one-body
    evaluation, i.e. the checker of lb200_onebody / lb200_onebody_forces (the GPU half compares against it)."""
    from libint_b200 import onebody
    from libint_b200.basis import Atom, BasisSet, Shell
    atoms = [Atom(8, 0., 0., 0.), Atom(8, 0., 0., 2.), Atom(1, 0., -1., -1.), Atom(1, 0., 1., 3.)]
    bs = BasisSet(shells=[Shell(2, [(1.0, 1.0), (3.0, 0.3)], origin=(0., 0., 0.), pure=True),
                          Shell(2, [(2.0, 1.0), (5.0, 0.2)], origin=(1., 1., 1.), pure=True)])
    _, _, V = onebody.compute_1body_ints(bs, atoms)
    ref00 = [-1.238239259091998e+01, 0, 0, -5.775996163160049e-02, 0, 0, -1.301230978657952e+01,
             -6.796143730068988e-02, 0, 1.139389632827834e-01, 0, -6.796143730068988e-02, -1.343732979153083e+01, 0,
             -1.478824785355970e-02, -5.775996163160049e-02, 0, 0, -1.284475452992947e+01, 0, 0,
             1.139389632827834e-01, -1.478824785355970e-02, 0, -1.241040347301479e+01]
    ref01 = [-4.769186621041819e-01, -9.303619356400431e-01, -1.559058302243514e+00, -9.290824121864600e-01,
             -5.835786921473129e-04, -1.159266418436018e+00, -3.770080831197964e-01, 9.572841308198474e-01,
             -8.291498398421207e-01, -1.663667687168316e+00, -2.171951144148577e+00, 1.074249956874296e+00,
             2.128355904665372e+00, 1.074590109905394e+00, -3.485163651594458e-03, -1.160865205880651e+00,
             -8.344173649626901e-01, 9.566621490332916e-01, -3.760919234260182e-01, 1.660514988916377e+00,
             -1.120272634615116e-03, -1.385603731947886e+00, -2.105750177166632e-03, 1.380654897976564e+00,
             2.115041199099945e+00]
    np.testing.assert_allclose(V[0:5, 0:5].ravel(), ref00, rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(V[0:5, 5:10].ravel(), ref01, rtol=1e-13, atol=1e-13)


def test_boys_host():
    from libint_b200.onebody import boys
    from scipy.special import hyp1f1
    for T in (0.0, 1e-3, 0.7, 5.0, 20.0, 34.9, 35.1, 80.0, 500.0):
        F = boys(8, T)
        ref = [hyp1f1(m + 0.5, m + 1.5, -T) / (2 * m + 1) for m in range(9)]
        np.testing.assert_allclose(F, ref, rtol=2e-13)


@pytest.mark.gpu
@pytest.mark.parametrize("name,geom,eref", GOLDEN)
def test_gpu_scf_golden_energy(ctx, name, geom, eref):
    e, scf = _run(name, geom, lambda bs: _gpu_builder(ctx, bs))
    assert abs(e - eref) < ETOL, "%s: %.12f vs %.12f" % (name, e, eref)


@pytest.mark.gpu
def test_gpu_scf_config1_h2o_ccpvdz(ctx, oracle):
    """BASELINE config 1 (h2o.xyz + cc-pVDZ, max_am 2): no golden exists for this pairing
    (SURVEY 8c); the GPU-driven SCF must agree with the oracle-driven one to 1e-10 Eh and
    iteration by iteration."""
    e_gpu, s_gpu = _run("cc-pvdz", "h2o", lambda bs: _gpu_builder(ctx, bs))
    e_cpu, s_cpu = _run("cc-pvdz", "h2o", lambda bs: _oracle_builder(oracle, bs))
    assert abs(e_gpu - e_cpu) < ETOL
    assert len(s_gpu.history) == len(s_cpu.history)
    for a, c in zip(s_gpu.history, s_cpu.history):
        assert abs(a[1] - c[1]) < 1e-9
    assert -76.0 < e_gpu < -75.98   # RHF/cc-pVDZ at h2o.xyz's stretched geometry (r_OH = 1.10 A)


@pytest.mark.gpu
def test_hartree_fock_cli_output_is_parsed_by_the_reference_validators(tmp_path):
    """`python -m libint_b200.hartree_fock` prints what hartree-fock-validate.py:19-24 and
    hartree-fock++-validate.py:60-70 look for; the regex and tolerances below are theirs."""
    import os
    import re
    import subprocess
    import sys
    from libint_b200 import basis as b
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    xyz = tmp_path / "h2o_rotated.xyz"
    xyz.write_text("3\nrotated water\n" + "".join(
        "%s %.17g %.17g %.17g\n" % ({8: "O", 1: "H"}[Z], *r) for Z, r in b.H2O_ROTATED_XYZ_ANGSTROM))
    cases = [(["--codata2010", str(tmp_path / "h2o.xyz"), "sto-3g"], -74.942080057696, 1e-11),
             ([str(xyz), "aug-cc-pVDZ"], -76.003354058439, 5e-12)]
    (tmp_path / "h2o.xyz").write_text("3\n\n" + "".join(
        "%s %.5f %.5f %.5f\n" % ({8: "O", 1: "H"}[Z], *r) for Z, r in b.H2O_XYZ_ANGSTROM))
    for argv, eref, tol in cases:
        r = subprocess.run([sys.executable, "-m", "libint_b200.hartree_fock"] + argv, capture_output=True,
                           text=True, cwd=root, timeout=600)
        assert r.returncode == 0, r.stderr[-1500:]
        found = [re.match(r"\*\* Hartree-Fock energy =\s*([-\d.]+)", ln) for ln in r.stdout.splitlines()]
        found = [m for m in found if m]
        assert len(found) == 1
        assert abs(eref - float(found[0].group(1))) < tol
        if "aug-cc-pVDZ" in argv:   # the force lines, hartree-fock++-validate.py:30-36 pattern, :128-132,:147-155
            num = r"\s*([+-]?(?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?)"
            for key, (ref, ftol) in REF_FORCES.items():
                m = [re.match(r"\*\* %s forces =" % re.escape(key) + num * 9, ln) for ln in r.stdout.splitlines()]
                m = [x for x in m if x]
                assert len(m) == 1, key
                assert max(abs(float(v) - x) for v, x in zip(m[0].groups(), ref)) < ftol, key


# forces of the reference's own validation run (h2o_rotated.xyz, aug-cc-pVDZ):
# tests/hartree-fock/hartree-fock++-validate.py:84-112, tolerances :89,:95,:101,:107,:112
REF_FORCES = {
    "1-body": ([-5.43569555903312, -1.88298017654395, -2.17427822361352, 3.47022732536532, -2.96798871167808,
                2.59189820350226, 1.9654682336678, 4.85096888822203, -0.417619979888738], 1e-9),
    "Pulay": ([0.355265310323155, 0.123067513529209, 0.142106124129298, -0.224258642015539, 0.180741977786854,
               -0.164305035772324, -0.131006668307617, -0.303809491316068, 0.0221989116430271], 1e-9),
    "2-body": ([2.95100851670393, 1.02225933689958, 1.18040340668181, -1.89116113654409, 1.64868682617684,
                -1.42151545972338, -1.05984738015984, -2.67094616307642, 0.241112053041571], 1e-9),
    "nuclear repulsion": ([2.01500148332517, 0.698016989289171, 0.806000593330069, -1.28187870168764,
                           1.07670120707691, -0.951756216715147, -0.733122781637531, -1.77471819636609,
                           0.145755623385078], 1e-10),
    "Hartree-Fock": ([-0.114420248680859, -0.0396363368259882, -0.0457680994723476, 0.0729288451180514,
                      -0.0618587006374771, 0.0543214912914121, 0.0414914035628126, 0.101495037463456,
                      -0.00855339181906306], 1e-9),
}


def test_oracle_forces_match_the_reference_golden_values(oracle):
    """CPU: pins the derivative oracle (closed-form derivative integrals digested as
    compute_2body_fock_deriv<1> does, oracle lbo_fock_grad_closed) and the host one-body derivative integrals
    against the reference's golden forces, with D from the oracle-driven SCF."""
    import os
    from libint_b200 import onebody
    from libint_b200.basis import BasisSet
    from libint_b200.scf import RHF
    from util import cartesianized
    atoms = _atoms("h2o_rotated")
    bs = BasisSet("aug-cc-pvdz", atoms)
    scf = RHF(bs, atoms, _oracle_builder(oracle, bs))
    scf.run()
    assert scf.converged
    S1, T1, V1 = onebody.compute_1body_ints_deriv(bs, atoms)
    Co = scf.C[:, :scf.ndocc]
    W = (Co * scf.evals[:scf.ndocc]) @ Co.T
    sh, Dc = cartesianized(oracle, bs, scf.D)
    f = {"1-body": 2.0 * np.einsum("kij,ij->k", T1 + V1, scf.D),
         "Pulay": -2.0 * np.einsum("kij,ij->k", S1, W),
         "2-body": oracle.fock_grad_closed(sh, Dc, bs.shell2atom, len(atoms), nthreads=os.cpu_count() or 4).ravel(),
         "nuclear repulsion": onebody.nuclear_repulsion_forces(atoms).ravel()}
    f["Hartree-Fock"] = sum(f.values())
    for key, (ref, tol) in REF_FORCES.items():
        err = np.abs(f[key] - np.array(ref)).max()
        assert err < tol, "%s forces: max deviation %.3g (tolerance %g)" % (key, err, tol)


@pytest.mark.gpu
def test_gpu_forces_match_the_reference_golden_values(ctx):
    """The reference's golden forces of its hartree-fock++ validation run: the 2-body part is
    compute_2body_fock_deriv<1> traced with D (hartree-fock++.cc:642-656) -- here lb200_fock_grad on the GPU --
    and the total closes with the host one-body derivative integrals; tolerances are the validator's."""
    from libint_b200.basis import BasisSet
    from libint_b200.fock import FockBuilder
    from libint_b200.scf import RHF, hf_forces
    atoms = _atoms("h2o_rotated")
    bs = BasisSet("aug-cc-pvdz", atoms)
    fb = FockBuilder(bs, ctx=ctx, rank=0, nranks=1)
    scf = RHF(bs, atoms, lambda D, prec: fb.build_partial(np.ascontiguousarray(D), prec))
    e = scf.run()
    assert scf.converged and abs(e - (-76.003354058439)) < ETOL
    f = hf_forces(scf, fb)
    for key, (ref, tol) in REF_FORCES.items():
        err = np.abs(f[key].ravel() - np.array(ref)).max()
        assert err < tol, "%s forces: max deviation %.3g (tolerance %g)" % (key, err, tol)


def _cxx_driver_inputs(tmp_path):
    import os
    from libint_b200 import basis as b
    from libint_b200 import build
    exe = build.build_driver()
    (tmp_path / "h2o.xyz").write_text("3\n\n" + "".join(
        "%s %.5f %.5f %.5f\n" % ({8: "O", 1: "H"}[Z], *r) for Z, r in b.H2O_XYZ_ANGSTROM))
    (tmp_path / "h2o_rotated.xyz").write_text("3\nrotated water\n" + "".join(
        "%s %.17g %.17g %.17g\n" % ({8: "O", 1: "H"}[Z], *r) for Z, r in b.H2O_ROTATED_XYZ_ANGSTROM))
    data = os.path.join(os.path.dirname(os.path.abspath(b.__file__)), "data", "basis")
    return exe, data


@pytest.mark.gpu
def test_cxx_driver_reproduces_the_reference_goldens(tmp_path):
    """hartree-fock-b200 (C++ host program on the C ABI, no Python in the loop): the reference's golden SCF
    energies and, for its validation run, all five golden force vectors -- one-body and Pulay (lb200_onebody_forces),
    two-body (lb200_fock_grad), nuclear repulsion and their sum (hartree-fock-validate.py:14,
    hartree-fock++-validate.py:49,84-112) -- parsed with the validators' patterns and checked to their tolerances."""
    import os
    import re
    import subprocess
    exe, data = _cxx_driver_inputs(tmp_path)
    num = r"\s*([+-]?(?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?)"
    cases = [([str(tmp_path / "h2o.xyz"), os.path.join(data, "sto-3g.json"), "--codata2010"], -74.942080057696, 1e-11, False),
             ([str(tmp_path / "h2o_rotated.xyz"), os.path.join(data, "cc-pvdz.json"),
               os.path.join(data, "augmentation-cc-pvdz.json")], -76.003354058439, 5e-12, True)]
    for argv, eref, tol, forces in cases:
        r = subprocess.run([exe] + argv, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        m = [re.match(r"\*\* Hartree-Fock energy =" + num, ln) for ln in r.stdout.splitlines()]
        m = [x for x in m if x]
        assert len(m) == 1 and abs(float(m[0].group(1)) - eref) < tol, r.stdout[-600:]
        if forces:
            for key, (ref, ftol) in REF_FORCES.items():   # 1-body, Pulay, 2-body, nuclear repulsion, Hartree-Fock
                mm = [re.match(r"\*\* %s forces =" % re.escape(key) + num * 9, ln) for ln in r.stdout.splitlines()]
                mm = [x for x in mm if x]
                assert len(mm) == 1, key
                assert max(abs(float(v) - x) for v, x in zip(mm[0].groups(), ref)) < ftol, key


def test_cxx_driver_basis_reader_matches_the_python_mirror(tmp_path):
    """CPU: the C++ driver's host side -- xyz reader, packed-basis reader, the pure / Cartesian-d rule of
    BasisSet (basis.h.in:368-386), component order of aug-cc-pVDZ (:388-400), Shell::renorm through the C ABI --
    builds the same shells as libint_b200.basis.BasisSet (itself checked against the reference's reader)."""
    import os
    import subprocess
    from libint_b200.basis import BasisSet
    exe, data = _cxx_driver_inputs(tmp_path)
    for geom, files, name in [("h2o", ["sto-3g.json"], "sto-3g"), ("h2o", ["6-31gs.json"], "6-31g*"),
                              ("h2o_rotated", ["cc-pvdz.json", "augmentation-cc-pvdz.json"], "aug-cc-pvdz"),
                              ("h2o", ["def2-tzvp.json"], "def2-tzvp"),
                              ("h2o_rotated", None, "aug-cc-pVDZ"), ("h2o", None, "6-31G*")]:   # by name, like hartree-fock++
        r = subprocess.run([exe, str(tmp_path / (geom + ".xyz"))] +
                           ([os.path.join(data, f) for f in files] if files else [name]) +
                           ["--dump-basis"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        lines = r.stdout.strip().splitlines()
        bs = BasisSet(name, _atoms(geom))
        assert int(lines[0]) == len(bs)
        for i, sh in enumerate(bs):
            h = lines[1 + 3 * i].split()
            assert [int(h[0]), int(h[1]), int(h[2]), int(h[6])] == [sh.l, int(sh.pure), sh.nprim, bs.shell2atom[i]]
            np.testing.assert_allclose([float(x) for x in h[3:6]], sh.O, rtol=0, atol=2e-10)   # xyz file: 5 digits
            np.testing.assert_array_equal([float(x) for x in lines[2 + 3 * i].split()], sh.alpha)
            np.testing.assert_array_equal([float(x) for x in lines[3 + 3 * i].split()], sh.coeff)


def test_cxx_driver_fails_loudly_without_gpu(tmp_path):
    import os
    import subprocess
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    exe, data = _cxx_driver_inputs(tmp_path)
    r = subprocess.run([exe, str(tmp_path / "h2o.xyz"), os.path.join(data, "sto-3g.json")], capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr
    assert "Hartree-Fock energy" not in r.stdout


def test_hartree_fock_cli_fails_loudly_without_gpu():
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "libint_b200.hartree_fock", "--codata2010"], capture_output=True,
                       text=True, cwd=root, timeout=300)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr
    assert "Hartree-Fock energy" not in r.stdout


@pytest.mark.gpu
def test_onebody_device_matches_host(ctx):
    """lb200_onebody (S, T, V on the GPU; Engine::compute1 of hartree-fock++.cc:267-275) against the host
    McMurchie-Davidson evaluation that reproduces the golden energies, pure d/f and Cartesian shells."""
    from libint_b200 import capi, onebody
    from libint_b200.basis import BasisSet
    atoms = _atoms("h2o_rotated")
    for name, pure in (("cc-pvdz", None), ("def2-tzvp", None), ("6-31g*", None), ("cc-pvdz", True)):
        bs = BasisSet(name, atoms)
        if pure is not None:
            bs.set_pure(pure)
        B = capi.Basis(ctx, *bs.flat())
        charges = [(float(a.atomic_number), a.xyz) for a in atoms]
        S, T, V = capi.onebody(ctx, B, charges)
        Sh, Th, Vh = onebody.compute_1body_ints(bs, atoms)
        np.testing.assert_allclose(S, Sh, rtol=1e-12, atol=1e-13, err_msg=name)
        np.testing.assert_allclose(T, Th, rtol=1e-12, atol=1e-12, err_msg=name)
        np.testing.assert_allclose(V, Vh, rtol=1e-12, atol=1e-12, err_msg=name)


@pytest.mark.gpu
def test_onebody_forces_device_matches_host(ctx):
    """lb200_onebody_forces (one-body and Pulay force sums on the GPU, hartree-fock++.cc:601-627) against the numpy
    derivative integrals contracted with the same random symmetric D and W: pure d / f, Cartesian d, diffuse
    functions; host and device (torch) inputs; argument errors."""
    import torch
    from libint_b200 import capi
    from libint_b200.basis import BasisSet
    from test_deriv_host import onebody_force_inputs
    for name, geom, pure in (("aug-cc-pvdz", "h2o_rotated", None), ("6-31g*", "h2o", None),
                             ("def2-tzvp", "h2o_rotated", None), ("cc-pvdz", "h2o", False)):
        atoms = _atoms(geom)
        bs = BasisSet(name, atoms)
        if pure is not None:
            bs.set_pure(pure)
        D, W, _, _, F1, FP = onebody_force_inputs(bs, atoms)
        B = capi.Basis(ctx, *bs.flat())
        charges = [(float(a.atomic_number), a.xyz) for a in atoms]
        g1, gp = capi.onebody_forces(ctx, B, charges, bs.shell2atom, D, W)
        np.testing.assert_allclose(g1.ravel(), F1, rtol=1e-11, atol=1e-11 * np.abs(F1).max(), err_msg=name)
        np.testing.assert_allclose(gp.ravel(), FP, rtol=1e-11, atol=1e-11 * np.abs(FP).max(), err_msg=name)
        dev = torch.device("cuda", ctx.device)
        Dt, Wt = torch.as_tensor(D, device=dev), torch.as_tensor(W, device=dev)
        torch.cuda.synchronize(dev)
        h1, hp = capi.onebody_forces(ctx, B, charges, bs.shell2atom, Dt, Wt)
        np.testing.assert_allclose(h1, g1, rtol=1e-12, atol=1e-12 * np.abs(F1).max(), err_msg=name)
        np.testing.assert_allclose(hp, gp, rtol=1e-12, atol=1e-12 * np.abs(FP).max(), err_msg=name)
        # translational invariance of the one-body energy terms: the forces sum to zero over the atoms
        assert np.abs(g1.sum(axis=0)).max() < 1e-10 * np.abs(F1).max()
        assert np.abs(gp.sum(axis=0)).max() < 1e-10 * np.abs(FP).max()
        with pytest.raises(capi.Lb200Error):
            capi.onebody_forces(ctx, B, charges, [len(atoms)] * B.nshell, D, W)   # shell2atom out of range
        with pytest.raises(ValueError):
            capi.onebody_forces(ctx, B, charges, bs.shell2atom[:-1], D, W)


@pytest.mark.gpu
@pytest.mark.parametrize("name,geom,eref", GOLDEN)
def test_device_scf_golden_energy(ctx, name, geom, eref):
    """the whole RHF loop on the device (RHFDevice: GPU one-body integrals, cuSOLVER eigensolves, GPU
    Fock builds with device buffers) reproduces the reference's golden energies to 1e-10 Eh."""
    from libint_b200.basis import BasisSet
    from libint_b200.fock import FockBuilder
    from libint_b200.scf import RHFDevice
    atoms = _atoms(geom)
    bs = BasisSet(name, atoms)
    fb = FockBuilder(bs, ctx=ctx, rank=0, nranks=1)
    scf = RHFDevice(bs, atoms, fb)
    e = scf.run()
    assert scf.converged
    assert abs(e - eref) < ETOL, "%s: %.12f vs %.12f" % (name, e, eref)
    assert scf.incremental_builds > 0 and scf.full_builds > 1   # hartree-fock++.cc:420-480 schedule was exercised
    scf2 = RHFDevice(bs, atoms, fb)
    e2 = scf2.run(incremental=False)
    assert scf2.incremental_builds == 0 and abs(e2 - eref) < ETOL
