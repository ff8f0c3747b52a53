"""First derivatives, CPU side: the derivative oracle (closed-form eri() with a derivative index, the
reference's own check of its eri1 kernels, tests/eri/test.cc:381-445) is pinned against finite differences
of the pinned reference Engine, and the host half of the GPU derivative path -- which six shifted shell
sets a class is assembled from and how they are laid out (lb200_eri_deriv1_plan) -- is checked by running
the same assembly in numpy on shell sets from the reference Engine."""
import numpy as np
import pytest

from util import nc, random_shell_table


def cart_components(l):
    return [(x, y, l - x - y) for x in range(l, -1, -1) for y in range(l - x, -1, -1)]


def cidx(q):
    l = sum(q)
    return ((l - q[0] + 1) * (l - q[0])) // 2 + l - q[0] - q[1]


def normalized_shells(po, rng, ls, K, spread=1.0):
    """random Cartesian shells with the normalization embedded (what the C ABI is handed)"""
    l, pure, nprim, O, al, co = random_shell_table(rng, ls, K, spread=spread)
    off = np.concatenate([[0], np.cumsum(nprim)])
    co = np.concatenate([po.shell_renorm(l[i], al[off[i]:off[i + 1]], co[off[i]:off[i + 1]])[0]
                         for i in range(len(l))])
    return po.Shells(l, pure, nprim, O, al, co, raw=False)


def shifted(po, sh, c, sgn):
    """the quartet with shell c raised (coefficients * 2 alpha) or lowered"""
    off = sh.offsets()
    l = sh.l.copy()
    l[c] += sgn
    co = sh.coeff.copy()
    if sgn > 0:
        co[off[c]:off[c + 1]] *= 2.0 * sh.alpha[off[c]:off[c + 1]]
    return po.Shells(l, sh.pure, sh.nprim, sh.O, sh.alpha, co, raw=False)


def assemble_from_plan(po, capi, sh):
    """numpy twin of deriv.cu: six shifted sets in the layout the store kernels write (the pair with the
    higher angular momentum first inside a twin), combined through the strides of lb200_eri_deriv1_plan"""
    ls = [int(x) for x in sh.l]
    plan = capi.eri_deriv1_plan(*ls)
    n = [nc(l) for l in ls]
    comps = [cart_components(l) for l in ls]
    out = np.zeros([12] + n)
    bufs = []
    for k in range(6):
        c, sgn = k // 2, (1 if k % 2 == 0 else -1)
        if ls[c] + sgn < 0:
            assert plan[k, 0] == 0
            bufs.append(None)
            continue
        s2 = shifted(po, sh, c, sgn)
        order = [0, 1, 2, 3]
        if c < 2 and s2.l[0] < s2.l[1]:
            order = [1, 0, 2, 3]
        if c == 2 and s2.l[2] < s2.l[3]:
            order = [0, 1, 3, 2]
        v = po.compute2(s2.subset(order), precision=0.0)
        assert v.size == plan[k, 0]
        bufs.append(v.ravel())
    for ia, qa in enumerate(comps[0]):
        for ib, qb in enumerate(comps[1]):
            for ic, qc in enumerate(comps[2]):
                for id_, qd in enumerate(comps[3]):
                    i = [ia, ib, ic, id_]
                    q = [qa, qb, qc, qd]
                    for c in range(3):
                        sp, sm = plan[2 * c, 1:], plan[2 * c + 1, 1:]
                        rest_p = sum(i[x] * sp[x] for x in range(4) if x != c)
                        rest_m = sum(i[x] * sm[x] for x in range(4) if x != c)
                        for d in range(3):
                            up = list(q[c]); up[d] += 1
                            val = bufs[2 * c][rest_p + cidx(up) * sp[c]]
                            if q[c][d] > 0:
                                dn = list(q[c]); dn[d] -= 1
                                val -= q[c][d] * bufs[2 * c + 1][rest_m + cidx(dn) * sm[c]]
                            out[3 * c + d][ia, ib, ic, id_] = val
    out[9:12] = -(out[0:3] + out[3:6] + out[6:9])
    return out.reshape(12, -1)


def test_closed_form_derivatives_match_finite_differences(oracle):
    """pins the derivative oracle: d/dR of the pinned reference Engine's integrals by central differences"""
    po = oracle
    rng = np.random.default_rng(5)
    for ls in [(0, 0, 0, 0), (1, 0, 1, 1), (2, 1, 1, 0)]:
        l, pure, nprim, O, al, co = random_shell_table(rng, ls, 2)
        d = po.deriv1_closed(po.Shells(l, pure, nprim, O, al, co, raw=True))
        h = 1e-4
        for c in range(4):
            for x in range(3):
                Op, Om = O.copy(), O.copy()
                Op[c, x] += h
                Om[c, x] -= h
                fp = po.compute2(po.Shells(l, pure, nprim, Op, al, co, raw=True), precision=0.0).ravel()
                fm = po.compute2(po.Shells(l, pure, nprim, Om, al, co, raw=True), precision=0.0).ravel()
                fd = (fp - fm) / (2 * h)
                assert np.abs(fd - d[3 * c + x]).max() <= 1e-6 * max(1.0, np.abs(d).max()), (ls, c, x)
    # translational invariance of the closed form itself
    assert np.abs(d[0:3] + d[3:6] + d[6:9] + d[9:12]).max() < 1e-12 * np.abs(d).max()


@pytest.mark.parametrize("ls", [(0, 0, 0, 0), (1, 0, 0, 0), (1, 1, 0, 0), (1, 0, 1, 0), (1, 1, 1, 1), (2, 0, 1, 1),
                                (2, 1, 1, 0), (2, 2, 1, 0), (2, 1, 2, 2), (1, 1, 2, 0)])
def test_derivative_assembly_plan(oracle, ls):
    """strides and shifted classes of lb200_eri_deriv1_plan: the numpy assembly reproduces the closed form
    (reference thresholds for derivative integrals, tests/eri/test.cc:77-88,434-437)"""
    from libint_b200 import capi
    po = oracle
    rng = np.random.default_rng(sum(ls) * 7 + ls[0])
    sh = normalized_shells(po, rng, ls, 1 if sum(ls) >= 5 else 2)
    got = assemble_from_plan(po, capi, sh)
    ref = po.deriv1_closed(sh)
    err = np.abs(got - ref)
    bad = (err > 1e-9 * np.abs(ref)) & (err > 5e-14)
    assert not bad.any(), (ls, err.max())
    # in practice the two agree far better than the reference's own thresholds ask
    assert err.max() <= 1e-11 * max(1.0, np.abs(ref).max()), (ls, err.max())


def test_derivative_plan_lmax():
    """raised twins without a kernel: the LIBINT2_MAX_AM_eri1 analogue"""
    from libint_b200 import capi
    with pytest.raises(capi.Lb200Error):
        capi.eri_deriv1_plan(3, 1, 0, 0)      # needs (g p|
    assert capi.eri_deriv1_plan(3, 0, 2, 2).shape == (6, 5)   # (g s| exists


def test_gradient_oracle_matches_finite_difference_of_fock_energy(oracle):
    """pins lbo_fock_grad_closed: F2 = d/dR trace(G(D; R) D) at fixed D, G from the pinned oracle Fock build"""
    po = oracle
    rng = np.random.default_rng(11)
    # three "atoms", Cartesian s/p/d shells
    centres = np.array([[0.0, -0.14, 0.0], [1.6, 1.1, 0.1], [-1.5, 1.2, -0.2]])
    ls = [0, 1, 2, 0, 0, 1]
    s2a = [0, 0, 0, 1, 2, 2]
    l, pure, nprim, _, al, co = random_shell_table(rng, ls, [2, 1, 1, 2, 1, 1])
    O = centres[s2a]
    sh = po.Shells(l, pure, nprim, O, al, co, raw=True)
    n = sum(nc(x) for x in ls)
    A = rng.standard_normal((n, n))
    D = 0.1 * (A + A.T)
    g = po.fock_grad_closed(sh, D, s2a, 3, nthreads=4)
    assert np.abs(g.sum(axis=0)).max() < 1e-10 * np.abs(g).max()    # no net force
    ns = len(ls)
    p1 = [a for a in range(ns) for b in range(a + 1)]
    p2 = [b for a in range(ns) for b in range(a + 1)]

    def energy(cen):
        f = po.Fock(po.Shells(l, pure, nprim, cen[s2a], al, co, raw=True), p1, p2, nthreads=2)
        G, _ = f.build(D, 1e-16, use_schwarz=False)
        f.close()
        return float((G * D).sum())

    h = 1e-4
    for a, x in [(0, 1), (1, 0), (2, 2)]:
        cp, cm = centres.copy(), centres.copy()
        cp[a, x] += h
        cm[a, x] -= h
        fd = (energy(cp) - energy(cm)) / (2 * h)
        assert abs(fd - g[a, x]) <= 1e-6 * max(1.0, np.abs(g).max()), (a, x, fd, g[a, x])


# ---------------------------------------------------------------------------------------------------
# one-body forces: the __host__ __device__ core of lb200_onebody_forces compiled for the CPU
# ---------------------------------------------------------------------------------------------------
def onebody_force_inputs(bs, atoms, seed=7):
    """random symmetric D, W in the basis functions of `bs`, their Cartesian-ised twins C^T D C, and the
    reference values from the numpy derivative integrals (pinned to the reference's golden forces in test_scf.py)"""
    from libint_b200 import onebody
    rng = np.random.default_rng(seed)
    n = bs.nbf
    D = rng.standard_normal((n, n)) * 0.3
    D = 0.5 * (D + D.T)
    W = rng.standard_normal((n, n)) * 0.3
    W = 0.5 * (W + W.T)
    S1, T1, V1 = onebody.compute_1body_ints_deriv(bs, atoms)
    F1 = 2.0 * np.einsum("kij,ij->k", T1 + V1, D)
    FP = -2.0 * np.einsum("kij,ij->k", S1, W)
    nbfc = sum(nc(s.l) for s in bs)
    C = np.zeros((n, nbfc))            # pure (or Cartesian) function <- Cartesian components, block diagonal
    oc = 0
    for i, s in enumerate(bs):
        k = nc(s.l)
        blk = onebody.cart_to_pure(s.l) if s.pure else np.eye(k)
        C[bs.shell2bf[i]:bs.shell2bf[i] + blk.shape[0], oc:oc + k] = blk
        oc += k
    return D, W, C.T @ D @ C, C.T @ W @ C, F1, FP


def host_onebody_forces(tmp_path_factory_dir, bs, atoms, Dc, Wc):
    import ctypes
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = os.path.join(str(tmp_path_factory_dir), "libob1host.so")
    if not os.path.exists(so):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared",
                               "-I", os.path.join(root, "libint_b200", "csrc"),
                               os.path.join(root, "tests", "cxx", "onebody_forces_host.cc"), "-o", so])
    lib = ctypes.CDLL(so)
    l, pure, nprim, O, alpha, coeff = bs.flat()
    off = np.concatenate([[0], np.cumsum(nprim)]).astype(np.int32)
    s2c = np.concatenate([[0], np.cumsum([nc(int(x)) for x in l])]).astype(np.int32)
    s2a = np.ascontiguousarray(bs.shell2atom, dtype=np.int32)
    ch = np.ascontiguousarray([[a.atomic_number, *a.xyz] for a in atoms], dtype=np.float64)
    F = np.zeros((2, 3 * len(atoms)))
    ip, dp = ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double)
    arrs = [np.ascontiguousarray(x) for x in (l, nprim, off, O, alpha, coeff, s2c, s2a, ch, Dc, Wc)]
    ptr = [x.ctypes.data_as(ip if x.dtype == np.int32 else dp) for x in arrs]
    lib.ob1_host_forces(ctypes.c_int(len(l)), ptr[0], ptr[1], ptr[2], ptr[3], ptr[4], ptr[5], ptr[6], ptr[7],
                        ctypes.c_int(len(atoms)), ptr[8], ptr[9], ptr[10], ctypes.c_int(Dc.shape[0]),
                        F.ctypes.data_as(dp))
    return F[0], F[1]


@pytest.mark.parametrize("name,geom,set_pure", [("aug-cc-pvdz", "h2o_rotated", None), ("6-31g*", "h2o", None),
                                                ("def2-tzvp", "h2o_rotated", None), ("cc-pvdz", "h2o", False)])
def test_onebody_force_kernel_core_on_the_cpu(tmp_path, name, geom, set_pure):
    """lb200_onebody_forces' per-shell-pair routine (onebody_deriv.cuh, compiled by g++ here) against
    2 sum (T1 + V1) o D and -2 sum S1 o W from the numpy derivative integrals: pure d / f, Cartesian d,
    diffuse functions, three centres."""
    from libint_b200.basis import BasisSet, H2O_ROTATED_XYZ_ANGSTROM, H2O_XYZ_ANGSTROM, atoms_from_tuples
    atoms = atoms_from_tuples(H2O_ROTATED_XYZ_ANGSTROM if geom == "h2o_rotated" else H2O_XYZ_ANGSTROM)
    bs = BasisSet(name, atoms)
    if set_pure is not None:
        bs.set_pure(set_pure)
    D, W, Dc, Wc, F1, FP = onebody_force_inputs(bs, atoms)
    g1, gp = host_onebody_forces(tmp_path, bs, atoms, Dc, Wc)
    np.testing.assert_allclose(g1, F1, rtol=1e-11, atol=1e-11 * np.abs(F1).max())
    np.testing.assert_allclose(gp, FP, rtol=1e-11, atol=1e-11 * np.abs(FP).max())
