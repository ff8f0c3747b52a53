"""First geometric derivatives on the GPU: the batched
Engine::compute2<coulomb, xx_xx, 1> (lb200_eri_deriv1_batch) and the two-body forces of the direct-SCF driver
(lb200_fock_grad; compute_2body_fock_deriv<1>, tests/hartree-fock/hartree-fock++.cc:1775-2055,:642-656),
through the C ABI, against the reference's closed-form derivative integrals (the check the reference itself
applies to its eri1 kernels, tests/eri/test.cc:381-445, same thresholds) and against finite differences."""
import os

import numpy as np
import pytest

from util import _sph_matrix, nc, random_shell_table

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# tests/eri/test.cc:77-88,434-437: an element fails only if it is off by more than 1e-9 relative AND 5e-14
# absolute
REL_DEV, ABS_DEV = 1.0e-9, 5.0e-14


def assert_reference_thresholds(got, ref, what):
    err = np.abs(np.asarray(got) - np.asarray(ref))
    bad = (err > REL_DEV * np.abs(ref)) & (err > ABS_DEV)
    assert not bad.any(), "%s: %d elements outside the reference's thresholds, max abs err %.3g" % (
        what, int(bad.sum()), err.max())
    # and far inside them in practice: both sides are the same recurrences in double precision
    assert err.max() <= 1e-11 * max(1.0, np.abs(ref).max()), (what, err.max())


def _normalized(po, rng, ls, K, pure=None):
    l, pu, nprim, O, al, co = random_shell_table(rng, ls, K, pure=pure)
    off = np.concatenate([[0], np.cumsum(nprim)])
    co = np.concatenate([po.shell_renorm(l[i], al[off[i]:off[i + 1]], co[off[i]:off[i + 1]])[0]
                         for i in range(len(l))])
    return po.Shells(l, pu, nprim, O, al, co, raw=False)


CLASSES = [(0, 0, 0, 0), (1, 0, 0, 0), (0, 0, 1, 0), (1, 1, 0, 0), (1, 0, 1, 0), (1, 1, 1, 1), (2, 0, 0, 0),
           (2, 0, 1, 1), (1, 1, 2, 0), (2, 1, 1, 0), (2, 2, 1, 0), (2, 1, 2, 1), (2, 2, 2, 2), (0, 0, 2, 2),
           (3, 0, 1, 0)]


@pytest.mark.parametrize("ls", CLASSES)
def test_deriv1_batch_vs_closed_form(ctx, oracle, ls):
    """twelve Cartesian derivative shell sets per quartet, a small batch per class (bra/ket classes in
    either order: the kernel-orientation swap), contracted shells"""
    from libint_b200 import capi
    po = oracle
    rng = np.random.default_rng(100 + sum(x * 5 ** i for i, x in enumerate(ls)))
    K = 1 if sum(ls) >= 6 else 2
    nb, nk = 2, 2
    bra_sh = _normalized(po, rng, [ls[0]] * nb + [ls[1]] * nb, K)
    ket_sh = _normalized(po, rng, [ls[2]] * nk + [ls[3]] * nk, K)
    Bb = capi.Basis(ctx, bra_sh.l, bra_sh.pure, bra_sh.nprim, bra_sh.O, bra_sh.alpha, bra_sh.coeff)
    Bk = capi.Basis(ctx, ket_sh.l, ket_sh.pure, ket_sh.nprim, ket_sh.O, ket_sh.alpha, ket_sh.coeff)
    bra = capi.Pairs(ctx, Bb, Bb, np.arange(nb), nb + np.arange(nb))
    ket = capi.Pairs(ctx, Bk, Bk, np.arange(nk), nk + np.arange(nk))
    tasks = np.array([(i, j) for i in range(nb) for j in range(nk)], dtype=np.int32)
    out = capi.eri_deriv1_batch(ctx, bra, ket, tasks)
    assert out.shape == (len(tasks), 12, nc(ls[0]) * nc(ls[1]) * nc(ls[2]) * nc(ls[3]))
    off_b, off_k = bra_sh.offsets(), ket_sh.offsets()
    for t, (i, j) in enumerate(tasks):
        idx_b, idx_k = [i, nb + i], [j, nk + j]
        sub_b, sub_k = bra_sh.subset(idx_b), ket_sh.subset(idx_k)
        q = po.Shells(np.concatenate([sub_b.l, sub_k.l]), np.zeros(4, dtype=np.int32),
                      np.concatenate([sub_b.nprim, sub_k.nprim]), np.concatenate([sub_b.O, sub_k.O]),
                      np.concatenate([sub_b.alpha, sub_k.alpha]), np.concatenate([sub_b.coeff, sub_k.coeff]),
                      raw=False)
        ref = po.deriv1_closed(q)
        assert_reference_thresholds(out[t], ref, "d(%d%d|%d%d) task %d" % (ls + (t,)))
        # translational invariance holds exactly by construction
        assert np.array_equal(out[t, 9:12], -(out[t, 0:3] + out[t, 3:6] + out[t, 6:9]))


def test_engine_mirror_deriv1_pure_and_permuted(ctx, oracle):
    """Engine(deriv_order=1).compute in the caller's shell order with pure shells: results()[3 * centre + xyz],
    centres in the caller's order (engine.impl.h:1996-2003), solid harmonics where flagged"""
    from libint_b200.basis import Shell
    from libint_b200.engine import Engine
    po = oracle
    rng = np.random.default_rng(77)
    for ls, pure in [((0, 1, 2, 1), (0, 0, 1, 0)), ((1, 2, 0, 2), (0, 1, 0, 1)), ((2, 2, 1, 1), (1, 1, 0, 0))]:
        l, pu, nprim, O, al, co = random_shell_table(rng, ls, 2, pure=pure)
        off = np.concatenate([[0], np.cumsum(nprim)])
        shells = [Shell(l[i], list(zip(al[off[i]:off[i + 1]], co[off[i]:off[i + 1]])), origin=O[i], pure=bool(pu[i]))
                  for i in range(4)]
        eng = Engine(max_nprim=2, max_l=max(ls), deriv_order=1, precision=0.0, ctx=ctx)
        res = eng.compute(*shells)
        assert len(res) == 12 and eng.results() is res
        ref = po.deriv1_closed(po.Shells(l, [0] * 4, nprim, O, al, co, raw=True)).reshape([12] + [nc(x) for x in ls])
        for ax in range(4):
            if pu[ax] and ls[ax] > 0:
                M = _sph_matrix(po, ls[ax])
                ref = np.moveaxis(np.tensordot(M, ref, axes=([1], [ax + 1])), 0, ax + 1)
        for d in range(12):
            assert_reference_thresholds(res[d], ref[d].ravel(), "Engine deriv %s set %d" % (ls, d))


def test_deriv1_batch_device_buffers_pure_output_and_edge_cases(ctx, oracle):
    """device-resident task list and output (torch), solid-harmonic output, an empty batch, a batch whose
    primitives are all screened out (zeros: the reference returns a null target, engine.impl.h:1781-1784), and a
    batch larger than one chunk of the derivative scratch"""
    import torch
    from libint_b200 import capi
    po = oracle
    rng = np.random.default_rng(808)
    nb = nk = 8
    bra_sh = _normalized(po, rng, [2] * nb + [1] * nb, 2, pure=[1] * nb + [0] * nb)
    ket_sh = _normalized(po, rng, [2] * nk + [0] * nk, 1, pure=[1] * nk + [0] * nk)
    Bb = capi.Basis(ctx, bra_sh.l, bra_sh.pure, bra_sh.nprim, bra_sh.O, bra_sh.alpha, bra_sh.coeff)
    Bk = capi.Basis(ctx, ket_sh.l, ket_sh.pure, ket_sh.nprim, ket_sh.O, ket_sh.alpha, ket_sh.coeff)
    bra = capi.Pairs(ctx, Bb, Bb, np.arange(nb), nb + np.arange(nb))
    ket = capi.Pairs(ctx, Bk, Bk, np.arange(nk), nk + np.arange(nk))
    tasks = np.array([(i, j) for i in range(nb) for j in range(nk)], dtype=np.int32)
    cart = capi.eri_deriv1_batch(ctx, bra, ket, tasks)                       # host buffers, Cartesian
    pure_h = capi.eri_deriv1_batch(ctx, bra, ket, tasks, pure_out=True)      # host buffers, solid harmonics
    assert cart.shape == (64, 12, 6 * 3 * 6 * 1) and pure_h.shape == (64, 12, 5 * 3 * 5 * 1)
    M = _sph_matrix(po, 2)
    ref = np.einsum("pa,tdabcx,qc->tdpbqx", M, cart.reshape(64, 12, 6, 3, 6, 1), M).reshape(64, 12, -1)
    np.testing.assert_allclose(pure_h, ref, rtol=1e-13, atol=1e-15)
    dev = torch.device("cuda", ctx.device)
    ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    t_dev = torch.from_numpy(tasks).to(dev)
    for pure_out, want in ((False, cart), (True, pure_h)):
        out = torch.empty(want.shape, dtype=torch.float64, device=dev)
        capi.eri_deriv1_batch(ctx, bra, ket, t_dev, out=out, pure_out=pure_out)
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), want)
    # empty batch
    assert capi.eri_deriv1_batch(ctx, bra, ket, np.zeros((0, 2), dtype=np.int32)).shape == (0, 12, 108)
    # everything screened out by the engine precision: zeros
    z = capi.eri_deriv1_batch(ctx, bra, ket, tasks, precision=1e30)
    assert not z.any()
    # many tasks: several chunks of the derivative scratch, identical blocks for repeated tasks
    big = np.tile(tasks, (700, 1))    # 44800 tasks: two chunks of the 512 MiB scratch (33 k tasks each)
    out = torch.empty((len(big), 12, 108), dtype=torch.float64, device=dev)
    capi.eri_deriv1_batch(ctx, bra, ket, torch.from_numpy(big).to(dev), out=out)
    torch.cuda.synchronize()
    assert torch.equal(out[:64], out[-64:]) and np.array_equal(out[64 * 600:64 * 601].cpu().numpy(), cart)


def test_deriv1_lmax_is_an_error(ctx, oracle):
    """(f p| would need a (g p| twin: LB200_ERR_LMAX, the analogue of LIBINT2_MAX_AM_eri1"""
    from libint_b200 import capi
    po = oracle
    sh = _normalized(po, np.random.default_rng(3), [3, 1, 0, 0], 1)
    B = capi.Basis(ctx, sh.l, sh.pure, sh.nprim, sh.O, sh.alpha, sh.coeff)
    bra = capi.Pairs(ctx, B, B, [0], [1])
    ket = capi.Pairs(ctx, B, B, [2], [3])
    with pytest.raises(capi.Lb200Error):
        capi.eri_deriv1_batch(ctx, bra, ket, np.array([[0, 0]], dtype=np.int32))


def _h2o(name):
    from libint_b200.basis import BasisSet, H2O_XYZ_ANGSTROM, atoms_from_tuples
    return BasisSet(name, atoms_from_tuples(H2O_XYZ_ANGSTROM))


@pytest.mark.parametrize("name", ["sto-3g", "6-31g*", "cc-pvdz"])
def test_forces_2body_vs_golden(ctx, name):
    """F2 of H2O for a seeded D against the committed oracle output (closed-form derivative integrals digested
    as compute_2body_fock_deriv<1> does; tests/golden/make_golden.py grad_h2o): Cartesian d (6-31G*) and
    pure d (cc-pVDZ, BASELINE configs[0]'s basis)"""
    from libint_b200.fock import FockBuilder
    d = np.load(os.path.join(GOLD, "grad_h2o.npz"))
    tag = name.replace("-", "").replace("*", "s")
    bs = _h2o(name)
    fb = FockBuilder(bs, ctx=ctx, rank=0, nranks=1)
    g, st = fb.forces_2body(d[tag + "_D"], precision=1e-16, use_schwarz=False, stats=True)
    ref = d[tag + "_F2"]
    ns = len(bs)
    npair = ns * (ns + 1) // 2
    assert st["nquartets"] == npair * (npair + 1) // 2
    assert np.abs(g - ref).max() <= 1e-11 * max(1.0, np.abs(ref).max()), (name, np.abs(g - ref).max())
    assert np.abs(g.sum(axis=0)).max() <= 1e-11 * np.abs(g).max()   # no net force


def test_forces_2body_finite_difference_and_ranks(ctx):
    """F2 = d/dR trace(G(D; R) D) at fixed D with G from the (parity-tested) GPU Fock build: water dimer,
    cc-pVDZ, Schwarz screening on; the partial gradients of two ranks add up to the whole"""
    from libint_b200.basis import Atom, BasisSet, H2O_XYZ_ANGSTROM, atoms_from_tuples
    from libint_b200.fock import FockBuilder
    a = atoms_from_tuples(H2O_XYZ_ANGSTROM)
    atoms = a + [Atom(x.atomic_number, x.x + 5.5, x.y + 0.4, x.z - 0.3) for x in a]
    bs = BasisSet("cc-pvdz", atoms)
    rng = np.random.default_rng(9)
    D = rng.standard_normal((bs.nbf, bs.nbf)) * 0.1
    D = 0.5 * (D + D.T)
    fb = FockBuilder(bs, ctx=ctx, rank=0, nranks=1)
    g = fb.forces_2body(D, precision=1e-14)
    parts = [FockBuilder(bs, ctx=ctx, rank=r, nranks=2).fock.gradient(D, bs.shell2atom, len(atoms), 1e-14,
                                                                      rank=r, nranks=2) for r in range(2)]
    assert np.abs(parts[0]).max() > 0 and np.abs(parts[1]).max() > 0
    assert np.abs(parts[0] + parts[1] - g).max() <= 1e-11 * np.abs(g).max()

    def energy(at):
        b = BasisSet("cc-pvdz", at)
        G = FockBuilder(b, ctx=ctx, rank=0, nranks=1)(D, precision=1e-14)
        return float((np.asarray(G) * D).sum())

    h = 1e-3
    for ia, x in [(0, 0), (1, 1), (4, 2)]:
        def moved(s):
            out = []
            for k, t in enumerate(atoms):
                c = [t.x, t.y, t.z]
                if k == ia:
                    c[x] += s
                out.append(Atom(t.atomic_number, *c))
            return out
        fd = (energy(moved(h)) - energy(moved(-h))) / (2 * h)
        assert abs(fd - g[ia, x]) <= 2e-5 * max(1.0, np.abs(g).max()), (ia, x, fd, g[ia, x])


def test_forces_2body_lmax(ctx):
    """def2-TZVP has f shells next to p/d shells: raised twins (g p| do not exist"""
    from libint_b200 import capi
    from libint_b200.fock import FockBuilder
    bs = _h2o("def2-tzvp")
    fb = FockBuilder(bs, ctx=ctx, rank=0, nranks=1)
    with pytest.raises(capi.Lb200Error):
        fb.forces_2body(np.eye(bs.nbf), precision=1e-12)


def test_pair_records_built_on_device_match_host(ctx):
    """ShellPair::init on the GPU (pairs_device.cu) against the host loop: same surviving primitive pairs, same
    records (exp / log of the CUDA math library are within an ulp of libm's), every screening method"""
    from libint_b200 import capi
    from libint_b200.basis import BasisSet, water_cluster
    bs = BasisSet("cc-pvdz", water_cluster(2, 1, 1))
    B = capi.Basis(ctx, *bs.flat())
    s1, s2 = capi.significant_pairs(B, 1e-12)
    l = np.array([s.l for s in bs])
    for la, lb in [(0, 0), (1, 0), (2, 1)]:
        m = ((l[s1] == la) & (l[s2] == lb)) | ((l[s1] == lb) & (l[s2] == la))
        a = np.where(l[s1[m]] >= l[s2[m]], s1[m], s2[m])
        b = np.where(l[s1[m]] >= l[s2[m]], s2[m], s1[m])
        for scr, lnp in [(capi.SCREEN_ORIGINAL, np.log(1e-12)), (capi.SCREEN_CONSERVATIVE, np.log(1e-12)),
                         (capi.SCREEN_SCHWARZ_INF, np.log(1e-20))]:
            blocks = []
            for dev_min in ("1", "1000000000"):
                os.environ["LB200_PAIRS_DEVICE_MIN"] = dev_min
                try:
                    blocks.append(capi.Pairs(ctx, B, B, a, b, screening=scr, ln_prec=lnp))
                finally:
                    del os.environ["LB200_PAIRS_DEVICE_MIN"]
            dev, host = blocks
            assert dev.nprimpair == host.nprimpair and dev.nprimpair > 0
            for i in range(0, len(a), max(1, len(a) // 25)):
                rd, rh = dev.get(i), host.get(i)
                assert rd.shape == rh.shape
                assert np.array_equal(rd[:, 7:], rh[:, 7:])     # same (p1, p2)
                # P: (a1 A + a2 B) / gamma contracts into an FMA on the GPU -- a few ulp of the
                # coordinate scale where the sum cancels; K, 1/gamma, nonsph: relative
                np.testing.assert_allclose(rd[:, :3], rh[:, :3], rtol=4e-15, atol=4e-15)
                # K = c exp(-rho |AB|^2) / gamma: the rounding of the exponent (|x| up to ~40) shows up |x| ulp large
                np.testing.assert_allclose(rd[:, 3], rh[:, 3], rtol=1e-13, atol=1e-300)
                np.testing.assert_allclose(rd[:, 4:6], rh[:, 4:6], rtol=4e-15, atol=1e-300)
                # ln_scr = -rho |AB|^2 + ln c_a + ln c_b is a sum of O(1..10) terms that may cancel
                np.testing.assert_allclose(rd[:, 6], rh[:, 6], rtol=4e-15, atol=1e-13)
            # and the integrals made from them
            tasks = np.array([(i, (3 * i) % len(a)) for i in range(min(64, len(a)))], dtype=np.int32)
            x = capi.eri_batch(ctx, dev, dev, tasks, precision=0.0)
            y = capi.eri_batch(ctx, host, host, tasks, precision=0.0)
            np.testing.assert_allclose(x, y, rtol=1e-13, atol=1e-15)
