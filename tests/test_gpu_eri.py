"""Parity of the CUDA ERI path (through the C ABI) with the CPU oracle and the committed
golden vectors.  Tolerance: 1e-12 relative / 1e-14 absolute (BASELINE.json north_star)."""
import itertools
import os

import numpy as np
import pytest

from util import (ATOL, PAIR_CLASSES, all_classes, assert_batch_close, assert_close_to_oracle, assert_parity,
                  assert_parity_screened, nc, pair_key, random_shell_table)

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# classes of the l<=3 (+ g-bra) grid without a kernel yet; must fail cleanly with an LMAX error
KNOWN_GAPS = set()


def _supported(capi, cl):
    ok = capi.eri_class_supported(*cl)
    assert ok or tuple(cl) in KNOWN_GAPS, "class %s lost its kernel" % (tuple(cl),)
    return ok


def _batch_one(ctx, capi, table, pure_out=False, **kw):
    l, pure, nprim, O, al, co = table
    bs = capi.Basis(ctx, l, pure, nprim, O, al, co)
    bra = capi.Pairs(ctx, bs, bs, [0], [1])
    ket = capi.Pairs(ctx, bs, bs, [2], [3])
    return capi.eri_batch(ctx, bra, ket, np.array([[0, 0]], dtype=np.int32), pure_out=pure_out, **kw)[0]


def test_every_class_vs_committed_goldens(ctx):
    """all (la lb|lc ld), la>=lb, lc>=ld, l<=3 plus (g s| bra/ket: 121 classes (any bra/ket
    order: the non-canonical orders exercise the transposed write-out)."""
    from libint_b200 import capi
    d = np.load(os.path.join(GOLD, "eri_classes.npz"))
    for ci, cl in enumerate(d["classes"]):
        K = int(d["c%d_K" % ci])
        table = (list(cl), [0] * 4, [K] * 4, d["c%d_O" % ci], d["c%d_alpha" % ci], d["c%d_coeff" % ci])
        if not _supported(capi, cl):
            with pytest.raises(capi.Lb200Error):
                _batch_one(ctx, capi, table)
            continue
        got = _batch_one(ctx, capi, table)
        assert_parity(got, d["c%d_eri" % ci], "class (%d%d|%d%d) vs golden" % tuple(cl))


@pytest.mark.parametrize("K", [1, 3])
def test_classes_vs_oracle_live(ctx, oracle, K):
    """fresh random contracted quartets per class, oracle computed on the box's CPU."""
    from libint_b200 import capi
    rng = np.random.default_rng(1000 + K)
    for cl in all_classes():
        if (K == 3 and sum(cl) > 9) or not _supported(capi, cl):
            continue
        table = random_shell_table(rng, cl, K)
        got = _batch_one(ctx, capi, table)
        osh = oracle.Shells(*table, raw=False)
        ref = oracle.compute2(osh, precision=0.0).ravel()
        assert_close_to_oracle(got, ref, osh, [0, 1, 2, 3], "class (%d%d|%d%d) K=%d" % (tuple(cl) + (K,)))


def test_pure_output_and_mixed_purity(ctx, oracle):
    """cart->pure transform (solidharmonics.h:281-463 via engine.impl.h:1965-1985), any
    purity pattern, including pure p shells."""
    from libint_b200 import capi
    rng = np.random.default_rng(77)
    cases = [((2, 2, 2, 2), (1, 1, 1, 1)), ((3, 1, 2, 0), (1, 0, 1, 0)), ((2, 1, 1, 1), (1, 1, 0, 1)),
             ((3, 3, 2, 2), (1, 0, 0, 1)), ((4, 0, 3, 2), (1, 0, 1, 1)), ((1, 0, 0, 0), (1, 0, 0, 0)),
             ((3, 2, 3, 1), (1, 1, 1, 0))]
    for cl, pure in cases:
        table = random_shell_table(rng, cl, 2, pure=pure)
        got = _batch_one(ctx, capi, table, pure_out=True)
        osh = oracle.Shells(*table, raw=False)
        ref = oracle.compute2(osh, precision=0.0).ravel()
        assert_close_to_oracle(got, ref, osh, [0, 1, 2, 3], "pure %s %s" % (cl, pure))


def test_three_center_vs_goldens(ctx):
    """xs_xx: (X s|c d) with Shell::unit() as bra2 (engine.impl.h:165-167,1836-1873)."""
    from libint_b200 import capi
    d = np.load(os.path.join(GOLD, "eri3_classes.npz"))
    unit = capi.Basis.unit(ctx)
    for ci, cl in enumerate(d["classes"]):
        for pv in (0, 1):
            pure = [int(pv and x > 1) for x in cl]
            bs = capi.Basis(ctx, list(cl), pure, [2] * 3, d["c%d_%d_O" % (ci, pv)],
                            d["c%d_%d_alpha" % (ci, pv)], d["c%d_%d_coeff" % (ci, pv)])
            bra = capi.Pairs(ctx, bs, unit, [0], [0])
            ket = capi.Pairs(ctx, bs, bs, [1], [2])
            got = capi.eri_batch(ctx, bra, ket, np.array([[0, 0]], dtype=np.int32), pure_out=True)[0]
            assert_parity(got, d["c%d_%d_eri" % (ci, pv)], "3-centre (%d s|%d %d) pure=%d" % (tuple(cl) + (pv,)))


def test_engine_mirror_permutations(ctx, oracle):
    """Engine.compute in any shell order == reference Engine (canonicalisation + un-permute,
    engine.impl.h:1183-1211,1988-2067; tests/unit/test-permute.cc)."""
    from libint_b200.basis import Shell
    from libint_b200.engine import BraKet, Engine
    rng = np.random.default_rng(5)
    ls = (0, 2, 1, 3)
    table = random_shell_table(rng, ls, 2, pure=[0, 1, 0, 1])
    l, pure, nprim, O, al, co = table
    shells = [Shell(l[i], list(zip(al[2 * i:2 * i + 2], co[2 * i:2 * i + 2])), O[i], pure=bool(pure[i]),
                    embed_normalization=False) for i in range(4)]
    eng = Engine(max_nprim=2, max_l=3, precision=0.0, ctx=ctx)
    osh = oracle.Shells(*table, raw=False)
    for perm in [(0, 1, 2, 3), (1, 0, 2, 3), (2, 3, 0, 1), (3, 2, 1, 0), (1, 0, 3, 2), (2, 3, 1, 0)]:
        got = eng.compute(*[shells[i] for i in perm])
        ref = oracle.compute2(osh.subset(list(perm)), precision=0.0).ravel()
        assert_close_to_oracle(got, ref, osh, list(perm), "perm %s" % (perm,))
    eng3 = Engine(max_nprim=2, max_l=3, precision=0.0, braket=BraKet.xs_xx, ctx=ctx)
    got = eng3.compute(shells[3], shells[1], shells[2])
    ref = oracle.compute2(osh.subset([3, 1, 2]), braket=1, precision=0.0).ravel()
    assert_parity(got, ref, "xs_xx engine")
    eng2 = Engine(max_nprim=2, max_l=3, precision=0.0, braket=BraKet.xs_xs, ctx=ctx)
    got = eng2.compute(shells[3], shells[1])
    ref = oracle.compute2(osh.subset([3, 1]), braket=2, precision=0.0).ravel()
    assert_parity(got, ref, "xs_xs engine")


def test_python_goldens_gpu(ctx):
    """python/tests/test_libint2.py:37-49 through the CUDA path."""
    from libint_b200.basis import Shell
    from libint_b200.engine import BraKet, Engine
    s = Shell(0, [(1.0, 10.0)])
    p = Shell(1, [(1.0, 10.0)])
    eng = Engine(max_nprim=1, max_l=1, ctx=ctx)
    assert np.linalg.norm(eng.compute(p, p, s, s)) == pytest.approx(1.62867503968, abs=5e-11)
    e3 = Engine(max_nprim=1, max_l=1, braket=BraKet.xs_xx, ctx=ctx)
    assert np.linalg.norm(e3.compute(s, s, s)) == pytest.approx(3.6563211198, abs=5e-11)
    basis = [s, p, s, p]
    tot = sum(np.sum(eng.compute(*[basis[i] for i in q]) ** 2) for q in itertools.product(range(4), repeat=4))
    assert np.sqrt(tot) == pytest.approx(14.7036075402, abs=5e-10)


def test_pair_data_matches_shellpair(ctx, oracle):
    """device pair records == ShellPair::init (shell.h:1138-1256), Original and Conservative,
    including which primitive pairs survive ln_prec."""
    from libint_b200 import capi
    rng = np.random.default_rng(21)
    for scr in (capi.SCREEN_ORIGINAL, capi.SCREEN_CONSERVATIVE):
        for (la, lb) in [(0, 0), (1, 0), (2, 1), (3, 3)]:
            table = random_shell_table(rng, (la, lb), (4, 3), spread=2.0, amax=12.0)
            l, pure, nprim, O, al, co = table
            bs = capi.Basis(ctx, l, pure, nprim, O, al, co)
            for ln_prec in (-1e300, np.log(1e-10), np.log(1e-4)):
                P = capi.Pairs(ctx, bs, bs, [0], [1], screening=scr, ln_prec=ln_prec)
                ref, AB = oracle.shellpair(oracle.Shells(*table, raw=False), ln_prec, scr)
                got = P.get(0)
                assert got.shape == ref.shape, (scr, la, lb, ln_prec)
                np.testing.assert_allclose(got, ref, rtol=2e-15, atol=0)


def test_primitive_screening_matches_engine(ctx, oracle):
    """Engine precision semantics (engine.impl.h:1313-1314,1371-1386): with the same pair data
    and precision the same primitive quartets are skipped, so results agree to parity
    tolerance, and |I_eps - I_0| <= 2 eps (tests/unit/test-precision.cc:82-216)."""
    from libint_b200 import capi
    rng = np.random.default_rng(33)
    for scr in (capi.SCREEN_ORIGINAL, capi.SCREEN_CONSERVATIVE):
        for cl in [(0, 0, 0, 0), (1, 0, 1, 0), (2, 1, 1, 1), (2, 2, 2, 0)]:
            table = random_shell_table(rng, cl, 3, spread=2.5, amax=8.0)
            l, pure, nprim, O, al, co = table
            bs = capi.Basis(ctx, l, pure, nprim, O, al, co)
            exact = _batch_one(ctx, capi, table)
            for eps in (1e-8, 1e-10, 1e-12):
                bra = capi.Pairs(ctx, bs, bs, [0], [1], screening=scr, ln_prec=np.log(eps))
                ket = capi.Pairs(ctx, bs, bs, [2], [3], screening=scr, ln_prec=np.log(eps))
                got = capi.eri_batch(ctx, bra, ket, np.array([[0, 0]], dtype=np.int32), screening=scr,
                                     precision=eps)[0]
                ref = oracle.compute2(oracle.Shells(*table, raw=False), precision=eps, screening=scr)
                ref = np.zeros_like(got) if ref is None else ref.ravel()
                assert_parity_screened(got, ref, oracle.Shells(*table, raw=False), [0, 1, 2, 3],
                                       "screen %x %s eps=%g" % (scr, cl, eps))
                if scr == capi.SCREEN_CONSERVATIVE:  # the bound the reference tests (Conservative only)
                    assert np.max(np.abs(got - exact)) <= 2 * eps


def test_empty_and_edge_inputs(ctx):
    from libint_b200 import capi
    rng = np.random.default_rng(2)
    table = random_shell_table(rng, (1, 0, 1, 0), 2)
    l, pure, nprim, O, al, co = table
    bs = capi.Basis(ctx, l, pure, nprim, O, al, co)
    bra = capi.Pairs(ctx, bs, bs, [0], [1])
    ket = capi.Pairs(ctx, bs, bs, [2], [3])
    out = capi.eri_batch(ctx, bra, ket, np.zeros((0, 2), dtype=np.int32))
    assert out.shape == (0, 9)
    # pairs of mixed classes in one block are rejected
    with pytest.raises(capi.Lb200Error):
        capi.Pairs(ctx, bs, bs, [0, 1], [1, 0])
    # l(s1) < l(s2) is rejected (caller canonicalises, engine.impl.h:1183-1188)
    with pytest.raises(capi.Lb200Error):
        capi.Pairs(ctx, bs, bs, [1], [0])
    # angular momentum beyond LB200_MAX_AM -> LMAX error, like Engine::lmax_exceeded
    with pytest.raises(capi.Lb200Error):
        capi.Basis(ctx, [5], [0], [1], [[0, 0, 0]], [1.0], [1.0])
    # same-centre quartet: exact zeros by symmetry are returned as zeros, not skipped
    t2 = ([1, 0, 0, 0], [0] * 4, [1] * 4, np.zeros((4, 3)), np.ones(4), np.ones(4))
    assert np.all(_batch_one(ctx, capi, t2) == 0.0)
    # all primitive pairs screened out -> zeros (results()[0] == nullptr in the reference)
    far = ([0, 0, 0, 0], [0] * 4, [1] * 4, np.array([[0, 0, 0], [60.0, 0, 0], [0, 0, 0], [0, 0, 1.0]]),
           np.full(4, 5.0), np.ones(4))
    bs2 = capi.Basis(ctx, *far)
    bra2 = capi.Pairs(ctx, bs2, bs2, [0], [1], ln_prec=np.log(1e-12))
    assert bra2.nprimpair == 0
    ket2 = capi.Pairs(ctx, bs2, bs2, [2], [3], ln_prec=np.log(1e-12))
    assert np.all(capi.eri_batch(ctx, bra2, ket2, np.array([[0, 0]], dtype=np.int32), precision=1e-12) == 0)


def test_large_batch_properties(ctx, oracle):
    """BASELINE config 2 shape at reduced count: many random primitive quartets of one class;
    (i) a subsample equals the oracle, (ii) bra<->ket swap symmetry (ab|cd) == (cd|ab) holds
    for the whole batch (tests/unit/test-2body.cc:126-189), (iii) device-resident torch
    buffers give the same bits as host buffers."""
    import torch
    from libint_b200 import capi
    rng = np.random.default_rng(99)
    nsh = 64
    for (la, lb, lc, ld) in [(0, 0, 0, 0), (1, 1, 1, 0), (2, 1, 2, 0), (2, 2, 2, 2)]:
        ls = [la] * nsh + [lb] * nsh + [lc] * nsh + [ld] * nsh
        table = random_shell_table(rng, ls, 1, spread=3.0, amin=0.1, amax=6.0)
        l, pure, nprim, O, al, co = table
        bs = capi.Basis(ctx, l, pure, nprim, O, al, co)
        i = np.arange(nsh, dtype=np.int32)
        bra = capi.Pairs(ctx, bs, bs, i, nsh + i)
        ket = capi.Pairs(ctx, bs, bs, 2 * nsh + i, 3 * nsh + i)
        ntask = 20000
        tasks = rng.integers(0, nsh, (ntask, 2)).astype(np.int32)
        got = capi.eri_batch(ctx, bra, ket, tasks)
        blk = nc(la) * nc(lb) * nc(lc) * nc(ld)
        assert got.shape == (ntask, blk)
        osh = oracle.Shells(*table, raw=False)
        # the whole batch against the reference Engine and the extended-precision arbiter
        q4 = np.stack([tasks[:, 0], nsh + tasks[:, 0], 2 * nsh + tasks[:, 1], 3 * nsh + tasks[:, 1]], axis=1)
        nthr = os.cpu_count() or 1
        orc = oracle.compute_batch(osh, q4, nthreads=nthr)
        hi, lo = oracle.truth_batch(osh, q4, nthreads=nthr)
        assert_batch_close(got, orc, hi, lo, "batch (%d%d|%d%d)" % (la, lb, lc, ld))
        swapped = capi.eri_batch(ctx, ket, bra, tasks[:, ::-1].copy())
        sw = swapped.reshape(ntask, nc(lc) * nc(ld), nc(la) * nc(lb)).transpose(0, 2, 1).reshape(ntask, blk)
        assert_batch_close(sw, orc, hi, lo, "bra<->ket swapped batch (%d%d|%d%d)" % (la, lb, lc, ld))
        tt = torch.from_numpy(tasks).cuda()
        out = torch.empty((ntask, blk), dtype=torch.float64, device="cuda")
        capi.eri_batch(ctx, bra, ket, tt, out=out)
        ctx.synchronize()
        assert np.array_equal(out.cpu().numpy(), got)


def test_boys_branches_through_ss(ctx, oracle):
    """(ss|ss) over separations that hit every Boys branch: interpolation table, T just below /
    above 117, asymptotic (boys.h:345-454)."""
    from libint_b200 import capi
    for x in (0.0, 0.3, 2.0, 7.61, 7.65, 7.7, 12.0, 40.0):
        O = np.array([[0, 0, 0], [0, 0, 0.1], [x, 0, 0], [x, 0.1, 0]], dtype=float)
        for m_cl in [(0, 0, 0, 0), (2, 2, 2, 2), (3, 3, 3, 2)]:
            table = (list(m_cl), [0] * 4, [1] * 4, O, np.array([2.0, 2.0, 2.0, 2.0]), np.ones(4))
            got = _batch_one(ctx, capi, table)
            ref = oracle.compute2(oracle.Shells(*table, raw=False), precision=0.0).ravel()
            assert_parity(got, ref, "boys x=%g class %s" % (x, m_cl))


def test_uncontracted_pipeline_ragged(ctx, oracle):
    """The pipelined uncontracted kernel (eri_rowreg_prim.cuh) on ragged input: batches that are not
    a multiple of the quartets per CTA (1, 2, 7, 1001 tasks), pairs whose only primitive pair is
    screened out by ShellPair::init, quartets dropped by the primitive screen at a finite engine
    precision, both output orientations -- against the reference Engine with the same precision."""
    from libint_b200 import capi
    rng = np.random.default_rng(2024)
    nsh = 12
    for cl in [(1, 0, 1, 0), (1, 1, 1, 1), (2, 1, 2, 0), (2, 2, 2, 1), (2, 2, 2, 2), (3, 2, 1, 0), (3, 3, 2, 1)]:
        la, lb, lc, ld = cl
        ls = [la] * nsh + [lb] * nsh + [lc] * nsh + [ld] * nsh
        table = random_shell_table(rng, ls, 1, spread=4.0, amin=0.3, amax=12.0)
        l, pure, nprim, O, al, co = table
        O[1] = O[0] + np.array([9.0, 0, 0])          # a far, tight pair: nothing survives ShellPair::init
        O[nsh + 1] = O[0] + np.array([-9.0, 0, 0])
        al[1] = al[nsh + 1] = 11.0
        eps = 1e-9
        bs = capi.Basis(ctx, l, pure, nprim, O, al, co)
        i = np.arange(nsh, dtype=np.int32)
        bra = capi.Pairs(ctx, bs, bs, i, nsh + i, ln_prec=np.log(eps))
        ket = capi.Pairs(ctx, bs, bs, 2 * nsh + i, 3 * nsh + i, ln_prec=np.log(eps))
        assert bra.nprimpair < nsh                       # the far pair kept no primitive
        osh = oracle.Shells(*table, raw=False)
        blk = nc(la) * nc(lb) * nc(lc) * nc(ld)
        for ntask in (1, 2, 7, 1001):
            tasks = rng.integers(0, nsh, (ntask, 2)).astype(np.int32)
            tasks[0] = (1, 3)                            # the empty pair is always present
            got = capi.eri_batch(ctx, bra, ket, tasks, precision=eps)
            swapped = capi.eri_batch(ctx, ket, bra, tasks[:, ::-1].copy(), precision=eps)
            assert np.all(got[0] == 0.0)
            nzero = 0
            for t in range(min(ntask, 25)):
                b, k = tasks[t]
                ref = oracle.compute2(osh.subset([b, nsh + b, 2 * nsh + k, 3 * nsh + k]), precision=eps)
                ref = np.zeros(blk) if ref is None else ref.ravel()
                nzero += int(not ref.any())
                idx = [b, nsh + b, 2 * nsh + k, 3 * nsh + k]
                assert_parity_screened(got[t], ref, osh, idx, "ragged %s n=%d task %d" % (cl, ntask, t))
                sw = swapped[t].reshape(nc(lc) * nc(ld), nc(la) * nc(lb)).T.ravel()
                assert_parity_screened(sw, ref, osh, idx, "ragged swapped %s n=%d task %d" % (cl, ntask, t))
            assert nzero >= 1
