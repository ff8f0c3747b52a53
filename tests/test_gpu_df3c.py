"""3-centre (P|mu nu) class sweep (libint_b200.df3c) against the oracle's
Engine(Operator::coulomb, BraKet::xs_xx).compute2(dfbs[P], unit, obs[mu], obs[nu])
(tests/hartree-fock/hartree-fock++.cc:2241-2242) on a small alkane in the BASELINE configs[3] bases."""
import numpy as np
import pytest

from util import assert_parity

pytestmark = pytest.mark.gpu


def test_three_center_sweep_vs_oracle(oracle):
    import torch
    from libint_b200 import capi
    from libint_b200.basis import BasisSet, alkane
    from libint_b200.df3c import ThreeCenter
    po = oracle
    ctx = capi.Context(0)   # own context: the sweep runs on a torch stream that dies with this test
    atoms = alkane(2)
    obs, dfbs = BasisSet("def2-tzvp", atoms), BasisSet("def2-tzvp-jk", atoms)
    for b in (obs, dfbs):       # the sweep returns Cartesian shell sets
        b.set_pure(False)
    dev = torch.device("cuda", ctx.device)
    stream = torch.cuda.Stream(dev)
    ctx.set_stream(stream.cuda_stream)
    tc = ThreeCenter(ctx, obs, dfbs)
    assert tc.ntriplets() == len(dfbs) * tc.npairs
    assert max(c[0] for c in tc.classes()) == 4          # g functions in the JKFIT set
    got = {}

    def sink(cls, t0, n, view):
        got.setdefault(cls, []).append(view.cpu().numpy().copy())

    out = torch.empty(1 << 22, dtype=torch.float64, device=dev)   # 32 MiB: forces chunking
    with torch.cuda.stream(stream):
        n = tc.sweep(out, sink=sink)
        torch.cuda.synchronize()
    assert n == tc.ntriplets()
    rng = np.random.default_rng(3)
    ol, op, on, oO, oa, oc = obs.flat()
    dl, dp_, dn, dO, da, dc = dfbs.flat()
    ooff = np.concatenate([[0], np.cumsum(on)])
    doff = np.concatenate([[0], np.cumsum(dn)])
    checked = 0
    for kb, kk in tc.blocks():
        cls = (kb[0], kk[0], kk[1])
        bra, ket = tc.bras[kb], tc.kets[kk]
        blocks = np.concatenate(got[(kb, kk)], axis=0)
        assert blocks.shape[0] == bra.npair * ket.npair
        for _ in range(3):
            ip, jk = int(rng.integers(bra.npair)), int(rng.integers(ket.npair))
            P, a, b = int(bra.s1[ip]), int(ket.s1[jk]), int(ket.s2[jk])
            sh = po.Shells([dl[P], ol[a], ol[b]], [0, 0, 0], [dn[P], on[a], on[b]],
                           np.array([dO[P], oO[a], oO[b]]),
                           np.concatenate([da[doff[P]:doff[P + 1]], oa[ooff[a]:ooff[a + 1]], oa[ooff[b]:ooff[b + 1]]]),
                           np.concatenate([dc[doff[P]:doff[P + 1]], oc[ooff[a]:ooff[a + 1]], oc[ooff[b]:ooff[b + 1]]]),
                           raw=False)
            ref = po.compute2(sh, braket=1, precision=0.0)
            assert_parity(blocks[ip * ket.npair + jk], ref, "(%d s|%d %d) P=%d mu=%d nu=%d" % (cls + (P, a, b)),
                          rtol=1e-12, atol=1e-14)
            checked += 1
    assert checked == 3 * len(tc.blocks())
    assert len(tc.blocks()) > len(tc.classes())   # contraction buckets split the classes


def _oracle_df_tensors(po, obs, dfbs):
    """dense Zxy[ndf][n][n] and V[ndf][ndf] through the reference Engine (xs_xx / xs_xs), as the reference's
    DF set-up computes them (hartree-fock++.cc:2215-2262, :1517-1571)."""
    ol, op, on, oO, oa, oc = obs.flat()
    dl, dp_, dn, dO, da, dc = dfbs.flat()
    ooff = np.concatenate([[0], np.cumsum(on)])
    doff = np.concatenate([[0], np.cumsum(dn)])
    n, ndf = obs.nbf, dfbs.nbf
    Z = np.zeros((ndf, n, n))
    V = np.zeros((ndf, ndf))

    def shells(idx):   # idx: list of ("d"|"o", shell)
        l, pu, npm, O, al, co = [], [], [], [], [], []
        for kind, s in idx:
            if kind == "d":
                l.append(dl[s]); pu.append(dp_[s]); npm.append(dn[s]); O.append(dO[s])
                al.append(da[doff[s]:doff[s + 1]]); co.append(dc[doff[s]:doff[s + 1]])
            else:
                l.append(ol[s]); pu.append(op[s]); npm.append(on[s]); O.append(oO[s])
                al.append(oa[ooff[s]:ooff[s + 1]]); co.append(oc[ooff[s]:ooff[s + 1]])
        return po.Shells(l, pu, npm, np.array(O), np.concatenate(al), np.concatenate(co), raw=False)

    for P in range(len(dfbs)):
        p0, npf = dfbs.shell2bf[P], dfbs[P].size()
        for Q in range(len(dfbs)):
            q0, nqf = dfbs.shell2bf[Q], dfbs[Q].size()
            V[p0:p0 + npf, q0:q0 + nqf] = po.compute2(shells([("d", P), ("d", Q)]), braket=2, precision=0.0)
        for a in range(len(obs)):
            a0, na = obs.shell2bf[a], obs[a].size()
            for b in range(len(obs)):
                b0, nb = obs.shell2bf[b], obs[b].size()
                Z[p0:p0 + npf, a0:a0 + na, b0:b0 + nb] = po.compute2(shells([("d", P), ("o", a), ("o", b)]),
                                                                    braket=1, precision=0.0)
    return Z, V


def test_df_slab_metric_and_fock_vs_reference_formulas(ctx, oracle):
    """lb200_df3c_slab / lb200_df3c_metric against the reference Engine, and the streamed DF Fock builder
    (libint_b200.dfjk) against compute_2body_fock_dfC's dense formulas (hartree-fock++.cc:2264-2320:
    L = chol(V), xyK = Zxy L^-T, exchange + Coulomb contractions, G = 2J - K)."""
    import torch
    from libint_b200 import capi
    from libint_b200.basis import BasisSet, H2O_XYZ_ANGSTROM, atoms_from_tuples
    from libint_b200.dfjk import DFFockBuilder
    atoms = atoms_from_tuples(H2O_XYZ_ANGSTROM)
    obs, dfbs = BasisSet("cc-pvdz", atoms), BasisSet("cc-pvdz-ri", atoms)
    Zref, Vref = _oracle_df_tensors(oracle, obs, dfbs)
    n, ndf = obs.nbf, dfbs.nbf
    # small slabs (3 DF functions' worth of memory would be one shell): forces several slabs
    fb = DFFockBuilder(obs, dfbs, ctx=ctx, slab_bytes=8 * n * n * 40)
    assert len(fb.slabs) > 2
    dev = fb.dev
    V = torch.empty((ndf, ndf), dtype=torch.float64, device=dev)
    fb.df.metric(V)
    np.testing.assert_allclose(V.cpu().numpy(), Vref, rtol=1e-12, atol=1e-13)
    Z = torch.empty((ndf, n, n), dtype=torch.float64, device=dev)
    done, total = fb.df.slab(Z, 0, len(dfbs))
    assert done == total == len(dfbs) * fb.df.npairs
    np.testing.assert_allclose(Z.cpu().numpy(), Zref, rtol=1e-12, atol=1e-13)
    # a slab in the middle, with screening on: only negligible triplets may be dropped
    s0, ns, nf = fb.slabs[1]
    Zs = torch.empty((nf, n, n), dtype=torch.float64, device=dev)
    d2, t2 = fb.df.slab(Zs, s0, ns, threshold=1e-9)
    r0 = dfbs.shell2bf[s0]
    assert d2 <= t2
    assert np.max(np.abs(Zs.cpu().numpy() - Zref[r0:r0 + nf])) < 1e-8
    # DF Fock matrix
    rng = np.random.default_rng(5)
    nocc = 5
    C = np.linalg.qr(rng.standard_normal((n, nocc)))[0]
    G = fb(C).cpu().numpy()
    L = np.linalg.cholesky(Vref)
    Linv_t = np.linalg.inv(L).T
    xyK = np.einsum("Pxy,PQ->xyQ", Zref, Linv_t)
    xiK = np.einsum("xyK,yi->xiK", xyK, C)
    Kmat = np.einsum("xiK,yiK->xy", xiK, xiK)
    Jtmp = np.einsum("xiK,xi->K", xiK, C)
    Gref = 2.0 * np.einsum("xyK,K->xy", xyK, Jtmp) - Kmat
    np.testing.assert_allclose(G, Gref, rtol=1e-10, atol=1e-11)
    assert fb.stats["sweeps"] == 2
