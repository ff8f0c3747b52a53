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
