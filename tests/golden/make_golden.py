#!/usr/bin/env python3
"""Generate the committed golden vectors of the Coulomb-ERI / Fock path.

Run in the build container (needs /root/reference for oracle/_ref/liboracle.so, i.e. the
reference's own Engine / ShellPair / FmEval_Chebyshev7 / eri() compiled where they lie):

    python tests/golden/make_golden.py

    python tests/golden/make_golden.py grad_h2o      # one file only

Outputs (small, committed): tests/golden/eri_classes.npz, eri3_classes.npz, boys.npz,
fock_h2o.npz, closed_form.npz, grad_h2o.npz.  The GPU box has no reference tree; there the parity tests
compare the CUDA path against these files and against the prebuilt oracle library.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

from oracle import pyoracle as po  # noqa: E402
from util import all_classes, cartesianized, nc, random_shell_table  # noqa: E402


def eri_classes():
    rng = np.random.default_rng(20240607)
    out = {}
    classes = all_classes()
    out["classes"] = np.array(classes, dtype=np.int32)
    for ci, cl in enumerate(classes):
        K = 2 if sum(cl) <= 8 else 1
        l, pure, nprim, O, al, co = random_shell_table(rng, cl, K)
        ref = po.compute2(po.Shells(l, pure, nprim, O, al, co, raw=False), precision=0.0)
        out["c%d_O" % ci] = O
        out["c%d_alpha" % ci] = al
        out["c%d_coeff" % ci] = co
        out["c%d_K" % ci] = np.int32(K)
        out["c%d_eri" % ci] = ref.ravel()
    np.savez_compressed(os.path.join(HERE, "eri_classes.npz"), **out)
    print("eri_classes.npz:", len(classes), "classes")


def eri3_classes():
    """(X s|c d) with the unit shell as bra2 (engine.impl.h:165-167), pure and Cartesian."""
    rng = np.random.default_rng(31)
    out = {}
    cls = [(X, c, d) for X in range(5) for c in range(4) for d in range(c + 1)]
    out["classes"] = np.array(cls, dtype=np.int32)
    for ci, cl in enumerate(cls):
        for pv in (0, 1):
            pure = [pv and (x > 1) for x in cl]
            l, pure, nprim, O, al, co = random_shell_table(rng, cl, 2, pure=[int(p) for p in pure])
            ref = po.compute2(po.Shells(l, pure, nprim, O, al, co, raw=False), braket=1, precision=0.0)
            out["c%d_%d_O" % (ci, pv)] = O
            out["c%d_%d_alpha" % (ci, pv)] = al
            out["c%d_%d_coeff" % (ci, pv)] = co
            out["c%d_%d_eri" % (ci, pv)] = ref.ravel()
    np.savez_compressed(os.path.join(HERE, "eri3_classes.npz"), **out)
    print("eri3_classes.npz:", len(cls), "classes")


def boys():
    """FmEval_Chebyshev7 (boys.h:345-454) on a T grid covering every branch."""
    # T == 117.0 exactly is left out: the reference indexes interval 819 of an 819-interval
    # table there (boys.h:348 tests x > T_crit, :363 iv = int(x/delta)), i.e. reads past the end
    T = np.concatenate([[0.0, 1e-12, 1e-5], 10 ** np.linspace(-3, 2.3, 400),
                        [116.9999, np.nextafter(117.0, 0.0), 117.0001, 130.0, 1e3, 1e5]])
    mmax = 16
    F = np.array([po.boys_cheb7(t, mmax, 24) for t in T])
    Fref = np.array([po.boys_reference(t, mmax) for t in T])
    np.savez_compressed(os.path.join(HERE, "boys.npz"), T=T, cheb7=F, reference2=Fref)
    print("boys.npz:", F.shape, "max rel dev cheb7 vs reference2:",
          np.max(np.abs(F - Fref) / np.maximum(np.abs(Fref), 1e-300)))


def closed_form():
    """Primitive integrals from the reference's independent closed form eri()
    (src/bin/test_eri/eri.h:121-380), un-normalized Gaussians (norm_flag 0)."""
    rng = np.random.default_rng(5)
    rows = []
    for _ in range(200):
        ls = rng.integers(0, 4, 4)
        lmn = []
        for l in ls:
            x = rng.integers(0, l + 1)
            y = rng.integers(0, l - x + 1)
            lmn += [x, y, l - x - y]
        alpha = rng.uniform(0.1, 3.0, 4)
        centers = rng.uniform(0.7, 1.3, 12) - 1.0
        v = po.eri_closed(lmn, alpha, centers, 0)
        rows.append(np.concatenate([lmn, alpha, centers, [v]]))
    np.savez_compressed(os.path.join(HERE, "closed_form.npz"), rows=np.array(rows))
    print("closed_form.npz:", len(rows))


def fock_h2o():
    """G(D) of compute_2body_fock (hartree-fock++.cc:1574-1772) for H2O / cc-pVDZ and
    a seeded symmetric D; the pair list is all pairs (every H2O pair passes the overlap
    screen of hartree-fock++.cc:1353-1361)."""
    from libint_b200.basis import BasisSet, read_dotxyz, H2O_XYZ_ANGSTROM, atoms_from_tuples
    out = {}
    for name in ("cc-pvdz", "6-31g", "sto-3g"):
        atoms = atoms_from_tuples(H2O_XYZ_ANGSTROM)
        bs = BasisSet(name, atoms)
        sh = po.Shells(*bs.flat(), raw=False)
        ns = len(bs)
        s1, s2 = np.array([(a, b) for a in range(ns) for b in range(a + 1)], dtype=np.int32).T
        f = po.Fock(sh, s1, s2, nthreads=4)
        rng = np.random.default_rng(11)
        D = rng.standard_normal((bs.nbf, bs.nbf)) * 0.3
        D = 0.5 * (D + D.T)
        G, st = f.build(D, 1e-12)
        tag = name.replace("-", "")
        out[tag + "_D"] = D
        out[tag + "_G"] = G
        out[tag + "_K"] = f.schwarz()
        out[tag + "_nquartets"] = st["nquartets"]
    np.savez_compressed(os.path.join(HERE, "fock_h2o.npz"), **out)
    print("fock_h2o.npz done")


def grad_h2o():
    """Two-body forces F2 of hartree-fock++.cc:642-656 (compute_2body_fock_deriv<1> traced with D) for H2O
    and a seeded symmetric D, from the reference's closed-form derivative integrals (oracle
    lbo_fock_grad_closed); no screening."""
    from libint_b200.basis import BasisSet, H2O_XYZ_ANGSTROM, atoms_from_tuples
    out = {}
    for name in ("sto-3g", "6-31g*", "cc-pvdz"):
        atoms = atoms_from_tuples(H2O_XYZ_ANGSTROM)
        bs = BasisSet(name, atoms)
        rng = np.random.default_rng(23)
        D = rng.standard_normal((bs.nbf, bs.nbf)) * 0.2
        D = 0.5 * (D + D.T)
        sh, Dc = cartesianized(po, bs, D)
        g = po.fock_grad_closed(sh, Dc, bs.shell2atom, len(atoms), nthreads=os.cpu_count() or 4)
        tag = name.replace("-", "").replace("*", "s")
        out[tag + "_D"] = D
        out[tag + "_F2"] = g
        print(name, g)
    np.savez_compressed(os.path.join(HERE, "grad_h2o.npz"), **out)
    print("grad_h2o.npz done")


if __name__ == "__main__":
    which = sys.argv[1:] or ["eri_classes", "eri3_classes", "boys", "closed_form", "fock_h2o", "grad_h2o"]
    for name in which:
        globals()[name]()
