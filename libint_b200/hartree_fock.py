"""Command-line RHF driver with the reference's interface and output format:

    python -m libint_b200.hartree_fock [geometry.xyz] [basis] [--codata2010]

mirrors `hartree-fock++ [geometry.xyz] [basis]` (tests/hartree-fock/hartree-fock++.cc:233-244:
default geometry h2o.xyz, default basis aug-cc-pVDZ) and prints the lines the reference's
validation scripts parse (`** Hartree-Fock energy = ...`, hartree-fock++-validate.py:60-70,
hartree-fock-validate.py:19-24; `** 1-body forces = ...` to `** Hartree-Fock forces = ...`,
hartree-fock++-validate.py:128-132), so those scripts can be pointed at this driver unchanged.
S, T, V come from lb200_onebody, the two-electron part of every Fock matrix from lb200_fock_build, the one-body
and Pulay forces from lb200_onebody_forces, the two-body forces from lb200_fock_grad (all on the GPU); `--codata2010` converts Angstrom with the constant the plain `hartree-fock` test uses
(hartree-fock.cc:306) instead of libint2's CODATA-2018 default.
"""
import argparse
import sys
import time

import numpy as np

from . import basis as B
from .fock import FockBuilder
from . import capi
from .scf import RHF, hf_forces


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m libint_b200.hartree_fock")
    ap.add_argument("geometry", nargs="?", default=None, help=".xyz file (Angstrom); default: h2o.xyz of the reference")
    ap.add_argument("basis", nargs="?", default="aug-cc-pVDZ")
    ap.add_argument("--codata2010", action="store_true")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--no-forces", action="store_true")
    args = ap.parse_args(argv)
    b2a = B.BOHR_TO_ANGSTROM_CODATA2010 if args.codata2010 else B.BOHR_TO_ANGSTROM
    if args.geometry is None:
        atoms = [B.Atom(Z, r[0] / b2a, r[1] / b2a, r[2] / b2a) for Z, r in B.H2O_XYZ_ANGSTROM]
    else:
        atoms = B.read_dotxyz(args.geometry, b2a)
    obs = B.BasisSet(args.basis, atoms)
    print("Atomic Cartesian coordinates (a.u.):")
    for a in atoms:
        print("%d %.10f %.10f %.10f" % (a.atomic_number, a.x, a.y, a.z))
    print("orbital basis set rank = %d" % obs.nbf)
    fb = FockBuilder(obs, device=args.device, rank=0, nranks=1)
    charges = [(float(a.atomic_number), a.xyz) for a in atoms]
    scf = RHF(obs, atoms, lambda D, prec: fb.build_partial(np.ascontiguousarray(D), prec),
              stv=capi.onebody(fb.ctx, fb.basis, charges))   # S, T, V from lb200_onebody
    print("Nuclear repulsion energy = %.12f" % scf.enuc)
    print("\n\nIter         E(HF)                 D(E)/E         RMS([F,D])/nn       Time(s)")
    t0 = time.time()
    e = scf.run()
    for it, etot, ediff, rms in scf.history:
        print(" %02d %20.12f %20.12e %20.12e" % (it, etot, ediff, rms))
    print("SCF wall time %.3f s, %s" % (time.time() - t0, "converged" if scf.converged else "NOT converged"))
    print("** Hartree-Fock energy = %20.12f" % e)
    # forces, in the reference's output format (hartree-fock++.cc:613-716; parsed by
    # hartree-fock++-validate.py:128-132).  A basis whose raised derivative classes have no kernel
    # (f shells next to p/d shells) skips them, like a reference library built with LIBINT2_DERIV_ERI_ORDER 0.
    if not args.no_forces:
        try:
            f = hf_forces(scf, fb)
        except capi.Lb200Error as err:
            print("forces skipped: %s" % err)
        else:
            for key in ("1-body", "Pulay", "2-body", "nuclear repulsion", "Hartree-Fock"):
                print("** %s forces = %s " % (key, " ".join("%.15g" % x for x in f[key].ravel())))
    return 0 if scf.converged else 1


if __name__ == "__main__":
    sys.exit(main())
