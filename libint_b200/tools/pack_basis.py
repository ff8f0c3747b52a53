#!/usr/bin/env python3
"""Pack a few Gaussian-94 basis-set library files into the compact JSON tables shipped in
libint_b200/data/basis/ (the GPU box has no basis library; the reference looks its .g94 files
up through LIBINT_DATA_PATH / SRCDATADIR, include/libint2/basis.h.in:404-422).

  python -m libint_b200.tools.pack_basis <dir with *.g94> [names...]

Only the elements listed in ELEMENTS are kept.  Values are the numbers of the .g94 files
(raw contraction coefficients, not normalized); parsing is libint_b200.basis.read_g94.
"""
import json
import os
import sys

from libint_b200.basis import read_g94

ELEMENTS = (1, 2, 6, 7, 8, 9, 10)
NAMES = ("sto-3g", "6-31g", "6-31g*", "cc-pvdz", "augmentation-cc-pvdz", "cc-pvtz", "def2-svp",
         "def2-tzvp", "def2-tzvp-jk", "cc-pvdz-ri")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "data", "basis")


def main(argv):
    src = argv[1]
    names = argv[2:] or NAMES
    os.makedirs(OUT, exist_ok=True)
    for name in names:
        lib = read_g94(os.path.join(src, name + ".g94"))
        packed = {}
        for Z in ELEMENTS:
            if Z in lib:
                packed[str(Z)] = [[l, list(map(float, ex)), list(map(float, co))] for l, ex, co in lib[Z]]
        fn = os.path.join(OUT, name.replace("*", "s") + ".json")
        with open(fn, "w") as f:
            json.dump({"name": name, "shells": packed}, f, separators=(",", ":"))
        print(fn, {z: len(v) for z, v in packed.items()})


if __name__ == "__main__":
    main(sys.argv)
