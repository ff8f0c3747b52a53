"""Host-side mirror of the reference's Shell / BasisSet / Atom API for the Coulomb-ERI path.

What it mirrors (evaleev/libint @ 7a1a9d8):
  * libint2::Atom, read_dotxyz            include/libint2/atom.h:41-44,83-160 (Angstrom -> bohr
                                          with the CODATA-2018 constant, :53)
  * libint2::Shell                        include/libint2/shell.h:720-1012 (one contraction per
                                          shell here, as Engine::compute2 requires,
                                          engine.impl.h:1167-1171); normalization = Shell::renorm
                                          (:958-999), done by the C ABI (lb200_shell_renorm)
  * libint2::BasisSet(name, atoms)        include/libint2/basis.h.in: G94 reader :473-617,
                                          SP splitting, pure iff l > 1 (l > 2 for the Pople
                                          Cartesian-d sets :368-386), aug-cc-pVXZ =
                                          cc-pVXZ + augmentation-cc-pVXZ (:388-400),
                                          shell2bf / nbf / max_nprim / max_l
  * python binding spelling               python/src/libint2/libint2.cc (Shell(l, [(exp, coeff)..],
                                          origin), BasisSet(name, atoms), basis.pure = ...)
Basis data: libint_b200/data/basis/*.json (H, He, C-Ne subsets packed by tools/pack_basis.py) or
any directory of .g94 files given by LIBINT_DATA_PATH (<path>/basis/<name>.g94), as in the
reference (basis.h.in:404-422).
"""
import json
import math
import os
import re

import numpy as np

BOHR_TO_ANGSTROM = 0.529177210903  # CODATA 2018, atom.h:53
ANGSTROM_TO_BOHR = 1 / BOHR_TO_ANGSTROM

_SYMBOLS = ["X", "H", "He", "Li", "Be", "B", "C", "N", "O", "F", "Ne", "Na", "Mg", "Al", "Si", "P",
            "S", "Cl", "Ar", "K", "Ca", "Sc", "Ti", "V", "Cr", "Mn", "Fe", "Co", "Ni", "Cu", "Zn",
            "Ga", "Ge", "As", "Se", "Br", "Kr"]
_Z_OF = {s.lower(): z for z, s in enumerate(_SYMBOLS)}
_AM = "spdfghikl"
_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "basis")


class Atom:
    """libint2::Atom: atomic number and position in bohr."""
    __slots__ = ("atomic_number", "x", "y", "z")

    def __init__(self, atomic_number, x, y, z):
        self.atomic_number, self.x, self.y, self.z = int(atomic_number), float(x), float(y), float(z)

    @property
    def xyz(self):
        return (self.x, self.y, self.z)


def read_dotxyz(path_or_text, bohr_to_angstrom=BOHR_TO_ANGSTROM):
    """libint2::read_dotxyz: XYZ file (Angstrom) -> list of Atom in bohr (atom.h:83-160)."""
    text = open(path_or_text).read() if os.path.exists(str(path_or_text)) else str(path_or_text)
    lines = text.splitlines()
    natom = int(lines[0].split()[0])
    atoms = []
    a2b = 1 / bohr_to_angstrom
    for ln in lines[2:2 + natom]:
        sym, x, y, z = ln.split()[:4]
        if sym.lower() not in _Z_OF:
            raise ValueError("read_dotxyz: element symbol \"%s\" is not recognized" % sym)
        atoms.append(Atom(_Z_OF[sym.lower()], float(x) * a2b, float(y) * a2b, float(z) * a2b))
    if len(atoms) != natom:
        raise ValueError("read_dotxyz: expected %d atoms" % natom)
    return atoms


def atoms_from_tuples(tuples, angstrom=True):
    """[(Z, [x, y, z]), ...] as the reference's python tests write geometries (Angstrom)."""
    f = ANGSTROM_TO_BOHR if angstrom else 1.0
    return [Atom(Z, r[0] * f, r[1] * f, r[2] * f) for Z, r in tuples]


class Shell:
    """One contracted Gaussian shell with a single contraction (shell.h:720).

    `coeff` holds normalization-embedded coefficients exactly as Shell::contr[0].coeff does
    after Shell::renorm(); `raw_coeff` keeps the input."""

    do_enforce_unit_normalization = True  # Shell::do_enforce_unit_normalization(), shell.h:895-902

    def __init__(self, l, primitives, origin=(0.0, 0.0, 0.0), pure=None, embed_normalization=True):
        from . import capi
        self.l = int(l)
        self.alpha = np.array([p[0] for p in primitives], dtype=np.float64)
        self.raw_coeff = np.array([p[1] for p in primitives], dtype=np.float64)
        self.O = np.array(origin, dtype=np.float64)
        self.pure = bool(self.l > 1 if pure is None else pure)
        if embed_normalization:
            self.coeff, self.max_ln_coeff = capi.shell_renorm(
                self.l, self.alpha, self.raw_coeff, Shell.do_enforce_unit_normalization)
        else:
            self.coeff = self.raw_coeff.copy()
            with np.errstate(divide="ignore"):
                self.max_ln_coeff = np.log(np.abs(self.coeff))

    @classmethod
    def unit(cls):
        """Shell::unit(), shell.h:906-909,949-953."""
        return cls(0, [(0.0, 1.0)], pure=False, embed_normalization=False)

    @property
    def nprim(self):
        return len(self.alpha)

    def size(self):
        return 2 * self.l + 1 if self.pure else (self.l + 1) * (self.l + 2) // 2

    def cartesian_size(self):
        return (self.l + 1) * (self.l + 2) // 2

    def moved(self, origin):
        s = object.__new__(Shell)
        s.l, s.alpha, s.raw_coeff, s.pure = self.l, self.alpha, self.raw_coeff, self.pure
        s.coeff, s.max_ln_coeff = self.coeff, self.max_ln_coeff
        s.O = np.array(origin, dtype=np.float64)
        return s


def _fortran_float(tok):
    return float(tok.replace("D", "E").replace("d", "e"))


def read_g94(path):
    """Gaussian-94 library file -> {Z: [(l, exps, coeffs), ...]} with SP shells split
    (semantics of BasisSet::read_g94_basis_library, basis.h.in:473-617)."""
    out = {}
    with open(path) as f:
        lines = [ln.strip() for ln in f]
    i, n = 0, len(lines)
    Z = None
    expect_element = True
    first_element = True
    while i < n:
        ln = lines[i]
        i += 1
        if not ln or ln[0] == "!":
            continue
        if ln == "****":
            if first_element:
                continue
            expect_element = True
            continue
        if expect_element:
            sym = ln.split()[0]
            if sym.lower() not in _Z_OF:
                # elements beyond the table are skipped up to the next ****
                Z = None
            else:
                Z = _Z_OF[sym.lower()]
                out.setdefault(Z, [])
            expect_element = False
            first_element = False
            continue
        tok = ln.split()
        label, nprim = tok[0].lower(), int(tok[1])
        rows = []
        while len(rows) < nprim:
            r = lines[i]
            i += 1
            if not r or r[0] == "!":
                continue
            rows.append([_fortran_float(t) for t in r.split()])
        if Z is None:
            continue
        exps = [r[0] for r in rows]
        if label == "sp":
            out[Z].append((0, exps, [r[1] for r in rows]))
            out[Z].append((1, exps, [r[2] for r in rows]))
        else:
            if label == "j":
                l = 7
            else:
                l = _AM.index(label)
                if l >= 7:
                    l += 1  # Gaussian's K means L etc. (basis.h.in:540-546)
            out[Z].append((l, exps, [r[1] for r in rows]))
    return out


def _canonical(name):
    return name.strip().lower()


def gaussian_cartesian_d_convention(cname):
    """basis.h.in:368-386: 3-21G, 4-31G and 6-31G families use Cartesian d shells."""
    if cname.startswith("3-21") or cname.startswith("4-31g"):
        return True
    if cname.startswith("6-31") and len(cname) > 4 and cname[4] != "1":
        g = cname.find("g")
        if g < 0:
            return False
        if g + 1 == len(cname):
            return True
        if cname[g + 1] in "*s":
            return True
    return False


def decompose_name_into_components(name):
    """basis.h.in:388-400."""
    if name.startswith("aug-cc-pv") and "cabs" not in name:
        base = name[4:]
        return [base, "augmentation-" + base]
    return [name]


def _load_component(cname):
    path = os.environ.get("LIBINT_DATA_PATH")
    if path:
        fn = os.path.join(path, "basis", cname + ".g94")
        if os.path.exists(fn):
            return read_g94(fn)
    fn = os.path.join(_DATA, cname.replace("*", "s") + ".json")
    if os.path.exists(fn):
        with open(fn) as f:
            d = json.load(f)
        return {int(z): [(s[0], s[1], s[2]) for s in v] for z, v in d["shells"].items()}
    raise FileNotFoundError(
        "BasisSet: basis \"%s\" not found (set LIBINT_DATA_PATH to a directory holding "
        "basis/%s.g94)" % (cname, cname))


class BasisSet(list):
    """std::vector<Shell> for a molecule (BasisSet(name, atoms), basis.h.in:99-148)."""

    def __init__(self, name=None, atoms=None, shells=None, throw_if_no_match=True):
        super().__init__()
        self.name = name
        if shells is not None:
            self.extend(shells)
            self.shell2atom = [-1] * len(self)
        elif name is not None:
            cname = _canonical(name)
            force_cart_d = gaussian_cartesian_d_convention(cname)
            comps = [_load_component(c) for c in decompose_name_into_components(cname)]
            self.shell2atom = []
            proto = {}
            for ia, a in enumerate(atoms):
                Z = a.atomic_number
                if Z not in proto:
                    lst = []
                    for comp in comps:
                        if Z not in comp:
                            if throw_if_no_match:
                                raise KeyError("BasisSet: basis %s lacks element Z=%d" % (name, Z))
                            continue
                        for l, ex, co in comp[Z]:
                            pure = (l > 2) if force_cart_d else (l > 1)
                            lst.append(Shell(l, list(zip(ex, co)), pure=pure))
                    proto[Z] = lst
                for sh in proto[Z]:
                    self.append(sh.moved(a.xyz))
                    self.shell2atom.append(ia)
        self._refresh()

    def _refresh(self):
        self.shell2bf = []
        n = 0
        for s in self:
            self.shell2bf.append(n)
            n += s.size()
        self.nbf = n
        self.max_nprim = max([s.nprim for s in self], default=0)
        self.max_l = max([s.l for s in self], default=0)

    def set_pure(self, solid):
        """BasisSet::set_pure (basis.h.in:165-171)."""
        for s in self:
            s.pure = bool(solid)
        self._refresh()

    def flat(self):
        """-> (l, pure, nprim, origin, alpha, coeff) arrays of the C ABI (lb200_basis_create)."""
        l = np.array([s.l for s in self], dtype=np.int32)
        pure = np.array([int(s.pure) for s in self], dtype=np.int32)
        nprim = np.array([s.nprim for s in self], dtype=np.int32)
        O = np.array([s.O for s in self], dtype=np.float64).reshape(-1, 3)
        alpha = np.concatenate([s.alpha for s in self]) if len(self) else np.zeros(0)
        coeff = np.concatenate([s.coeff for s in self]) if len(self) else np.zeros(0)
        return l, pure, nprim, O, alpha, coeff


# ---------------------------------------------------------------------------------------
# synthetic geometries of BASELINE.json's configs
# ---------------------------------------------------------------------------------------
H2O_XYZ_ANGSTROM = [(8, (0.00000, -0.07579, 0.00000)), (1, (0.86681, 0.60144, 0.00000)),
                    (1, (-0.86681, 0.60144, 0.00000))]  # tests/hartree-fock/h2o.xyz


# tests/hartree-fock/h2o_rotated.xyz (h2o.xyz rotated by EulerMatrix[{pi/4,pi/6,pi/6}]), Angstrom:
# the default geometry of the reference's hartree-fock++ validation run
H2O_ROTATED_XYZ_ANGSTROM = [(8, (-0.06698952868266053, -0.02320585345069213, -0.026795811473064216)),
                            (1, (0.6848346853461241, -0.6120631876157682, 0.5191047657385741)),
                            (1, (0.37837107084596855, 0.9803684653406467, -0.09382246326173704))]
BOHR_TO_ANGSTROM_CODATA2010 = 0.52917721092  # atom.h:63; what tests/hartree-fock/hartree-fock.cc:306 uses


def _random_rotation(rng):
    q = rng.standard_normal(4)
    q /= np.linalg.norm(q)
    a, b, c, d = q
    return np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                     [2 * (b * c + a * d), a * a - b * b + c * c - d * d, 2 * (c * d - a * b)],
                     [2 * (b * d - a * c), 2 * (c * d + a * b), a * a - b * b - c * c + d * d]])


def water_cluster(nx, ny, nz, spacing_angstrom=3.10, seed=64):
    """(H2O)_{nx*ny*nz}: cubic lattice of h2o.xyz monomers, seeded random orientations."""
    rng = np.random.default_rng(seed)
    mono = np.array([r for _, r in H2O_XYZ_ANGSTROM])
    mono = mono - mono[0]
    atoms = []
    for ix in range(nx):
        for iy in range(ny):
            for iz in range(nz):
                R = _random_rotation(rng)
                c = np.array([ix, iy, iz], dtype=float) * spacing_angstrom
                for (Z, _), r in zip(H2O_XYZ_ANGSTROM, mono):
                    p = (R @ r + c) * ANGSTROM_TO_BOHR
                    atoms.append(Atom(Z, p[0], p[1], p[2]))
    return atoms


def alkane(ncarbon):
    """all-trans C_n H_{2n+2}: C-C 1.54 A, CCC 112 deg, C-H 1.09 A, HCH 109.5 deg."""
    cc, ch = 1.54, 1.09
    half = math.radians(112.0) / 2
    dx, dy = cc * math.sin(half), cc * math.cos(half)
    C = [np.array([i * dx, (i % 2) * dy, 0.0]) for i in range(ncarbon)]
    atoms = [(6, c) for c in C]
    hh = math.radians(109.5) / 2
    for i, c in enumerate(C):
        up = 1.0 if i % 2 else -1.0
        for sz in (+1.0, -1.0):
            atoms.append((1, c + ch * np.array([0.0, up * math.cos(hh), sz * math.sin(hh)])))
    # terminal hydrogens continue the zig-zag
    atoms.append((1, C[0] + ch * np.array([-math.sin(half), math.cos(half) * (1 if ncarbon > 1 else 1), 0.0])))
    sgn = -1.0 if (ncarbon - 1) % 2 else 1.0
    atoms.append((1, C[-1] + ch * np.array([math.sin(half), sgn * math.cos(half), 0.0])))
    return [Atom(Z, *(r * ANGSTROM_TO_BOHR)) for Z, r in atoms]
