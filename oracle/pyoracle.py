"""oracle/pyoracle.py -- TEST INFRASTRUCTURE (CPU oracle), not product code.

ctypes front-end to oracle/_ref/liboracle.so = the reference's unmodified
header-only libint2::Engine / Shell / ShellPair / BasisSet / FmEval_Chebyshev7 /
eri() (compiled from /root/reference where they lie, see oracle/Makefile) on top
of the restated CPU kernels of oracle/oracle_kernels.cc.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module; the product (libint_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "liboracle.so")

SCREEN_ORIGINAL = 0x0001
SCREEN_CONSERVATIVE = 0x0010
SCREEN_SCHWARZ = 0x0100
SCREEN_SCHWARZ_INF = 0x1000

FAST_LIB_PATH = os.path.join(HERE, "_ref", "liboracle_fast.so")
TRUTH_LIB_PATH = os.path.join(HERE, "_ref", "libtruth.so")

_libs = {}


def build(force=False):
    """Build oracle/_ref/*.so (needs /root/reference; prebuilt files are used otherwise)."""
    ref = os.environ.get("LIBINT_REFERENCE", "/root/reference")
    if os.path.exists(LIB_PATH) and not force and not os.path.isdir(ref):
        return LIB_PATH
    if not os.path.isdir(ref):
        raise RuntimeError("oracle not built and %s absent" % ref)
    subprocess.check_call(["make", "-C", HERE, "-j4", "REF=" + ref])
    return LIB_PATH


def _cpu_has(*flags):
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("flags"):
                have = set(ln.split(":", 1)[1].split())
                return all(f in have for f in flags)
    except OSError:
        pass
    return False


def fast_available():
    """The timing build (-O3 -march=x86-64-v3) needs AVX2 + FMA on the host it runs on."""
    return os.path.exists(FAST_LIB_PATH) and _cpu_has("avx2", "fma", "bmi2")


REFGPU_LIB_PATH = os.path.join(HERE, "_ref", "librefengine_b200.so")


def lib(fast=False, b200=False):
    """fast=False: the parity build (-O2 -ffp-contract=off, x86-64-v2).  fast=True: the same
    sources built -O3 -march=x86-64-v3 for the CPU-baseline timings (falls back to the parity
    build on a host without AVX2/FMA).  b200=True: NOT the oracle -- the same reference-Engine
    wrappers (oracle_capi.cc) compiled against the PRODUCT's generated headers (include/libint2) and
    linked to liblibint_b200_iface.so, i.e. the reference's unmodified libint2::Engine running on the
    GPU library's Libint_t / libint2_build_* boundary; the thing under test in tests/test_gpu_iface.py."""
    path = REFGPU_LIB_PATH if b200 else (FAST_LIB_PATH if (fast and fast_available()) else LIB_PATH)
    if path not in _libs:
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int)
        L.lbo_init.restype = C.c_int
        L.lbo_boys_cheb7.argtypes = [C.c_double, C.c_int, C.c_int, dp]
        L.lbo_boys_reference.argtypes = [C.c_double, C.c_int, dp]
        L.lbo_eri_closed.argtypes = [ip, dp, dp, C.c_int]
        L.lbo_eri_closed.restype = C.c_double
        L.lbo_shell_renorm.argtypes = [C.c_int, C.c_int, dp, dp, dp, dp]
        L.lbo_solidharmonic_coeff.argtypes = [C.c_int] * 5
        L.lbo_solidharmonic_coeff.restype = C.c_double
        L.lbo_compute2.argtypes = [C.c_int, ip, ip, ip, dp, dp, dp, C.c_int, C.c_int, C.c_double,
                                   C.c_int, dp, C.c_long]
        L.lbo_compute2.restype = C.c_long
        L.lbo_shellpair.argtypes = [ip, ip, ip, dp, dp, dp, C.c_int, C.c_double, C.c_int, dp,
                                    C.c_int, dp]
        L.lbo_fock_create.argtypes = [C.c_int, ip, ip, ip, dp, dp, dp, C.c_int, C.c_int, ip, ip,
                                      C.c_int]
        L.lbo_fock_create.restype = C.c_void_p
        L.lbo_fock_destroy.argtypes = [C.c_void_p]
        L.lbo_fock_nbf.argtypes = [C.c_void_p]
        L.lbo_fock_nbf.restype = C.c_long
        L.lbo_fock_schwarz.argtypes = [C.c_void_p, dp]
        L.lbo_fock_pairdata.argtypes = [C.c_void_p, C.c_int, C.c_int, dp, C.c_int]
        L.lbo_fock_build.argtypes = [C.c_void_p, dp, C.c_double, C.c_int, C.c_long, C.c_long, dp,
                                     dp]
        L.lbo_time_quartets.argtypes = [C.c_int, ip, ip, ip, dp, dp, dp, C.c_int, C.c_long, ip,
                                        C.c_int, C.c_int, dp]
        L.lbo_time_quartets.restype = C.c_double
        L.lbo_time_triplets.argtypes = [C.c_int, ip, ip, ip, dp, dp, dp, C.c_int, ip, ip, ip, dp, dp, dp,
                                        C.c_long, ip, C.c_int, dp]
        L.lbo_time_triplets.restype = C.c_double
        L.lbo_compute_batch.argtypes = [C.c_int, ip, ip, ip, dp, dp, dp, C.c_int, C.c_long, ip,
                                        C.c_int, C.c_double, dp]
        L.lbo_compute_batch.restype = C.c_long
        L.lbo_basis_load.argtypes = [C.c_char_p, C.c_int, ip, dp, C.c_int, C.c_int, ip, ip, ip, dp,
                                     dp, dp, ip]
        if hasattr(L, "lbo_compute2_deriv1"):
            L.lbo_compute2_deriv1.argtypes = [ip, ip, ip, dp, dp, dp, C.c_int, C.c_double, dp, C.c_long]
            L.lbo_compute2_deriv1.restype = C.c_long
        if hasattr(L, "lbo_deriv1_closed"):
            L.lbo_deriv1_closed.argtypes = [ip, ip, dp, dp, dp, C.c_int, dp, C.c_long]
            L.lbo_deriv1_closed.restype = C.c_long
            L.lbo_fock_grad_closed.argtypes = [C.c_int, ip, ip, dp, dp, dp, C.c_int, dp, C.c_int, ip, C.c_int, dp]
        L.lbo_init()
        _libs[path] = L
    return _libs[path]


def truth_lib():
    """Extended-precision arbiter (oracle/truth.cc); own code only, so it can be (re)built on
    any host with g++."""
    if TRUTH_LIB_PATH not in _libs:
        if not os.path.exists(TRUTH_LIB_PATH):
            subprocess.check_call(["make", "-C", HERE, "_ref/libtruth.so"])
        L = C.CDLL(TRUTH_LIB_PATH)
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int)
        L.lbt_eri_batch.argtypes = [C.c_int, C.c_int, ip, ip, dp, dp, dp, C.c_long, ip, C.c_int, dp, dp]
        L.lbt_boys.argtypes = [C.c_int, C.c_double, C.c_int, dp]
        _libs[TRUTH_LIB_PATH] = L
    return _libs[TRUTH_LIB_PATH]


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class Shells:
    """Flat shell table: l, pure, nprim (int32), O (n,3), alpha/coeff concatenated."""

    def __init__(self, l, pure, nprim, O, alpha, coeff, raw=True):
        self.l = np.ascontiguousarray(l, dtype=np.int32)
        self.pure = np.ascontiguousarray(pure, dtype=np.int32)
        self.nprim = np.ascontiguousarray(nprim, dtype=np.int32)
        self.O = np.ascontiguousarray(O, dtype=np.float64).reshape(-1, 3)
        self.alpha = np.ascontiguousarray(alpha, dtype=np.float64)
        self.coeff = np.ascontiguousarray(coeff, dtype=np.float64)
        self.raw = bool(raw)
        assert self.alpha.size == self.nprim.sum() == self.coeff.size

    def __len__(self):
        return len(self.l)

    def offsets(self):
        return np.concatenate([[0], np.cumsum(self.nprim)]).astype(np.int64)

    def subset(self, idx):
        off = self.offsets()
        al = np.concatenate([self.alpha[off[i]:off[i + 1]] for i in idx])
        co = np.concatenate([self.coeff[off[i]:off[i + 1]] for i in idx])
        idx = np.asarray(idx)
        return Shells(self.l[idx], self.pure[idx], self.nprim[idx], self.O[idx], al, co, self.raw)

    def size(self, i):
        l = int(self.l[i])
        return 2 * l + 1 if self.pure[i] else (l + 1) * (l + 2) // 2

    def args(self):
        return (_i(self.l), _i(self.pure), _i(self.nprim), _d(self.O), _d(self.alpha),
                _d(self.coeff), int(self.raw))


def set_unit_normalization(flag):
    lib().lbo_set_unit_normalization(int(flag))


def boys_cheb7(T, mmax, table_mmax=None):
    out = np.zeros(mmax + 1)
    lib().lbo_boys_cheb7(float(T), int(mmax), int(table_mmax if table_mmax is not None else mmax),
                         _d(out))
    return out


def boys_reference(T, mmax):
    out = np.zeros(mmax + 1)
    lib().lbo_boys_reference(float(T), int(mmax), _d(out))
    return out


def eri_closed(lmn, alpha, centers, norm_flag=0):
    lmn = np.ascontiguousarray(lmn, dtype=np.int32).reshape(12)
    alpha = np.ascontiguousarray(alpha, dtype=np.float64).reshape(4)
    centers = np.ascontiguousarray(centers, dtype=np.float64).reshape(12)
    return lib().lbo_eri_closed(_i(lmn), _d(alpha), _d(centers), int(norm_flag))


def shell_renorm(l, alpha, coeff):
    alpha = np.ascontiguousarray(alpha, dtype=np.float64)
    coeff = np.ascontiguousarray(coeff, dtype=np.float64)
    oc = np.zeros_like(alpha)
    om = np.zeros_like(alpha)
    lib().lbo_shell_renorm(int(l), len(alpha), _d(alpha), _d(coeff), _d(oc), _d(om))
    return oc, om


def solidharmonic_coeff(l, m, lx, ly, lz):
    return lib().lbo_solidharmonic_coeff(l, m, lx, ly, lz)


def compute2(shells, braket=0, screening=SCREEN_ORIGINAL, precision=np.finfo(float).eps,
             uniform_cart_norm=False, b200=False):
    """One shell set through the reference Engine. Returns None if screened out.
    b200=True: the same Engine on the GPU library's Libint_t boundary (see lib())."""
    n = 1
    for i in range(len(shells)):
        n *= shells.size(i)
    out = np.zeros(n)
    r = lib(b200=b200).lbo_compute2(int(braket), *shells.args(), int(screening), float(precision),
                           int(uniform_cart_norm), _d(out), n)
    if r < 0:
        raise RuntimeError("lbo_compute2 failed (%d)" % r)
    if r == 0:
        return None
    return out.reshape([shells.size(i) for i in range(len(shells))])


def shellpair(shells2, ln_prec, screening=SCREEN_ORIGINAL):
    cap = int(shells2.nprim[0] * shells2.nprim[1])
    out = np.zeros((cap, 9))
    AB = np.zeros(3)
    n = lib().lbo_shellpair(*shells2.args(), float(ln_prec), int(screening), _d(out), cap, _d(AB))
    if n < 0:
        raise RuntimeError("lbo_shellpair failed")
    return out[:n].copy(), AB


class Fock:
    """Direct Fock build through the reference Engine (hartree-fock++.cc pattern)."""

    def __init__(self, shells, pair_s1, pair_s2, nthreads=1, fast=False, b200=False):
        self.shells = shells
        self.L = lib(fast, b200)   # fast=True: the -O3 timing build (CPU-baseline legs only)
        p1 = np.ascontiguousarray(pair_s1, dtype=np.int32)
        p2 = np.ascontiguousarray(pair_s2, dtype=np.int32)
        self.h = self.L.lbo_fock_create(len(shells), *shells.args(), len(p1), _i(p1), _i(p2),
                                       int(nthreads))
        self.nbf = self.L.lbo_fock_nbf(self.h)
        self.nshell = len(shells)

    def schwarz(self):
        K = np.zeros((self.nshell, self.nshell))
        self.L.lbo_fock_schwarz(self.h, _d(K))
        return K

    def pairdata(self, s1, s2):
        cap = int(self.shells.nprim[s1] * self.shells.nprim[s2])
        out = np.zeros((cap, 9))
        n = self.L.lbo_fock_pairdata(self.h, int(s1), int(s2), _d(out), cap)
        if n < 0:
            raise KeyError((s1, s2))
        return out[:n].copy()

    def build(self, D, precision, use_schwarz=True, task_stride=1, task_offset=0):
        D = np.ascontiguousarray(D, dtype=np.float64)
        G = np.zeros((self.nbf, self.nbf))
        stats = np.zeros(3)
        self.L.lbo_fock_build(self.h, _d(D), float(precision), int(use_schwarz), int(task_stride),
                             int(task_offset), _d(G), _d(stats))
        return G, {"nints": stats[0], "nquartets": stats[1], "seconds": stats[2]}

    def close(self):
        if self.h:
            self.L.lbo_fock_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def time_quartets(shells, quartets, nthreads=1, use_pairs=True, fast=True):
    """wall seconds of the reference Engine over the quartet list (one Engine per thread).
    use_pairs: precomputed ShellPairs handed to compute2 (hartree-fock++.cc:1697); fast: the
    -O3 -march=x86-64-v3 build of the same sources."""
    q = np.ascontiguousarray(quartets, dtype=np.int32).reshape(-1, 4)
    s = C.c_double(0)
    t = lib(fast).lbo_time_quartets(len(shells), *shells.args(), len(q), _i(q), int(nthreads),
                                    int(bool(use_pairs)), C.byref(s))
    return t, s.value


def time_triplets(dfshells, obsshells, triplets, nthreads=1, fast=True):
    """wall seconds of Engine(xs_xx).compute2(df[P], unit, obs[a], obs[b]) over the triplet list."""
    t = np.ascontiguousarray(triplets, dtype=np.int32).reshape(-1, 3)
    s = C.c_double(0)
    a1, a2 = dfshells.args()[:-1], obsshells.args()[:-1]
    sec = lib(fast).lbo_time_triplets(len(dfshells), *a1, len(obsshells), *a2, len(t), _i(t), int(nthreads),
                                      C.byref(s))
    return sec, s.value


def compute_batch(shells, quartets, nthreads=1, precision=0.0):
    """Reference Engine (parity build) over a list of quartets of one class -> (n, blk)."""
    q = np.ascontiguousarray(quartets, dtype=np.int32).reshape(-1, 4)
    if len(q) == 0:
        return np.zeros((0, 0))
    blk = 1
    for i in q[0]:
        blk *= shells.size(int(i))
    out = np.empty((len(q), blk))
    r = lib().lbo_compute_batch(len(shells), *shells.args(), len(q), _i(q), int(nthreads),
                                float(precision), _d(out))
    if r != blk:
        raise RuntimeError("lbo_compute_batch failed (%d)" % r)
    return out


def compute2_deriv1(shells4, precision=0.0, b200=True):
    """The reference's unmodified Engine with deriv_order = 1 on one quartet -> (12, n1*n2*n3*n4), or None when
    screened out.  b200=True (the only build whose headers have derivative order 1): the Engine runs on the GPU
    library's Libint_t boundary (libint2_build_eri1 of liblibint_b200_iface.so)."""
    n = 1
    for i in range(4):
        n *= shells4.size(i)
    out = np.zeros((12, n))
    r = lib(b200=b200).lbo_compute2_deriv1(*shells4.args(), float(precision), _d(out), out.size)
    if r == -3:
        raise RuntimeError("this oracle build has LIBINT2_MAX_DERIV_ORDER 0")
    if r < 0:
        raise RuntimeError("lbo_compute2_deriv1 failed (%d)" % r)
    return None if r == 0 else out


def deriv1_closed(shells4):
    """The twelve Cartesian derivative shell sets of one quartet from the reference's closed-form eri()
    with a derivative index (tests/eri/test.cc:381-445) -> (12, n1*n2*n3*n4)."""
    blk = 1
    for i in range(4):
        l = int(shells4.l[i])
        blk *= (l + 1) * (l + 2) // 2
    out = np.zeros((12, blk))
    r = lib().lbo_deriv1_closed(_i(shells4.l), _i(shells4.nprim), _d(shells4.O), _d(shells4.alpha),
                                _d(shells4.coeff), int(shells4.raw), _d(out), out.size)
    if r != blk:
        raise RuntimeError("lbo_deriv1_closed failed (%d)" % r)
    return out


def fock_grad_closed(shells, D, shell2atom, natoms, nthreads=1):
    """F2[natoms, 3] as hartree-fock++.cc:648-656 forms it from compute_2body_fock_deriv<1>, with the
    closed-form derivative sets; Cartesian shells only, no screening."""
    assert not shells.pure.any(), "Cartesian shells only: back-transform the density of pure shells first"
    D = np.ascontiguousarray(D, dtype=np.float64)
    s2a = np.ascontiguousarray(shell2atom, dtype=np.int32)
    g = np.zeros((natoms, 3))
    r = lib().lbo_fock_grad_closed(len(shells), _i(shells.l), _i(shells.nprim), _d(shells.O), _d(shells.alpha),
                                   _d(shells.coeff), int(shells.raw), _d(D), int(natoms), _i(s2a), int(nthreads),
                                   _d(g))
    if r != 0:
        raise RuntimeError("lbo_fock_grad_closed failed (%d)" % r)
    return g


def truth_batch(shells, quartets, nthreads=1, quad=False, with_lo=True):
    """Extended-precision Cartesian shell sets (long double, or __float128 with quad=True) for a
    list of quartets of one class: returns (hi, lo) with truth = hi + lo (lo None if not asked)."""
    assert not shells.raw, "the arbiter takes normalization-embedded coefficients"
    q = np.ascontiguousarray(quartets, dtype=np.int32).reshape(-1, 4)
    blk = 1
    for i in q[0]:
        l = int(shells.l[int(i)])
        blk *= (l + 1) * (l + 2) // 2
    hi = np.empty((len(q), blk))
    lo = np.empty((len(q), blk)) if with_lo else None
    truth_lib().lbt_eri_batch(int(bool(quad)), len(shells), _i(shells.l), _i(shells.nprim), _d(shells.O),
                              _d(shells.alpha), _d(shells.coeff), len(q), _i(q), int(nthreads),
                              _d(hi), _d(lo) if with_lo else None)
    return hi, lo


def truth_boys(T, mmax, quad=True):
    out = np.zeros(mmax + 1)
    truth_lib().lbt_boys(int(bool(quad)), float(T), int(mmax), _d(out))
    return out


def basis_load(name, Z, xyz_bohr, data_path):
    """Reference BasisSet(name, atoms); data_path must contain basis/<name>.g94."""
    os.environ["LIBINT_DATA_PATH"] = data_path
    Z = np.ascontiguousarray(Z, dtype=np.int32)
    xyz = np.ascontiguousarray(xyz_bohr, dtype=np.float64)
    npt = C.c_int(0)
    z = np.zeros(1, dtype=np.int32)
    zd = np.zeros(3)
    ns = lib().lbo_basis_load(name.encode(), len(Z), _i(Z), _d(xyz), 0, 0, _i(z), _i(z), _i(z),
                              _d(zd), _d(zd), _d(zd), C.byref(npt))
    if ns < 0:
        raise RuntimeError("basis load failed")
    l = np.zeros(ns, dtype=np.int32)
    pure = np.zeros(ns, dtype=np.int32)
    nprim = np.zeros(ns, dtype=np.int32)
    O = np.zeros((ns, 3))
    alpha = np.zeros(npt.value)
    coeff = np.zeros(npt.value)
    lib().lbo_basis_load(name.encode(), len(Z), _i(Z), _d(xyz), ns, npt.value, _i(l), _i(pure),
                         _i(nprim), _d(O), _d(alpha), _d(coeff), C.byref(npt))
    return Shells(l, pure, nprim, O, alpha, coeff, raw=False)


# ---------------------------------------------------------------------------------------
# parity against the arbiter
# ---------------------------------------------------------------------------------------
RTOL, ATOL = 1e-12, 1e-14   # BASELINE.json north_star


def truth_errors(x, hi, lo):
    """|x - truth| with truth = hi + lo (extended precision split into two doubles)."""
    return np.abs((np.asarray(x, dtype=np.float64) - hi) - lo)


def parity_stats(got, orc, hi, lo):
    """GPU result `got` and reference-Engine result `orc` of the same shell sets (n, blk) against the
    extended-precision truth: the numbers the parity criterion is stated on.
      *_outside : elements outside the literal 1e-12 rel / 1e-14 abs tolerance vs the truth
      *_max_abs : max |x - truth|;  *_rms : root mean square error
      *_max_scaled : max |x - truth| / (1e-14 + 1e-12 |truth|)  (<= 1 <=> literal tolerance met)
      worse_sets : shell sets where max|got - truth| > max(literal, max|orc - truth|) -- sets on
                   which the GPU is further from the truth than both the tolerance and the
                   reference itself."""
    eg, eo = truth_errors(got, hi, lo), truth_errors(orc, hi, lo)
    tol = ATOL + RTOL * np.abs(hi)
    gq, oq = eg.max(axis=1), eo.max(axis=1)
    lit = (eg <= tol).all(axis=1)
    return {
        "integrals": int(eg.size), "shell_sets": int(eg.shape[0]),
        "gpu_outside": int((eg > tol).sum()), "oracle_outside": int((eo > tol).sum()),
        "gpu_max_abs": float(eg.max()), "oracle_max_abs": float(eo.max()),
        "gpu_rms": float(np.sqrt(np.mean(eg * eg))), "oracle_rms": float(np.sqrt(np.mean(eo * eo))),
        "gpu_max_scaled": float((eg / tol).max()), "oracle_max_scaled": float((eo / tol).max()),
        "gpu_vs_oracle_max_abs": float(np.abs(np.asarray(got) - np.asarray(orc)).max()),
        "worse_sets": int((~lit & (gq > oq)).sum()),
        # over the shell sets where the GPU misses the literal tolerance: worst ratio of its max error
        # to the reference's own max error on the same set
        "max_ratio_nonliteral": float((gq[~lit] / np.maximum(oq[~lit], 1e-300)).max()) if (~lit).any() else 0.0,
        "nonliteral_sets": int((~lit).sum()),
        "sets_gpu_closer": int((gq < oq).sum()), "sets_oracle_closer": int((oq < gq).sum()),
    }
