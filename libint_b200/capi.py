"""ctypes binding of the C ABI declared in include/libint_b200.h.

This is the only way Python reaches the CUDA path; there is no CPU fallback: `load()`
raises if the shared library is missing, and every compute call raises `Lb200Error` when
the library reports a failure (no GPU, unsupported class, CUDA error).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_lib", "liblibint_b200%s.so" % os.environ.get("LB200_LIB_SUFFIX", ""))

OK = 0
SCREEN_ORIGINAL = 0x0001
SCREEN_CONSERVATIVE = 0x0010
SCREEN_SCHWARZ = 0x0100
SCREEN_SCHWARZ_INF = 0x1000
MAX_AM = 4


class Lb200Error(RuntimeError):
    pass


_lib = None
dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)
vp = C.c_void_p

# name -> (restype, argtypes); also the list tests check against include/libint_b200.h
SIGNATURES = {
    "lb200_version": (C.c_int, []),
    "lb200_device_count": (C.c_int, []),
    "lb200_context_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
    "lb200_context_destroy": (C.c_int, [vp]),
    "lb200_last_error": (C.c_char_p, [vp]),
    "lb200_context_set_stream": (C.c_int, [vp, vp]),
    "lb200_context_synchronize": (C.c_int, [vp]),
    "lb200_context_launch_count": (C.c_longlong, [vp]),
    "lb200_shell_renorm": (C.c_int, [C.c_int, C.c_int, dp, dp, C.c_int, dp]),
    "lb200_basis_create": (C.c_int, [vp, C.c_int, ip, ip, ip, dp, dp, dp, C.POINTER(vp)]),
    "lb200_basis_create_unit": (C.c_int, [vp, C.POINTER(vp)]),
    "lb200_basis_destroy": (C.c_int, [vp]),
    "lb200_basis_nbf": (C.c_int, [vp]),
    "lb200_basis_nshell": (C.c_int, [vp]),
    "lb200_basis_shell2bf": (C.c_int, [vp, ip]),
    "lb200_pairs_create": (C.c_int, [vp, vp, vp, C.c_int, ip, ip, C.c_int, C.c_double, dp,
                                     C.POINTER(vp)]),
    "lb200_pairs_destroy": (C.c_int, [vp]),
    "lb200_pairs_info": (C.c_int, [vp, C.POINTER(C.c_longlong)]),
    "lb200_pairs_get": (C.c_int, [vp, C.c_int, dp, C.c_int]),
    "lb200_eri_batch": (C.c_int, [vp, vp, vp, C.c_longlong, vp, C.c_int, C.c_int, C.c_double,
                                  C.c_int, vp, C.c_int]),
    "lb200_eri_deriv1_batch": (C.c_int, [vp, vp, vp, C.c_longlong, vp, C.c_int, C.c_int, C.c_double,
                                         C.c_int, vp, C.c_int]),
    "lb200_eri_deriv1_plan": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong)]),
    "lb200_fock_grad": (C.c_int, [vp, vp, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, ip, dp, dp]),
    "lb200_eri_prereq_batch": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_longlong, ip, dp, dp, dp]),
    "lb200_eri_block_size": (C.c_longlong, [vp, vp, C.c_int]),
    "lb200_eri_class_supported": (C.c_int, [C.c_int] * 4),
    "lb200_significant_pairs": (C.c_int, [vp, C.c_double, ip, ip, C.c_longlong,
                                          C.POINTER(C.c_longlong)]),
    "lb200_significant_pairs_device": (C.c_int, [vp, vp, C.c_double, ip, ip, C.c_longlong,
                                                 C.POINTER(C.c_longlong)]),
    "lb200_fock_create": (C.c_int, [vp, vp, C.c_longlong, ip, ip, C.POINTER(vp)]),
    "lb200_fock_destroy": (C.c_int, [vp]),
    "lb200_fock_schwarz": (C.c_int, [vp, dp]),
    "lb200_onebody": (C.c_int, [vp, vp, C.c_int, dp, vp, vp, vp, C.c_int]),
    "lb200_onebody_forces": (C.c_int, [vp, vp, C.c_int, dp, ip, vp, vp, C.c_int, dp, dp]),
    "lb200_eri_product": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, vp]),
    "lb200_df3c_create": (C.c_int, [vp, vp, vp, C.c_longlong, ip, ip, C.POINTER(vp)]),
    "lb200_df3c_destroy": (C.c_int, [vp]),
    "lb200_df3c_slab": (C.c_int, [vp, C.c_int, C.c_int, C.c_double, C.c_double, vp, dp]),
    "lb200_df3c_metric": (C.c_int, [vp, vp]),
    "lb200_df3c_info": (C.c_int, [vp, C.POINTER(C.c_longlong)]),
    "lb200_comm_unique_id": (C.c_int, [C.c_char_p, C.c_int]),
    "lb200_comm_create": (C.c_int, [vp, C.c_int, C.c_int, C.c_char_p, C.POINTER(vp)]),
    "lb200_comm_from_nccl": (C.c_int, [vp, vp, C.POINTER(vp)]),
    "lb200_comm_destroy": (C.c_int, [vp]),
    "lb200_comm_rank": (C.c_int, [vp]),
    "lb200_comm_size": (C.c_int, [vp]),
    "lb200_fock_allreduce": (C.c_int, [vp, vp, C.c_longlong]),
    "lb200_fock_set_profile": (C.c_int, [vp, C.c_int]),
    "lb200_fock_get_profile": (C.c_longlong, [vp, dp, C.c_longlong]),
    "lb200_fock_task_owner": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "lb200_fp64_peak_probe": (C.c_int, [vp, C.c_int, dp, dp]),
    "lb200_fock_build": (C.c_int, [vp, vp, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, vp,
                                   C.c_int, dp]),
}


def load():
    """Load the CUDA extension; fails loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Lb200Error(
                "libint_b200 CUDA extension not built: %s missing "
                "(run `python -m libint_b200.build`); there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(dp)


def _i(a):
    return a.ctypes.data_as(ip)


def _ptr(x):
    """numpy array -> (pointer, on_device=0); torch CUDA tensor -> (pointer, 1)."""
    if isinstance(x, np.ndarray):
        return C.c_void_p(x.ctypes.data), 0
    if hasattr(x, "data_ptr"):
        return C.c_void_p(x.data_ptr()), 1 if x.is_cuda else 0
    raise TypeError("expected numpy array or torch tensor")


class Context:
    def __init__(self, device=0):
        L = load()
        h = vp()
        rc = L.lb200_context_create(int(device), C.byref(h))
        if rc != OK:
            raise Lb200Error("lb200_context_create(device=%d) failed with %d: no usable CUDA "
                             "device (this library has no CPU fallback)" % (device, rc))
        self.h = h
        self.device = device

    def check(self, rc, what=""):
        if rc != OK:
            msg = load().lb200_last_error(self.h)
            raise Lb200Error("%s failed (%d): %s" % (what, rc, msg.decode() if msg else ""))

    def set_stream(self, cuda_stream_ptr):
        self.check(load().lb200_context_set_stream(self.h, vp(cuda_stream_ptr)), "set_stream")

    def synchronize(self):
        self.check(load().lb200_context_synchronize(self.h), "synchronize")

    @property
    def launch_count(self):
        return load().lb200_context_launch_count(self.h)

    def close(self):
        if self.h:
            load().lb200_context_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def shell_renorm(l, alpha, coeff, enforce_unit_normalization=True):
    alpha = np.ascontiguousarray(alpha, dtype=np.float64)
    c = np.array(coeff, dtype=np.float64)
    m = np.zeros_like(c)
    rc = load().lb200_shell_renorm(int(l), len(alpha), _d(alpha), _d(c),
                                   int(enforce_unit_normalization), _d(m))
    if rc != OK:
        raise Lb200Error("lb200_shell_renorm failed (%d)" % rc)
    return c, m


class Basis:
    """Flat shell table on the library side (coefficients carry the normalization)."""

    def __init__(self, ctx, l, pure, nprim, origin, alpha, coeff, unit=False):
        self.ctx = ctx
        h = vp()
        if unit:
            ctx.check(load().lb200_basis_create_unit(ctx.h, C.byref(h)), "basis_create_unit")
            l, pure, nprim, origin, alpha, coeff = [0], [0], [1], [[0, 0, 0]], [0.0], [1.0]
        self.l = np.ascontiguousarray(l, dtype=np.int32)
        self.pure = np.ascontiguousarray(pure, dtype=np.int32)
        self.nprim = np.ascontiguousarray(nprim, dtype=np.int32)
        self.origin = np.ascontiguousarray(origin, dtype=np.float64).reshape(-1, 3)
        self.alpha = np.ascontiguousarray(alpha, dtype=np.float64)
        self.coeff = np.ascontiguousarray(coeff, dtype=np.float64)
        if not unit:
            ctx.check(load().lb200_basis_create(ctx.h, len(self.l), _i(self.l), _i(self.pure),
                                                _i(self.nprim), _d(self.origin), _d(self.alpha),
                                                _d(self.coeff), C.byref(h)), "basis_create")
        self.h = h
        self.nshell = len(self.l)
        self.nbf = load().lb200_basis_nbf(h)
        s2b = np.zeros(self.nshell, dtype=np.int32)
        load().lb200_basis_shell2bf(h, _i(s2b))
        self.shell2bf = s2b

    @classmethod
    def unit(cls, ctx):
        return cls(ctx, None, None, None, None, None, None, unit=True)

    def size(self, s):
        l = int(self.l[s])
        return 2 * l + 1 if self.pure[s] else (l + 1) * (l + 2) // 2

    def close(self):
        if self.h:
            load().lb200_basis_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Pairs:
    """A block of shell pairs of one class, resident on the GPU."""

    def __init__(self, ctx, bs1, bs2, s1, s2, screening=SCREEN_ORIGINAL, ln_prec=-np.inf,
                 prim_schwarz=None):
        self.ctx, self.bs1, self.bs2 = ctx, bs1, bs2
        self.s1 = np.ascontiguousarray(s1, dtype=np.int32)
        self.s2 = np.ascontiguousarray(s2, dtype=np.int32)
        if not np.isfinite(ln_prec):
            ln_prec = -np.finfo(np.float64).max
        ps = None if prim_schwarz is None else np.ascontiguousarray(prim_schwarz, dtype=np.float64)
        h = vp()
        ctx.check(load().lb200_pairs_create(ctx.h, bs1.h, bs2.h, len(self.s1), _i(self.s1),
                                            _i(self.s2), int(screening), float(ln_prec),
                                            _d(ps) if ps is not None else None, C.byref(h)),
                  "pairs_create")
        self.h = h
        info = (C.c_longlong * 6)()
        load().lb200_pairs_info(h, info)
        self.la, self.lb, self.npair, self.nprimpair, self.pure_a, self.pure_b = [int(x) for x in info]

    def get(self, i):
        cap = int(self.bs1.nprim[self.s1[i]] * self.bs2.nprim[self.s2[i]])
        out = np.zeros((cap, 9))
        n = load().lb200_pairs_get(self.h, int(i), _d(out), cap)
        if n < 0:
            raise Lb200Error("pairs_get failed")
        return out[:n].copy()

    def close(self):
        if self.h:
            load().lb200_pairs_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def eri_block_size(bra, ket, pure_out=False):
    return int(load().lb200_eri_block_size(bra.h, ket.h, int(pure_out)))


def eri_batch(ctx, bra, ket, tasks, out=None, screening=SCREEN_ORIGINAL, precision=0.0,
              pure_out=False):
    """Batched Engine::compute2. tasks: (n,2) int32 numpy array or torch int32 CUDA tensor;
    out: numpy array / torch CUDA tensor of n*block doubles (allocated as numpy if None)."""
    blk = eri_block_size(bra, ket, pure_out)
    if isinstance(tasks, np.ndarray):
        tasks = np.ascontiguousarray(tasks, dtype=np.int32).reshape(-1, 2)
        n = tasks.shape[0]
    else:
        n = tasks.shape[0]
    if out is None:
        out = np.empty((n, blk))
    tp, tdev = _ptr(tasks)
    op, odev = _ptr(out)
    ctx.check(load().lb200_eri_batch(ctx.h, bra.h, ket.h, n, tp, tdev, int(screening),
                                     float(precision), int(pure_out), op, odev), "eri_batch")
    return out


def eri_deriv1_plan(la, lb, lc, ld):
    """(6, 5) int64: per shifted set A+, A-, B+, B-, C+, C- the doubles per task and the strides of the
    (a, b, c, d) component indices (lb200_eri_deriv1_plan)."""
    plan = (C.c_longlong * 30)()
    rc = load().lb200_eri_deriv1_plan(int(la), int(lb), int(lc), int(ld), plan)
    if rc != OK:
        raise Lb200Error("lb200_eri_deriv1_plan(%d%d|%d%d) failed (%d)" % (la, lb, lc, ld, rc))
    return np.array(list(plan), dtype=np.int64).reshape(6, 5)


def eri_deriv1_batch(ctx, bra, ket, tasks, out=None, screening=SCREEN_ORIGINAL, precision=0.0,
                     pure_out=False):
    """Batched Engine::compute2<coulomb, xx_xx, 1>: out[n, 12, block]; block d = 3 * centre + xyz, centres
    in the order (bra.first, bra.second, ket.first, ket.second)."""
    blk = eri_block_size(bra, ket, pure_out)
    if isinstance(tasks, np.ndarray):
        tasks = np.ascontiguousarray(tasks, dtype=np.int32).reshape(-1, 2)
    n = tasks.shape[0]
    if out is None:
        out = np.empty((n, 12, blk))
    tp, tdev = _ptr(tasks)
    op, odev = _ptr(out)
    ctx.check(load().lb200_eri_deriv1_batch(ctx.h, bra.h, ket.h, n, tp, tdev, int(screening),
                                            float(precision), int(pure_out), op, odev), "eri_deriv1_batch")
    return out


def significant_pairs(bs, threshold=1e-12, device=True):
    """obs_shellpair_list of the reference (hartree-fock++.cc:1305-1381) as two int32 arrays (s1 >= s2).
    device=True: evaluated by a GPU kernel on the basis' context; False: the serial host loop."""
    if device and getattr(bs, "ctx", None) is not None:
        cap = bs.nshell * (bs.nshell + 1) // 2
        s1 = np.zeros(cap, dtype=np.int32)
        s2 = np.zeros(cap, dtype=np.int32)
        cnt = C.c_longlong(0)
        bs.ctx.check(load().lb200_significant_pairs_device(bs.ctx.h, bs.h, float(threshold), _i(s1), _i(s2), cap,
                                                           C.byref(cnt)), "significant_pairs_device")
        return s1[:cnt.value].copy(), s2[:cnt.value].copy()
    cnt = C.c_longlong(0)
    load().lb200_significant_pairs(bs.h, float(threshold), None, None, 0, C.byref(cnt))
    s1 = np.zeros(cnt.value, dtype=np.int32)
    s2 = np.zeros(cnt.value, dtype=np.int32)
    rc = load().lb200_significant_pairs(bs.h, float(threshold), _i(s1), _i(s2), cnt.value,
                                        C.byref(cnt))
    if rc != OK:
        raise Lb200Error("significant_pairs failed (%d)" % rc)
    return s1, s2


class Fock:
    """Direct Fock builder (compute_2body_fock of the reference's hartree-fock++ driver)."""

    def __init__(self, ctx, obs, pair_s1=None, pair_s2=None, threshold=1e-12):
        self.ctx, self.obs = ctx, obs
        if pair_s1 is None:
            pair_s1, pair_s2 = significant_pairs(obs, threshold)
        self.pair_s1 = np.ascontiguousarray(pair_s1, dtype=np.int32)
        self.pair_s2 = np.ascontiguousarray(pair_s2, dtype=np.int32)
        h = vp()
        ctx.check(load().lb200_fock_create(ctx.h, obs.h, len(self.pair_s1), _i(self.pair_s1),
                                           _i(self.pair_s2), C.byref(h)), "fock_create")
        self.h = h

    def schwarz(self):
        K = np.zeros((self.obs.nshell, self.obs.nshell))
        load().lb200_fock_schwarz(self.h, _d(K))
        return K

    def build(self, D, precision, use_schwarz=True, rank=0, nranks=1, out=None, stats=False):
        n = self.obs.nbf
        if out is None:
            out = np.empty((n, n))
        if isinstance(D, np.ndarray):
            D = np.ascontiguousarray(D, dtype=np.float64)
        Dp, Ddev = _ptr(D)
        Gp, Gdev = _ptr(out)
        st = np.zeros(4)
        self.ctx.check(load().lb200_fock_build(self.h, Dp, Ddev, float(precision), int(use_schwarz),
                                               int(rank), int(nranks), Gp, Gdev,
                                               _d(st) if stats else None), "fock_build")
        if stats:
            return out, {"nquartets": st[0], "launches": st[1], "ms": st[2], "candidates": st[3]}
        return out

    def gradient(self, D, shell2atom, natoms, precision, use_schwarz=True, rank=0, nranks=1, stats=False):
        """two-body forces F2[natoms, 3] = sum_ij G1[3 atom + xyz]_ij D_ij (compute_2body_fock_deriv<1> of the
        reference contracted with D, hartree-fock++.cc:648-656) -- the partial sum of this rank's quartets"""
        if isinstance(D, np.ndarray):
            D = np.ascontiguousarray(D, dtype=np.float64)
        Dp, Ddev = _ptr(D)
        s2a = np.ascontiguousarray(shell2atom, dtype=np.int32)
        if len(s2a) != self.obs.nshell:
            raise ValueError("shell2atom needs one entry per shell")
        g = np.zeros((int(natoms), 3))
        st = np.zeros(3)
        self.ctx.check(load().lb200_fock_grad(self.h, Dp, Ddev, float(precision), int(use_schwarz), int(rank),
                                              int(nranks), int(natoms), _i(s2a), _d(g), _d(st)), "fock_grad")
        if stats:
            return g, {"nquartets": st[0], "launches": st[1], "ms": st[2]}
        return g

    def set_profile(self, on=True):
        load().lb200_fock_set_profile(self.h, int(bool(on)))

    def profile(self):
        """rows of the last profiled build: (la, lb, lc, ld, bra bucket, ket bucket, ms, quartets,
        surviving primitive quartets), slowest first."""
        n = load().lb200_fock_get_profile(self.h, None, 0)
        rows = np.zeros((max(n, 1), 9))
        load().lb200_fock_get_profile(self.h, _d(rows), n)
        return rows[:n]

    def close(self):
        if self.h:
            load().lb200_fock_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def onebody(ctx, basis, charges, device=False):
    """S, T, V of `basis` with point charges [(Z, (x, y, z)), ...] on the GPU (lb200_onebody).
    device=False: numpy arrays; True: torch CUDA tensors."""
    ch = np.ascontiguousarray([[z, c[0], c[1], c[2]] for z, c in charges], dtype=np.float64).reshape(-1, 4)
    n = basis.nbf
    if device:
        import torch
        dev = torch.device("cuda", ctx.device)
        M = [torch.empty((n, n), dtype=torch.float64, device=dev) for _ in range(3)]
        ctx.check(load().lb200_onebody(ctx.h, basis.h, len(ch), _d(ch), vp(M[0].data_ptr()), vp(M[1].data_ptr()),
                                       vp(M[2].data_ptr()), 1), "onebody")
        return M
    M = [np.empty((n, n)) for _ in range(3)]
    ctx.check(load().lb200_onebody(ctx.h, basis.h, len(ch), _d(ch), vp(M[0].ctypes.data), vp(M[1].ctypes.data),
                                   vp(M[2].ctypes.data), 0), "onebody")
    return M


def onebody_forces(ctx, basis, charges, shell2atom, D, W):
    """(F1, F_Pulay), each [natom, 3]: 2 sum (T1 + V1) o D and -2 sum S1 o W of hartree-fock++.cc:601-627 on the GPU
    (lb200_onebody_forces).  D, W: numpy arrays or torch CUDA tensors (both of one kind), nbf x nbf."""
    ch = np.ascontiguousarray([[z, c[0], c[1], c[2]] for z, c in charges], dtype=np.float64).reshape(-1, 4)
    s2a = np.ascontiguousarray(shell2atom, dtype=np.int32)
    if len(s2a) != basis.nshell:
        raise ValueError("onebody_forces: shell2atom needs one entry per shell")
    n = basis.nbf
    F1, FP = np.zeros((len(ch), 3)), np.zeros((len(ch), 3))
    if isinstance(D, np.ndarray) != isinstance(W, np.ndarray):
        raise ValueError("onebody_forces: D and W must both be numpy arrays or both be torch CUDA tensors")
    if isinstance(D, np.ndarray):
        D = np.ascontiguousarray(D, dtype=np.float64)
        W = np.ascontiguousarray(W, dtype=np.float64)
        pD, pW, on_dev = vp(D.ctypes.data), vp(W.ctypes.data), 0
    else:
        import torch
        for M in (D, W):
            if M.dtype != torch.float64 or not M.is_cuda or M.device.index != ctx.device:
                raise ValueError("onebody_forces: device inputs must be float64 CUDA tensors on the context's GPU")
        D, W = D.contiguous(), W.contiguous()
        pD, pW, on_dev = vp(D.data_ptr()), vp(W.data_ptr()), 1
    if tuple(D.shape) != (n, n) or tuple(W.shape) != (n, n):
        raise ValueError("onebody_forces: D and W must be nbf x nbf")
    ctx.check(load().lb200_onebody_forces(ctx.h, basis.h, len(ch), _d(ch), _i(s2a), pD, pW, on_dev, _d(F1), _d(FP)),
              "onebody_forces")
    return F1, FP


def eri_product(ctx, bra, ket, b0, nb, k0, nk, out, screening=SCREEN_ORIGINAL, precision=0.0, pure_out=False):
    """Engine::compute2 over the implicit product (bra pairs [b0, b0+nb)) x (ket pairs [k0, k0+nk)) into the
    torch CUDA tensor `out`."""
    ctx.check(load().lb200_eri_product(ctx.h, bra.h, ket.h, int(b0), int(nb), int(k0), int(nk), int(screening),
                                       float(precision), int(pure_out), vp(out.data_ptr())), "eri_product")
    return out


class Df3c:
    """(P|mu nu) slabs and the (P|Q) metric on the GPU (lb200_df3c_*)."""

    def __init__(self, ctx, obs, dfbs, pair_s1=None, pair_s2=None, threshold=1e-12):
        self.ctx, self.obs, self.dfbs = ctx, obs, dfbs
        if pair_s1 is None:
            pair_s1, pair_s2 = significant_pairs(obs, threshold)
        self.pair_s1 = np.ascontiguousarray(pair_s1, dtype=np.int32)
        self.pair_s2 = np.ascontiguousarray(pair_s2, dtype=np.int32)
        h = vp()
        ctx.check(load().lb200_df3c_create(ctx.h, obs.h, dfbs.h, len(self.pair_s1), _i(self.pair_s1),
                                           _i(self.pair_s2), C.byref(h)), "df3c_create")
        self.h = h
        info = (C.c_longlong * 6)()
        load().lb200_df3c_info(h, info)
        self.nbf, self.ndf, self.ndfshell, self.npairs, self.ntriplets, self.ngroups = [int(x) for x in info]

    def slab(self, Z, P0, nP, threshold=0.0, precision=0.0):
        """fills the torch CUDA float64 tensor Z[nPfun, nbf, nbf] for DF shells [P0, P0+nP); returns
        (triplets computed, triplets total)"""
        st = np.zeros(2)
        self.ctx.check(load().lb200_df3c_slab(self.h, int(P0), int(nP), float(threshold), float(precision),
                                              vp(Z.data_ptr()), _d(st)), "df3c_slab")
        return st[0], st[1]

    def metric(self, V):
        self.ctx.check(load().lb200_df3c_metric(self.h, vp(V.data_ptr())), "df3c_metric")
        return V

    def close(self):
        if self.h:
            load().lb200_df3c_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Comm:
    """NCCL communicator behind the C ABI (lb200_comm_*): the all-reduce of partial Fock matrices."""

    def __init__(self, ctx, nranks, rank, unique_id):
        self.ctx = ctx
        h = vp()
        ctx.check(load().lb200_comm_create(ctx.h, int(nranks), int(rank), bytes(unique_id), C.byref(h)),
                  "comm_create")
        self.h = h
        self.rank, self.nranks = rank, nranks

    @staticmethod
    def unique_id():
        buf = C.create_string_buffer(128)
        n = load().lb200_comm_unique_id(buf, 128)
        if n < 0:
            raise Lb200Error("lb200_comm_unique_id failed (%d): NCCL not available" % n)
        return buf.raw

    def allreduce_(self, G):
        """in-place sum over ranks of a torch CUDA float64 tensor, on the context's stream"""
        self.ctx.check(load().lb200_fock_allreduce(self.h, vp(G.data_ptr()), G.numel()), "fock_allreduce")
        return G

    def close(self):
        if self.h:
            load().lb200_comm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def task_owner(bra_pair_index, ket_pair_index, nranks):
    """Rank owning the quartet (bra pair | ket pair); pairs by canonical index s1(s1+1)/2+s2."""
    r = load().lb200_fock_task_owner(int(bra_pair_index), int(ket_pair_index), int(nranks))
    if r < 0:
        raise Lb200Error("lb200_fock_task_owner: invalid argument")
    return r


def fp64_peak_probe(ctx, iters=4096):
    """Measured FP64 FMA throughput (TFLOP/s, ms) of the context's GPU."""
    t = C.c_double(0)
    ms = C.c_double(0)
    ctx.check(load().lb200_fp64_peak_probe(ctx.h, int(iters), C.byref(t), C.byref(ms)), "fp64_probe")
    return t.value, ms.value


def eri_class_supported(la, lb, lc, ld):
    return bool(load().lb200_eri_class_supported(int(la), int(lb), int(lc), int(ld)))
