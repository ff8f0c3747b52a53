"""Host-side mirror of libint2::Engine for Operator::coulomb (deriv_order 0) on top of the C ABI.

Reference interface (evaleev/libint @ 7a1a9d8):
  Engine(Operator, max_nprim, max_l, deriv_order, precision, params, BraKet, ScreeningMethod)
                                               include/libint2/engine.h:503-526
  compute(s1, s2, s3, s4) / compute(s1, s2, s3)  engine.impl.h:139-175 (xs_xx inserts Shell::unit()
                                               as bra2, :165-167)
  results()[0] is None <=> all primitives screened out          engine.impl.h:1781-1784
  result = row-major n1*n2*n3*n4 in the caller's shell order, pure shells transformed
                                               engine.impl.h:1917-2093
  exceptions: lmax_exceeded (engine.h:893-916)

The per-quartet call is the correctness path (a batch of one shell set through
lb200_eri_batch); `compute_batch` is the entry point a real consumer uses.  There is no CPU
fallback: constructing an Engine needs a GPU context.
"""
import enum

import numpy as np

from . import capi
from .basis import BasisSet, Shell


class Operator(enum.Enum):
    coulomb = "coulomb"


class BraKet(enum.Enum):
    xx_xx = 0
    xs_xx = 1
    xs_xs = 2


class ScreeningMethod(enum.IntEnum):  # shell.h:1041-1059
    Original = capi.SCREEN_ORIGINAL
    Conservative = capi.SCREEN_CONSERVATIVE
    Schwarz = capi.SCREEN_SCHWARZ
    SchwarzInf = capi.SCREEN_SCHWARZ_INF


class lmax_exceeded(RuntimeError):
    """Engine::lmax_exceeded, engine.h:893-916."""


def basis_handle(ctx, shells):
    """Upload a list of Shell (or a BasisSet) as an lb200_basis."""
    bs = shells if isinstance(shells, BasisSet) else BasisSet(shells=list(shells))
    return capi.Basis(ctx, *bs.flat())


class Engine:
    def __init__(self, oper=Operator.coulomb, max_nprim=1, max_l=0, deriv_order=0,
                 precision=np.finfo(np.float64).eps, braket=BraKet.xx_xx,
                 screening_method=ScreeningMethod.Original, ctx=None, device=0):
        if oper != Operator.coulomb:
            raise NotImplementedError("only Operator.coulomb is on the B200 path")
        if deriv_order not in (0, 1):
            raise NotImplementedError("deriv_order 0 and 1 are on the B200 path")
        if deriv_order == 1 and braket != BraKet.xx_xx:
            raise NotImplementedError("first derivatives are built for BraKet.xx_xx")
        self.deriv_order = int(deriv_order)
        if max_l > capi.MAX_AM:
            raise lmax_exceeded("max_l=%d exceeds LB200_MAX_AM=%d" % (max_l, capi.MAX_AM))
        self.ctx = ctx if ctx is not None else capi.Context(device)
        self.max_nprim, self.max_l = int(max_nprim), int(max_l)
        self.braket = braket
        self.screening_method = ScreeningMethod(screening_method)
        self.set_precision(precision)
        self._results = [None] * (12 if deriv_order == 1 else 1)
        self._unit = None

    # Engine::set_precision, engine.h:809-826
    def set_precision(self, prec):
        self.precision = float(prec)
        return self

    def set(self, what):
        if isinstance(what, BraKet):
            self.braket = what
        elif isinstance(what, ScreeningMethod):
            self.screening_method = what
        else:
            raise TypeError(what)
        return self

    def results(self):
        return self._results

    def _pairs(self, sh_a, sh_b):
        """pair block for (a b| with the higher-l shell first; returns (pairs, swapped)."""
        swapped = sh_a.l < sh_b.l
        first, second = (sh_b, sh_a) if swapped else (sh_a, sh_b)
        for s in (first, second):
            if s.l > capi.MAX_AM:
                raise lmax_exceeded("shell l=%d exceeds LB200_MAX_AM=%d" % (s.l, capi.MAX_AM))
        b1 = basis_handle(self.ctx, [first])
        b2 = basis_handle(self.ctx, [second])
        # ShellPair::init with the engine's own ln_precision (engine.impl.h:1258-1276)
        lnp = np.log(self.precision) if self.precision > 0 else -np.inf
        scr = self._checked_screening()
        return capi.Pairs(self.ctx, b1, b2, [0], [0], int(scr), lnp), swapped, (b1, b2)

    def _checked_screening(self):
        """The Schwarz methods need ShellPairs built with a schwarz_factor_evaluator (shell.h:1259-1328);
        Engine::compute2 without precomputed pairs asserts on them in the reference
        (engine.impl.h:1259-1276 -> ShellPair::init).  This mirror has no per-call evaluator either:
        refuse instead of silently screening differently."""
        scr = self.screening_method
        if scr in (ScreeningMethod.Schwarz, ScreeningMethod.SchwarzInf):
            raise ValueError("Engine.compute: ScreeningMethod.%s needs precomputed shell pairs "
                             "(use capi.Pairs(..., screening=SCREEN_SCHWARZ_INF) / FockBuilder)" % scr.name)
        return scr

    def compute(self, *shells):
        """compute2<coulomb, braket, 0>; returns a flat array or None when screened out."""
        if self.braket == BraKet.xx_xx:
            if len(shells) != 4:
                raise ValueError("xx_xx needs 4 shells")
            s1, s2, s3, s4 = shells
        elif self.braket == BraKet.xs_xx:
            if len(shells) != 3:
                raise ValueError("xs_xx needs 3 shells")
            s1, s3, s4 = shells
            s2 = Shell.unit()
        else:
            if len(shells) != 2:
                raise ValueError("xs_xs needs 2 shells")
            s1, s3 = shells
            s2 = s4 = Shell.unit()
        bra, sw_b, keep1 = self._pairs(s1, s2)
        ket, sw_k, keep2 = self._pairs(s3, s4)
        scr = self._checked_screening()
        n = [s1.size(), s2.size(), s3.size(), s4.size()]
        nb = [n[1], n[0]] if sw_b else [n[0], n[1]]
        nk = [n[3], n[2]] if sw_k else [n[2], n[3]]
        if self.deriv_order == 1:
            # compute2<coulomb, xx_xx, 1>: twelve shell sets, index 3 * centre + xyz in the caller's shell
            # order (the reference re-maps the index when it permutes shells, engine.impl.h:1996-2003)
            if self.braket != BraKet.xx_xx:
                raise NotImplementedError("first derivatives are built for BraKet.xx_xx")
            if bra.nprimpair == 0 or ket.nprimpair == 0:
                self._results = [None] * 12
                return None
            out = capi.eri_deriv1_batch(self.ctx, bra, ket, np.array([[0, 0]], dtype=np.int32),
                                        screening=int(scr), precision=self.precision, pure_out=True)[0]
            centre = [1, 0] if sw_b else [0, 1]
            centre += [3, 2] if sw_k else [2, 3]
            res = [None] * 12
            for c_lib, c_user in enumerate(centre):
                for xyz in range(3):
                    t = out[3 * c_lib + xyz].reshape(nb + nk)
                    if sw_b:
                        t = t.transpose(1, 0, 2, 3)
                    if sw_k:
                        t = t.transpose(0, 1, 3, 2)
                    res[3 * c_user + xyz] = np.ascontiguousarray(t).ravel()
            self._results = res
            return res
        out = capi.eri_batch(self.ctx, bra, ket, np.array([[0, 0]], dtype=np.int32),
                             screening=int(scr), precision=self.precision, pure_out=True)[0]
        if bra.nprimpair == 0 or ket.nprimpair == 0:
            # every primitive pair screened out: the reference returns a null target
            # (a set whose primitive *quartets* are all screened yields zeros here)
            self._results = [None]
            return None
        t = out.reshape(nb + nk)
        if sw_b:
            t = t.transpose(1, 0, 2, 3)
        if sw_k:
            t = t.transpose(0, 1, 3, 2)
        res = np.ascontiguousarray(t).ravel()
        self._results = [res]
        return res
