"""Host-side one-body integrals (overlap, kinetic, nuclear attraction) for the SCF driver.

NOT part of the GPU hot path: the reference's driver gets S, T, V from Engine::compute1
(tests/hartree-fock/hartree-fock++.cc:267-275, compute_1body_ints :1064-1152), an O(N^2) set-up
step that SURVEY.md section 8(f)1 lists as "next" for the device.  They are needed here only so
that libint_b200.scf can close the loop around the GPU Fock build and reproduce the reference's
golden SCF energies.  Evaluated with the McMurchie-Davidson Hermite expansion in the *same*
basis-function convention as the two-electron path: coefficients as renormalized by
Shell::renorm (shell.h:958-999), every Cartesian component of a shell carrying the
normalization of x^l, pure shells through the solid-harmonic coefficients of
solidharmonics.h:114-174 (tools/gen_sph_header.py), STANDARD component orders.
"""
import math

import numpy as np

from .tools.gen_sph_header import cart as _cart
from .tools.gen_sph_header import coeff as _sph_coeff


def boys(mmax, T):
    """F_m(T), m = 0..mmax, to ~1e-16 relative: convergent series + downward recursion for
    T < 35, asymptotic value + upward recursion above."""
    F = np.empty(mmax + 1)
    if T < 35.0:
        # F_m(T) = exp(-T) sum_k (2T)^k / ((2m+1)(2m+3)...(2m+2k+1))
        m = mmax
        term = 1.0 / (2 * m + 1)
        s = term
        k = 0
        while True:
            k += 1
            term *= 2.0 * T / (2 * m + 2 * k + 1)
            s += term
            if term < 1e-17 * s:
                break
        eT = math.exp(-T)
        F[m] = eT * s
        for mm in range(mmax - 1, -1, -1):
            F[mm] = (2.0 * T * F[mm + 1] + eT) / (2 * mm + 1)
    else:
        eT = math.exp(-T)
        F[0] = 0.5 * math.sqrt(math.pi / T) * math.erf(math.sqrt(T))
        for mm in range(mmax):
            F[mm + 1] = ((2 * mm + 1) * F[mm] - eT) / (2.0 * T)
    return F


def _hermite_E(la, lb, PA, PB, p):
    """E[i][j][t] of one dimension without the exp(-mu X_AB^2) factor."""
    E = np.zeros((la + 1, lb + 1, la + lb + 2))
    E[0, 0, 0] = 1.0
    o2p = 0.5 / p
    for i in range(la):
        for t in range(i + 2):
            E[i + 1, 0, t] = (o2p * E[i, 0, t - 1] if t > 0 else 0.0) + PA * E[i, 0, t] + \
                (t + 1) * E[i, 0, t + 1]
    for j in range(lb):
        for i in range(la + 1):
            for t in range(i + j + 2):
                E[i, j + 1, t] = (o2p * E[i, j, t - 1] if t > 0 else 0.0) + PB * E[i, j, t] + \
                    (t + 1) * E[i, j, t + 1]
    return E


def _hermite_R(L, p, PC):
    """R[t][u][v] = R^0_{tuv}(p, P - C), t+u+v <= L."""
    T = p * float(PC @ PC)
    Fm = boys(L, T)
    R = {}
    for n in range(L + 1):
        R[(n, 0, 0, 0)] = (-2.0 * p) ** n * Fm[n]

    def get(n, t, u, v):
        if t < 0 or u < 0 or v < 0:
            return 0.0
        key = (n, t, u, v)
        if key in R:
            return R[key]
        if t > 0:
            val = (t - 1) * get(n + 1, t - 2, u, v) + PC[0] * get(n + 1, t - 1, u, v)
        elif u > 0:
            val = (u - 1) * get(n + 1, t, u - 2, v) + PC[1] * get(n + 1, t, u - 1, v)
        else:
            val = (v - 1) * get(n + 1, t, u, v - 2) + PC[2] * get(n + 1, t, u, v - 1)
        R[key] = val
        return val

    out = np.zeros((L + 1, L + 1, L + 1))
    for t in range(L + 1):
        for u in range(L + 1 - t):
            for v in range(L + 1 - t - u):
                out[t, u, v] = get(0, t, u, v)
    return out


def cart_to_pure(l):
    """(2l+1) x ncart(l) matrix, rows m = -l..l (solidharmonics.h:114-174)."""
    c = _cart(l)
    M = np.zeros((2 * l + 1, len(c)))
    for m in range(-l, l + 1):
        for k, (lx, ly, lz) in enumerate(c):
            M[m + l, k] = _sph_coeff(l, m, lx, ly, lz)
    return M


def _shell_pair_blocks(sa, sb, charges, per_charge=False):
    """Cartesian S, T, V blocks of one shell pair (contracted); per_charge: V as one block per charge."""
    la, lb = sa.l, sb.l
    ca, cb = _cart(la), _cart(lb)
    S = np.zeros((len(ca), len(cb)))
    T = np.zeros_like(S)
    V = np.zeros((len(charges),) + S.shape) if per_charge else np.zeros_like(S)
    A, B = np.asarray(sa.O, float), np.asarray(sb.O, float)
    AB = A - B
    for a, wa in zip(sa.alpha, sa.coeff):
        for b, wb in zip(sb.alpha, sb.coeff):
            p = a + b
            P = (a * A + b * B) / p
            pref = wa * wb * math.exp(-a * b / p * float(AB @ AB))
            # two extra quanta on the ket for the kinetic-energy relation
            E = [_hermite_E(la, lb + 2, P[d] - A[d], P[d] - B[d], p) for d in range(3)]
            s1 = math.sqrt(math.pi / p)

            def S1(d, i, j):
                return E[d][i, j, 0] * s1 if j >= 0 else 0.0

            def T1(d, i, j):
                return -2.0 * b * b * S1(d, i, j + 2) + b * (2 * j + 1) * S1(d, i, j) - \
                    0.5 * j * (j - 1) * S1(d, i, j - 2)

            Rs = [(-Z * 2.0 * math.pi / p, _hermite_R(la + lb, p, P - np.asarray(C, float)))
                  for Z, C in charges]
            for ia, (ax, ay, az) in enumerate(ca):
                for ib, (bx, by, bz) in enumerate(cb):
                    sx, sy, sz = S1(0, ax, bx), S1(1, ay, by), S1(2, az, bz)
                    S[ia, ib] += pref * sx * sy * sz
                    T[ia, ib] += pref * (T1(0, ax, bx) * sy * sz + sx * T1(1, ay, by) * sz +
                                         sx * sy * T1(2, az, bz))
                    ex = E[0][ax, bx, :ax + bx + 1]
                    ey = E[1][ay, by, :ay + by + 1]
                    ez = E[2][az, bz, :az + bz + 1]
                    v = 0.0
                    for k, (f, R) in enumerate(Rs):
                        vk = f * np.einsum("t,u,v,tuv->", ex, ey, ez,
                                           R[:ax + bx + 1, :ay + by + 1, :az + bz + 1])
                        if per_charge:
                            V[k, ia, ib] += pref * vk
                        v += vk
                    if not per_charge:
                        V[ia, ib] += pref * v
    return S, T, V


def compute_1body_ints(obs, atoms):
    """-> (S, T, V) dense nbf x nbf; the three compute_1body_ints<Operator::overlap|kinetic|
    nuclear> calls of hartree-fock++.cc:267-275 (point charges = make_point_charges(atoms))."""
    n = obs.nbf
    S = np.zeros((n, n))
    T = np.zeros((n, n))
    V = np.zeros((n, n))
    charges = [(float(a.atomic_number), a.xyz) for a in atoms]
    c2p = {}
    for i, sa in enumerate(obs):
        for j in range(i + 1):
            sb = obs[j]
            blocks = _shell_pair_blocks(sa, sb, charges)
            out = []
            for M in blocks:
                if sa.pure:
                    M = c2p.setdefault(sa.l, cart_to_pure(sa.l)) @ M
                if sb.pure:
                    M = M @ c2p.setdefault(sb.l, cart_to_pure(sb.l)).T
                out.append(M)
            bi, bj = obs.shell2bf[i], obs.shell2bf[j]
            for dst, M in zip((S, T, V), out):
                dst[bi:bi + M.shape[0], bj:bj + M.shape[1]] = M
                dst[bj:bj + M.shape[1], bi:bi + M.shape[0]] = M.T
    return S, T, V


def nuclear_repulsion(atoms):
    """hartree-fock++.cc:245-255."""
    e = 0.0
    for i, a in enumerate(atoms):
        for b in atoms[:i]:
            e += a.atomic_number * b.atomic_number / float(np.linalg.norm(np.asarray(a.xyz) - np.asarray(b.xyz)))
    return e


# ---------------------------------------------------------------------------------------------------
# first derivatives of S, T, V in numpy: the CHECKER of the GPU one-body force kernel (lb200_onebody_forces,
# csrc/onebody_deriv.cu), not a product path -- scf.hf_forces and the C++ driver call the GPU entry.  These
# derivative matrices are pinned on the CPU against the reference's golden forces (tests/test_scf.py) and then
# serve as the reference of the kernel's CPU-compiled core and of the kernel itself.
# compute_1body_ints_deriv<overlap|kinetic|nuclear>(1, obs, atoms), hartree-fock++.cc:1154-1228, used at
# :601-627.  d/dA_x (a|O|b) = 2 alpha (a+1_x|O|b) - a_x (a-1_x|O|b); the derivative with respect to a
# nuclear position follows from translational invariance of every single-charge term.
# ---------------------------------------------------------------------------------------------------
class _S:
    def __init__(self, l, alpha, coeff, O):
        self.l, self.alpha, self.coeff, self.O = l, alpha, coeff, O


def _raised(sh):
    return _S(sh.l + 1, sh.alpha, np.asarray(sh.coeff) * 2.0 * np.asarray(sh.alpha), sh.O)


def _lowered(sh):
    return _S(sh.l - 1, sh.alpha, sh.coeff, sh.O)


def _deriv_first_index(up, dn, l):
    """Cartesian derivative blocks d/dA_x of the first index from the raised / lowered blocks
    (arrays [..., ncart(l+1), nb] and [..., ncart(l-1), nb] or None) -> [3][..., ncart(l), nb]"""
    c = _cart(l)
    iu = {q: k for k, q in enumerate(_cart(l + 1))}
    idn = {q: k for k, q in enumerate(_cart(l - 1))} if l > 0 else {}
    out = []
    for d in range(3):
        rows = []
        for q in c:
            qu = list(q)
            qu[d] += 1
            v = up[..., iu[tuple(qu)], :].copy()
            if q[d] > 0:
                qd = list(q)
                qd[d] -= 1
                v -= q[d] * dn[..., idn[tuple(qd)], :]
            rows.append(v)
        out.append(np.stack(rows, axis=-2))
    return out


def compute_1body_ints_deriv(obs, atoms):
    """-> (S1, T1, V1), each [3 * natoms, nbf, nbf]: first derivatives with respect to the nuclear
    coordinates (basis-function centres and, for V, the point charges)."""
    n = obs.nbf
    na = len(atoms)
    S1 = np.zeros((3 * na, n, n))
    T1 = np.zeros((3 * na, n, n))
    V1 = np.zeros((3 * na, n, n))
    charges = [(float(a.atomic_number), a.xyz) for a in atoms]
    c2p = {}

    def pure(M, sa, sb):
        if sa.pure:
            M = np.einsum("pc,...cb->...pb", c2p.setdefault(sa.l, cart_to_pure(sa.l)), M)
        if sb.pure:
            M = np.einsum("...ac,pc->...ap", M, c2p.setdefault(sb.l, cart_to_pure(sb.l)))
        return M

    def side(sa, sb):
        """derivatives with respect to the centre of sa of the (sa|O|sb) blocks: dS[3], dT[3], dV[3][ncharge]"""
        up = _shell_pair_blocks(_raised(sa), sb, charges, per_charge=True)
        dn = _shell_pair_blocks(_lowered(sa), sb, charges, per_charge=True) if sa.l > 0 else (None, None, None)
        return [_deriv_first_index(u, d, sa.l) for u, d in zip(up, dn)]

    for i, sa in enumerate(obs):
        for j in range(i + 1):
            sb = obs[j]
            dA = side(sa, sb)
            dB = [[np.swapaxes(x, -1, -2) for x in blk] for blk in side(sb, sa)]
            ai, aj = obs.shell2atom[i], obs.shell2atom[j]
            bi, bj = obs.shell2bf[i], obs.shell2bf[j]

            def add(dst, coord, M):
                M = pure(M, sa, sb)
                dst[coord, bi:bi + M.shape[0], bj:bj + M.shape[1]] += M
                if i != j:
                    dst[coord, bj:bj + M.shape[1], bi:bi + M.shape[0]] += M.T

            for d in range(3):
                add(S1, 3 * ai + d, dA[0][d])
                add(S1, 3 * aj + d, dB[0][d])
                add(T1, 3 * ai + d, dA[1][d])
                add(T1, 3 * aj + d, dB[1][d])
                add(V1, 3 * ai + d, dA[2][d].sum(axis=0))
                add(V1, 3 * aj + d, dB[2][d].sum(axis=0))
                for k in range(na):   # the operator's own centre: -(d/dA + d/dB) of that charge's term
                    add(V1, 3 * k + d, -(dA[2][d][k] + dB[2][d][k]))
    return S1, T1, V1


def nuclear_repulsion_forces(atoms):
    """hartree-fock++.cc:668-701."""
    F = np.zeros((len(atoms), 3))
    for i in range(1, len(atoms)):
        for j in range(i):
            r = np.asarray(atoms[i].xyz) - np.asarray(atoms[j].xyz)
            f = -r * atoms[i].atomic_number * atoms[j].atomic_number / float(np.linalg.norm(r)) ** 3
            F[i] += f
            F[j] -= f
    return F
