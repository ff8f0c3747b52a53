"""Restricted Hartree-Fock driver around the GPU Fock build: the consumer that the reference's
tests/hartree-fock/hartree-fock++.cc is for libint (main(), :233-520).

Same steps and conventions as the reference driver: nuclear repulsion (:245-255), S/T/V
(:267-275), conditioned orthogonalizer X with XtX condition number (:281-290, :1957-2006),
core-Hamiltonian start (:300-303 when the basis is minimal; the reference's SOAD start for
larger bases converges to the same ground state), density D = C_occ C_occ^T, energy
E = sum D o (H + F) + E_nuc (:472), commutator error ||FDS - SDF|| / n^2 (:476-477), DIIS
(diis.h) and the "totally empirical" Fock precision schedule (:463-466).  The two-electron
part F - H = G(D) comes from `fock_builder`, any callable (D, precision) -> G: the GPU
FockBuilder in production; tests also feed the CPU oracle's G through the same loop to pin the
oracle against the reference's golden energies.
"""
import numpy as np

from . import onebody


class DIIS:
    """libint2::DIIS (include/libint2/diis.h): Pulay extrapolation over the last `ndiis`
    (Fock, error) pairs, started after `strt` iterations."""

    def __init__(self, strt=2, ndiis=5):
        self.strt, self.ndiis = strt, ndiis
        self.x, self.e = [], []
        self.iter = 0

    def extrapolate(self, F, err):
        self.iter += 1
        self.x.append(F.copy())
        self.e.append(err.copy())
        if len(self.x) > self.ndiis:
            self.x.pop(0)
            self.e.pop(0)
        n = len(self.x)
        if self.iter < self.strt or n < 2:
            return F
        B = np.empty((n + 1, n + 1))
        for i in range(n):
            for j in range(n):
                B[i, j] = float(np.vdot(self.e[i], self.e[j]))
        scale = max(abs(B[:n, :n]).max(), 1e-300)
        B[:n, :n] /= scale
        B[n, :n] = B[:n, n] = -1.0
        B[n, n] = 0.0
        rhs = np.zeros(n + 1)
        rhs[n] = -1.0
        try:
            c = np.linalg.solve(B, rhs)[:n]
        except np.linalg.LinAlgError:
            return F
        return sum(ci * xi for ci, xi in zip(c, self.x))


def conditioning_orthogonalizer(S, max_condition_number=1e8):
    """-> (X, Xinv-less) = (S^-1/2 restricted to the well-conditioned subspace, rank, condition
    number of the retained part); gensqrtinv of hartree-fock++.cc:1957-2006."""
    w, U = np.linalg.eigh(S)
    wmax = w[-1]
    keep = w >= wmax / max_condition_number
    cond = wmax / w[keep][0]
    X = U[:, keep] / np.sqrt(w[keep])
    return X, int(keep.sum()), float(cond)


class RHF:
    def __init__(self, obs, atoms, fock_builder, charge=0, stv=None):
        """stv: (S, T, V) numpy matrices, e.g. from the GPU (capi.onebody); default: the host numpy evaluation
        (onebody.compute_1body_ints), which the CPU tests use to pin the oracle to the golden energies."""
        self.obs, self.atoms, self.fock_builder = obs, atoms, fock_builder
        nelec = sum(a.atomic_number for a in atoms) - charge
        if nelec % 2:
            raise ValueError("RHF needs an even number of electrons")
        self.ndocc = nelec // 2
        self.enuc = onebody.nuclear_repulsion(atoms)
        self.S, self.T, self.V = stv if stv is not None else onebody.compute_1body_ints(obs, atoms)
        self.H = self.T + self.V
        self.X, self.rank, self.cond = conditioning_orthogonalizer(self.S)
        self.history = []
        self.energy = None
        self.D = self._density(self.H)

    def _density(self, F):
        e, Cp = np.linalg.eigh(self.X.T @ F @ self.X)
        self.evals = e
        self.C = self.X @ Cp
        Co = self.C[:, :self.ndocc]
        return Co @ Co.T

    def run(self, conv=1e-12, maxiter=100, verbose=False):
        H, S = self.H, self.S
        n2 = H.size
        diis = DIIS(2)
        D = self.D
        ehf, rms, it = 0.0, 1.0, 0
        eps = np.finfo(float).eps
        while True:
            it += 1
            ehf_last = ehf
            precision = min(min(1e-3 / self.cond, 1e-7), max(rms / 1e4, eps))
            F = H + np.asarray(self.fock_builder(D, precision))
            ehf = float(np.sum(D * (H + F)))
            ediff_rel = abs((ehf - ehf_last) / ehf)
            comm = F @ D @ S - S @ D @ F
            rms = float(np.linalg.norm(comm)) / n2
            D = self._density(diis.extrapolate(F, comm))
            self.history.append((it, ehf + self.enuc, ediff_rel, rms))
            if verbose:
                print(" %02d %20.12f %20.12e %20.12e" % self.history[-1])
            if not ((ediff_rel > conv or rms > conv) and it < maxiter):
                break
        self.D, self.F = D, F
        self.converged = ediff_rel <= conv and rms <= conv
        self.energy = ehf + self.enuc
        return self.energy


class RHFDevice:
    """The same driver with every step on the GPU: S, T, V from lb200_onebody, the
    orthogonalizer and the Fock diagonalisation through cuSOLVER (torch.linalg.eigh), density, energy,
    commutator and DIIS as torch tensors on the device, G(D) from the GPU FockBuilder with device
    buffers -- per iteration only scalars (energy, error norm) reach the host.  `builder` is a
    libint_b200.fock.FockBuilder (its context, basis and rank set-up are reused)."""

    def __init__(self, obs, atoms, builder, charge=0):
        import torch
        from . import capi
        self.torch = torch
        self.obs, self.atoms, self.builder = obs, atoms, builder
        self.dev = torch.device("cuda", builder.ctx.device)
        nelec = sum(a.atomic_number for a in atoms) - charge
        if nelec % 2:
            raise ValueError("RHF needs an even number of electrons")
        self.ndocc = nelec // 2
        self.enuc = onebody.nuclear_repulsion(atoms)
        builder.ctx.set_stream(torch.cuda.current_stream(self.dev).cuda_stream)
        charges = [(float(a.atomic_number), a.xyz) for a in atoms]
        self.S, self.T, self.V = capi.onebody(builder.ctx, builder.basis, charges, device=True)
        self.H = self.T + self.V
        w, U = torch.linalg.eigh(self.S)            # gensqrtinv, hartree-fock++.cc:1957-2006
        keep = w >= w[-1] / 1e8
        self.cond = float(w[-1] / w[keep][0])
        self.rank = int(keep.sum())
        self.X = U[:, keep] / torch.sqrt(w[keep])
        self.history = []
        self.energy = None
        self.D = self._density(self.H)

    def _density(self, F):
        torch = self.torch
        e, Cp = torch.linalg.eigh(self.X.T @ F @ self.X)
        self.evals = e
        self.C = self.X @ Cp
        Co = self.C[:, :self.ndocc]
        return Co @ Co.T

    def run(self, conv=1e-12, maxiter=100, verbose=False, incremental=True):
        """incremental=True: the reference's incremental Fock formation (hartree-fock++.cc:420-480): once
        the error is below 1e-5, F += G(D - D_last) with a density difference whose shell-block norms -- and
        so the number of quartets surviving the Schwarz x density screen -- shrink from iteration to
        iteration; reset to a full build when the error fell 10x or after 8 iterations."""
        torch = self.torch
        H, S = self.H, self.S
        n2 = H.numel()
        xs, es = [], []          # DIIS history (libint2::DIIS, start 2, depth 5)
        D = self.D
        ehf, rms, it = 0.0, 1.0, 0
        eps = float(np.finfo(float).eps)
        D_diff, F = D, H
        reset, started = False, False
        start_thr = 1e-5 if incremental else 0.0
        next_reset_thr, last_reset_it = 0.0, 0
        self.full_builds = self.incremental_builds = 0
        while True:
            it += 1
            ehf_last = ehf
            D_last = D
            if not started and rms < start_thr:
                started, reset = True, False
                last_reset_it = it - 1
                next_reset_thr = rms / 10.0
            if reset or not started:
                F = H
                D_diff = D
            if reset and started:
                reset = False
                last_reset_it = it
                next_reset_thr = rms / 10.0
            precision = min(min(1e-3 / self.cond, 1e-7), max(rms / 1e4, eps))
            if D_diff is D:
                self.full_builds += 1
            else:
                self.incremental_builds += 1
            F = F + self.builder(D_diff, precision)
            ehf = float((D * (H + F)).sum())
            ediff_rel = abs((ehf - ehf_last) / ehf)
            comm = F @ D @ S - S @ D @ F
            rms = float(torch.linalg.norm(comm)) / n2
            if rms < next_reset_thr or it - last_reset_it >= 8:
                reset = True
            xs.append(F.clone())
            es.append(comm)
            if len(xs) > 5:
                xs.pop(0)
                es.pop(0)
            Fx = F
            n = len(xs)
            if it >= 2 and n >= 2:
                E = torch.stack([e.reshape(-1) for e in es])
                B = torch.empty((n + 1, n + 1), dtype=torch.float64, device=self.dev)
                B[:n, :n] = E @ E.T
                B[:n, :n] /= max(float(B[:n, :n].abs().max()), 1e-300)
                B[n, :n] = -1.0
                B[:n, n] = -1.0
                B[n, n] = 0.0
                rhs = torch.zeros(n + 1, dtype=torch.float64, device=self.dev)
                rhs[n] = -1.0
                try:
                    c = torch.linalg.solve(B, rhs)[:n]
                    Fx = sum(ci * xi for ci, xi in zip(c, xs))
                except RuntimeError:
                    Fx = F
            D = self._density(Fx)
            D_diff = D - D_last
            self.history.append((it, ehf + self.enuc, ediff_rel, rms))
            if verbose:
                print(" %02d %20.12f %20.12e %20.12e" % self.history[-1])
            if not ((ediff_rel > conv or rms > conv) and it < maxiter):
                break
        self.D, self.F = D, F
        self.converged = ediff_rel <= conv and rms <= conv
        self.energy = ehf + self.enuc
        return self.energy


def hf_forces(scf, builder, precision=None, use_schwarz=False):
    """The force block of the reference's driver (tests/hartree-fock/hartree-fock++.cc:596-716) for a
    converged RHF / RHFDevice object: one-body, Pulay, two-body and nuclear-repulsion contributions and their
    sum, each [natoms, 3].  One-body and Pulay parts: FockBuilder.forces_1body (lb200_onebody_forces: the
    derivative integrals of compute_1body_ints_deriv contracted with D and W on the GPU); two-body part:
    FockBuilder.forces_2body (compute_2body_fock_deriv<1> traced with D on the GPU)."""
    to_np = (lambda x: x.detach().cpu().numpy()) if hasattr(scf.D, "detach") else np.asarray
    D, C, ev = to_np(scf.D), to_np(scf.C), to_np(scf.evals)
    atoms = scf.atoms
    Co = C[:, :scf.ndocc]
    W = (Co * ev[:scf.ndocc]) @ Co.T                                              # :617-619
    F1, FP = builder.forces_1body(D, W, atoms)                                    # :601-627
    if precision is None:
        precision = np.finfo(float).eps      # compute_2body_fock_deriv's default, hartree-fock++.cc:158-162
    # the reference calls compute_2body_fock_deriv<1>(obs, atoms, D) without a Schwarz matrix (:649): no screening
    F2 = builder.forces_2body(D, precision=precision, use_schwarz=use_schwarz)   # :648-656
    FN = onebody.nuclear_repulsion_forces(atoms)                                  # :668-701
    return {"1-body": F1, "Pulay": FP, "2-body": F2, "nuclear repulsion": FN, "Hartree-Fock": F1 + FP + F2 + FN}
