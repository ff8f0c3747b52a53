"""Density-fitted Fock builder on the GPU: G = 2 J - K from the three-centre integrals (P|mu nu) and the
two-centre metric (P|Q).

Mirrors DFFockEngine::compute_2body_fock_dfC of the reference's direct-SCF driver
(tests/hartree-fock/hartree-fock++.cc:2180-2334): there Zxy[ndf][n][n] is computed once, transformed
with L^-T of the Cholesky factor of V = (P|Q) into xyK[n][n][ndf] (107 GB for C40H82 / def2-TZVP /
def2-universal-JKFIT) and contracted with C_occ on every call.  Here the integrals are *streamed*: a
slab Z[P-range][n][n] (lb200_df3c_slab, class kernels + scatter on the device) is half-transformed with
C_occ while it is resident in HBM (cuBLAS DGEMM through torch.matmul), only W[ndf][n][nocc] is kept, the
metric is applied to W by a triangular solve (cuSOLVER/cuBLAS through torch.linalg), and the Coulomb
part takes a second sweep of the (cheap) integrals:

    W[P,x,i] = sum_y (P|xy) C[y,i]               per slab
    X        = L^-1 W          (V = L L^T)        K[x,y] = sum_{P,i} X[P,x,i] X[P,y,i]
    jt[P]    = sum_{x,i} X[P,x,i] C[x,i],   c = L^-T jt,    J[x,y] = sum_P c[P] (P|xy)     second sweep
    G = 2 J - K                                                      (hartree-fock++.cc:2296-2320)

The dense contractions are plain library GEMMs (FP64 tensor-core DGEMM in cuBLAS); the integral slabs are
this package's kernels.  Slabs are independent, so ranks may take disjoint DF-shell ranges (W rows).
"""
import numpy as np

from . import capi


class DFFockBuilder:
    def __init__(self, obs, dfbs, ctx=None, device=0, slab_bytes=8 << 30, threshold=0.0, pair_threshold=1e-12):
        import torch
        self.ctx = ctx or capi.Context(device)
        self.dev = torch.device("cuda", self.ctx.device)
        self.obs, self.dfbs = obs, dfbs
        self.B = capi.Basis(self.ctx, *obs.flat())
        self.Bdf = capi.Basis(self.ctx, *dfbs.flat())
        self.df = capi.Df3c(self.ctx, self.B, self.Bdf, threshold=pair_threshold)
        self.n, self.ndf = self.df.nbf, self.df.ndf
        self.threshold = threshold
        self.ctx.set_stream(torch.cuda.current_stream(self.dev).cuda_stream)
        # slabs: consecutive DF shells whose functions x n^2 doubles fit the budget
        sizes = np.array([s.size() for s in dfbs])
        per = max(1, int(slab_bytes // (8 * self.n * self.n)))
        self.slabs = []
        s0 = 0
        while s0 < len(sizes):
            s1, nf = s0, 0
            while s1 < len(sizes) and (s1 == s0 or nf + sizes[s1] <= per):
                nf += sizes[s1]
                s1 += 1
            self.slabs.append((s0, s1 - s0, int(nf)))
            s0 = s1
        self.maxf = max(nf for _, _, nf in self.slabs)
        # metric and its Cholesky factor (hartree-fock++.cc:2264-2270)
        V = torch.empty((self.ndf, self.ndf), dtype=torch.float64, device=self.dev)
        self.df.metric(V)
        self.L = torch.linalg.cholesky(V)
        self.stats = {}

    def _sweep(self):
        import torch
        Z = torch.empty((self.maxf, self.n, self.n), dtype=torch.float64, device=self.dev)
        row = 0
        for s0, ns, nf in self.slabs:
            done, total = self.df.slab(Z, s0, ns, threshold=self.threshold)
            yield row, nf, Z[:nf], done, total
            row += nf

    def __call__(self, Cocc):
        """G = 2J - K for D = Cocc Cocc^T (torch CUDA tensor [n, nocc] or numpy); returns a torch tensor."""
        import torch
        self.ctx.set_stream(torch.cuda.current_stream(self.dev).cuda_stream)
        C = torch.as_tensor(Cocc, dtype=torch.float64).to(self.dev).contiguous()
        n, ndf, nocc = self.n, self.ndf, C.shape[1]
        W = torch.empty((ndf, n, nocc), dtype=torch.float64, device=self.dev)
        done = total = 0.0
        for row, nf, Z, d, t in self._sweep():
            torch.matmul(Z.reshape(nf * n, n), C, out=W[row:row + nf].view(nf * n, nocc))
            done += d
            total += t
        X = torch.linalg.solve_triangular(self.L, W.view(ndf, n * nocc), upper=False).view(ndf, n, nocc)
        del W
        Xm = X.permute(1, 0, 2).reshape(n, ndf * nocc)
        K = Xm @ Xm.T
        jt = (X * C.unsqueeze(0)).sum(dim=(1, 2))
        c = torch.linalg.solve_triangular(self.L.T, jt.unsqueeze(1), upper=True).squeeze(1)
        del X, Xm
        J = torch.zeros((n, n), dtype=torch.float64, device=self.dev)
        for row, nf, Z, d, t in self._sweep():
            J += torch.tensordot(c[row:row + nf], Z, dims=1)
        self.stats = {"triplets_computed_per_sweep": done, "triplets_total": total, "sweeps": 2,
                      "slabs": len(self.slabs)}
        return 2.0 * J - K
