"""Direct Fock (J - K/2) builder: the consumer of the ERI path, sharded over GPUs.

Mirrors compute_2body_fock of the reference's direct-SCF driver
(tests/hartree-fock/hartree-fock++.cc:1574-1772; python spelling
python/src/libint2/engine.cc:230-277): G(D) = 1/2 (g + g^T) with
g_12 += D_34 (12|34) deg, g_13 -= 1/4 D_24 (12|34) deg, ... over the unique, Schwarz x density
screened shell quartets.  The reference runs one Engine per CPU thread with static
round-robin and sums thread-private G's (:1665,:1753-1755); here every GPU (one process per
GPU, torch.distributed/NCCL) takes the quartets it owns (hash partition inside the screening
kernel), accumulates a partial G in its own HBM, and one all-reduce(sum, FP64) over NVLink
combines them.
"""
import os

import numpy as np

from . import capi
from .basis import BasisSet


def dist_env():
    """(rank, world_size, local_rank) from the torchrun environment."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init_distributed(backend=None):
    """One process per GPU; rendezvous from MASTER_ADDR/MASTER_PORT (torchrun)."""
    import torch
    import torch.distributed as dist
    rank, world, local = dist_env()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def allreduce_sum_(G):
    """In-place sum over ranks of a partial Fock matrix (torch tensor, FP64)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(G, op=dist.ReduceOp.SUM)
    return G


def allreduce_forces(g, device=None):
    """Sum over ranks of the partial two-body forces (numpy [natoms, 3]); 3 * natoms doubles, so the plain
    torch.distributed collective is used (NCCL needs the buffer on the rank's GPU, gloo takes it as it is)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return g
    t = torch.as_tensor(np.ascontiguousarray(g, dtype=np.float64)).clone()
    if dist.get_backend() == "nccl":
        t = t.to(torch.device("cuda", torch.cuda.current_device() if device is None else device))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def make_comm(ctx):
    """NCCL communicator of the C ABI (lb200_comm_create) for the ranks of the current
    torch.distributed job: rank 0's ncclUniqueId travels over the process group that torchrun set up
    (plumbing only); the all-reduce itself is lb200_fock_allreduce.  None on a single rank."""
    import torch.distributed as dist
    rank, world, _ = dist_env()
    if world == 1 or not (dist.is_available() and dist.is_initialized()):
        return None
    box = [capi.Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return capi.Comm(ctx, world, rank, box[0])


class FockBuilder:
    def __init__(self, obs, ctx=None, device=None, pair_threshold=1e-12, rank=None, nranks=None):
        r, w, local = dist_env()
        self.rank = r if rank is None else rank
        self.nranks = w if nranks is None else nranks
        if ctx is None:
            ctx = capi.Context(local if device is None else device)
        self.ctx = ctx
        self.obs = obs if isinstance(obs, BasisSet) else BasisSet(shells=list(obs))
        self.basis = capi.Basis(ctx, *self.obs.flat())
        self.fock = capi.Fock(ctx, self.basis, threshold=pair_threshold)
        self.nbf = self.basis.nbf
        self.comm = None   # created on first use (needs a GPU and an initialised process group)

    def schwarz(self):
        return self.fock.schwarz()

    def build_partial(self, D, precision, out=None, use_schwarz=True, stats=False):
        """This rank's share of G (no communication). D/out: numpy or torch CUDA tensors."""
        return self.fock.build(D, precision, use_schwarz=use_schwarz, rank=self.rank,
                               nranks=self.nranks, out=out, stats=stats)

    def forces_2body(self, D, shell2atom=None, natoms=None, precision=1e-12, use_schwarz=True, stats=False):
        """Two-body contribution to the forces, F2[natoms, 3] (hartree-fock++.cc:642-656: the trace of
        compute_2body_fock_deriv<1> with D), summed over ranks.  shell2atom defaults to the map the
        BasisSet was built with (BasisSet::shell2atom, basis.h.in)."""
        if shell2atom is None:
            shell2atom = self.obs.shell2atom
        if natoms is None:
            natoms = max(shell2atom) + 1
        if min(shell2atom) < 0:
            raise ValueError("forces_2body needs shell2atom (the basis was built from bare shells)")
        if not isinstance(D, np.ndarray):
            import torch
            self.ctx.set_stream(torch.cuda.current_stream(torch.device("cuda", self.ctx.device)).cuda_stream)
        out = self.fock.gradient(D, shell2atom, natoms, precision, use_schwarz=use_schwarz, rank=self.rank,
                                 nranks=self.nranks, stats=stats)
        g, st = out if stats else (out, None)
        if self.nranks > 1:
            g = allreduce_forces(g, self.ctx.device)
        return (g, st) if stats else g

    def forces_1body(self, D, W, atoms):
        """One-body and Pulay contributions to the forces, (F1, F_Pulay) each [natoms, 3]
        (hartree-fock++.cc:601-627; lb200_onebody_forces on the GPU).  D: density, W = C_occ eps_occ C_occ^T, numpy
        or torch CUDA tensors; atoms: the molecule (point charges and force centres).  Every rank evaluates the
        whole (set-up-sized) sum: no reduction."""
        shell2atom = self.obs.shell2atom
        if len(shell2atom) == 0 or min(shell2atom) < 0:
            raise ValueError("forces_1body needs shell2atom (the basis was built from bare shells)")
        if not isinstance(D, np.ndarray):
            import torch
            self.ctx.set_stream(torch.cuda.current_stream(torch.device("cuda", self.ctx.device)).cuda_stream)
        charges = [(float(a.atomic_number), a.xyz) for a in atoms]
        return capi.onebody_forces(self.ctx, self.basis, charges, shell2atom, D, W)

    def __call__(self, D, precision=1e-12, use_schwarz=True):
        """Full G on every rank. With torch CUDA input the result is a torch CUDA tensor and the
        reduction is NCCL; with numpy input on one rank the result is numpy."""
        import torch
        if isinstance(D, np.ndarray) and self.nranks == 1:
            return self.build_partial(D, precision, use_schwarz=use_schwarz)
        dev = torch.device("cuda", self.ctx.device)
        Dt = torch.as_tensor(D, dtype=torch.float64).to(dev)
        G = torch.empty((self.nbf, self.nbf), dtype=torch.float64, device=dev)
        self.ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
        self.build_partial(Dt, precision, out=G, use_schwarz=use_schwarz)
        if self.nranks > 1:
            if self.comm is None:
                self.comm = make_comm(self.ctx)
            if self.comm is not None:
                self.comm.allreduce_(G)      # ncclAllReduce on the context's (= torch's current) stream
            else:
                allreduce_sum_(G)
        return G
