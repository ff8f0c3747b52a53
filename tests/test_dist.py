"""world_size-2 gloo test of the multi-GPU plumbing: each rank builds the partial G of the
quartets it owns and one all-reduce sums them.  On this CPU-only box the per-rank compute is
the oracle restricted to a task stride (test stand-in for the CUDA build); what is under test
is libint_b200.fock's rendezvous + reduction."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    from libint_b200 import fock
    from libint_b200.basis import BasisSet, H2O_XYZ_ANGSTROM, atoms_from_tuples
    from oracle import pyoracle as po
    r, w, _ = fock.init_distributed(backend="gloo")
    assert (r, w) == (rank, world)
    bs = BasisSet("6-31g", atoms_from_tuples(H2O_XYZ_ANGSTROM))
    ns = len(bs)
    s1, s2 = np.array([(a, b) for a in range(ns) for b in range(a + 1)], dtype=np.int32).T
    f = po.Fock(po.Shells(*bs.flat(), raw=False), s1, s2, nthreads=1)
    rng = np.random.default_rng(11)
    D = rng.standard_normal((bs.nbf, bs.nbf)) * 0.3
    D = 0.5 * (D + D.T)
    Gpart, _ = f.build(D, 1e-12, task_stride=world, task_offset=rank)
    G = torch.from_numpy(Gpart.copy())
    fock.allreduce_sum_(G)
    # the 3 * natoms partial forces go through the same process group (FockBuilder.forces_2body)
    part = np.arange(9, dtype=np.float64).reshape(3, 3) * (rank + 1)
    tot = fock.allreduce_forces(part)
    assert np.array_equal(tot, np.arange(9, dtype=np.float64).reshape(3, 3) * sum(range(1, world + 1)))
    assert np.array_equal(part, np.arange(9, dtype=np.float64).reshape(3, 3) * (rank + 1))   # input untouched
    if rank == 0:
        Gfull, _ = f.build(D, 1e-12)
        q.put(float(np.max(np.abs(G.numpy() - Gfull))))
    torch.distributed.destroy_process_group()


def test_partial_fock_allreduce_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=5) < 1e-13


def test_dist_env_defaults(monkeypatch):
    from libint_b200 import fock
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    assert fock.dist_env() == (0, 1, 0)
