"""N > 1 path on CPU: world_size-2 `gloo` processes run the ensemble sharding / gather / status-reduce logic of
bench.py and the host API (contiguous trajectory blocks, all_gather of solutions, all_reduce of counters) with the CPU
oracle standing in for the per-rank batch solve.  The data path has no collective; only the epilogue does."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, K, N, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="2")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import nonlinearsolve_jl_b200 as nls
    from oracle import pyoracle as po
    import bench
    n = 2 * N * N
    lo, hi = nls.shard_range(K, rank, world)
    A, B = bench.ensemble_params(K)
    P = po.OracleProblem.bruss2d(N)
    u0 = np.tile(P.u0(), (hi - lo, 1))
    u, resid, rc, ns, nj, res = po.ensemble_solve(N, u0, A[lo:hi], B[lo:hi], opts=po.default_newton_opts(abstol=1e-8, gmres_orth=po.ORTH_CGS2))
    # epilogue collectives, as in bench.run_ensemble / SURVEY.md §8e
    gathered = torch.empty(K * n, dtype=torch.float64)
    dist.all_gather_into_tensor(gathered, torch.from_numpy(u.ravel().copy()))
    stats = torch.tensor([res.nsuccess, res.total_nsteps, res.total_njvp], dtype=torch.float64)
    dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    worst = torch.tensor([res.worst_resid_inf], dtype=torch.float64)
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    if rank == 0:
        np.savez(os.path.join(out_dir, "gathered.npz"), u=gathered.numpy().reshape(K, n), stats=stats.numpy(), worst=worst.numpy())
    dist.destroy_process_group()


def test_ensemble_sharding_world2(tmp_path, po):
    K, N, world = 6, 8, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, K, N, str(tmp_path)), nprocs=world, join=True)
    got = np.load(os.path.join(str(tmp_path), "gathered.npz"))
    sys.path.insert(0, ROOT)
    import bench
    A, B = bench.ensemble_params(K)
    P = po.OracleProblem.bruss2d(N)
    u, resid, rc, ns, nj, res = po.ensemble_solve(N, np.tile(P.u0(), (K, 1)), A, B, opts=po.default_newton_opts(abstol=1e-8, gmres_orth=po.ORTH_CGS2))
    assert np.array_equal(got["u"], u)  # sharded + gathered == unsharded, in trajectory order
    assert got["stats"][0] == K and got["stats"][1] == ns.sum() and got["stats"][2] == nj.sum()
    assert got["worst"][0] == resid.max() < 1e-8
