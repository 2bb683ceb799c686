"""CPU, world_size 2, gloo: the shot-sharding + all-reduce host logic of fwiflow.jl_b200.dist.
The compute on each rank is stood in by the CPU oracle (test infrastructure) because this box has no GPU;
on a GPU box the same `sharded_gradient` is driven with ops.Plan (tests/test_parity_gpu.py)."""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ROOT, golden_cases, rel
from fwiflow.jl_b200 import dist as fdist


def test_shard_is_a_true_partition():
    ids = np.arange(30)
    for ws in (1, 2, 4, 8):
        parts = [fdist.shard_shots(ids, r, ws) for r in range(ws)]
        assert sorted(np.concatenate(parts).tolist()) == ids.tolist()        # no shot twice (unlike TestFWI.jl:65)
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    assert [len(fdist.shard_shots(ids, r, 8)) for r in range(8)] == [4, 4, 4, 4, 4, 4, 3, 3]


class _OraclePlan:
    """Same surface as ops.Plan, computed by the oracle on the CPU (tests only)."""

    def __init__(self, para, ids, c):
        from oracle import oracle_py as op
        assert len(ids) > 0, "a plan needs at least one shot (fwi_b200_plan_create refuses group_size <= 0)"
        self.op, self.para, self.ids, self.c = op, para, np.asarray(ids, np.int32), c
        self.nz, self.nx = c.nz_pad, c.nx_pad
        self.buf = torch.zeros(3 * self.nz * self.nx + 1, dtype=torch.float32)

    def set_model(self, lam, mu, den):
        self.m = (lam, mu, den)

    def set_stf(self, stf):
        self.stf = stf

    def load_obs_files(self):
        pass

    def run(self, calc_id):
        g = self.op.oracle_cufd(1, *self.m, self.stf, self.ids, self.para, threads=2)
        j = self.op.oracle_cufd(0, *self.m, self.stf, self.ids, self.para, threads=2)["misfit"]
        n = self.nz * self.nx
        self.buf[:n] = torch.from_numpy(g["grad_lambda"].ravel()).float()
        self.buf[n:2 * n] = torch.from_numpy(g["grad_mu"].ravel()).float()
        self.buf[2 * n:3 * n] = torch.from_numpy(g["grad_den"].ravel()).float()
        self.buf[3 * n] = j

    def result_tensor(self):
        return self.buf


def _worker(rank, ws, port, wd, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    c = golden_cases()["small_elastic"]
    para = os.path.join(wd, "para_file.json")
    lam0, mu0, rho0 = c.moduli("init")
    ids = np.arange(c.nShots)
    res = fdist.sharded_gradient(lambda local: _OraclePlan(para, local, c), ids, rank, ws, lam0, mu0, rho0, c.stf)
    gs = fdist.gather_stf_grads(fdist.shard_shots(ids, rank, ws), np.full((len(fdist.shard_shots(ids, rank, ws)), 4), rank + 1.0),
                                len(ids), 4)
    if rank == 0:
        np.savez(out, misfit=res[0], gl=res[1], gm=res[2], gd=res[3], gs=gs)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_equals_single_process():
    from oracle import oracle_py as op
    c = golden_cases()["small_elastic"]
    wd = tempfile.mkdtemp()
    para = c.write_files(wd)
    lam, mu, rho = c.moduli("true")
    lam0, mu0, rho0 = c.moduli("init")
    ids = np.arange(c.nShots)
    op.oracle_cufd(2, lam, mu, rho, c.stf, ids, para)
    out = os.path.join(wd, "two_rank.npz")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, wd, out), nprocs=2, join=True)
    two = np.load(out)
    one = op.oracle_cufd(1, lam0, mu0, rho0, c.stf, ids, para)
    j1 = op.oracle_cufd(0, lam0, mu0, rho0, c.stf, ids, para)["misfit"]
    assert rel(two["gl"], one["grad_lambda"]) < 1e-5       # 1-rank vs 2-rank: summation order only
    assert rel(two["gm"], one["grad_mu"]) < 1e-5
    assert rel(two["gd"], one["grad_den"]) < 1e-5
    assert abs(float(two["misfit"]) - j1) <= 1e-5 * j1
    assert two["gs"][0, 0] == 1.0 and two["gs"][1, 0] == 2.0   # stf rows land on their GLOBAL shot id


def test_more_ranks_than_shots():
    """world_size 3, two shots: the rank without a shot creates no plan and contributes zeros, nobody blocks."""
    from oracle import oracle_py as op
    c = golden_cases()["small_elastic"]
    wd = tempfile.mkdtemp()
    para = c.write_files(wd)
    lam, mu, rho = c.moduli("true")
    lam0, mu0, rho0 = c.moduli("init")
    ids = np.arange(c.nShots)
    assert c.nShots == 2
    op.oracle_cufd(2, lam, mu, rho, c.stf, ids, para)
    out = os.path.join(wd, "three_rank.npz")
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(3, port, wd, out), nprocs=3, join=True)
    three = np.load(out)
    one = op.oracle_cufd(1, lam0, mu0, rho0, c.stf, ids, para)
    assert rel(three["gl"], one["grad_lambda"]) < 1e-5 and rel(three["gd"], one["grad_den"]) < 1e-5
    assert three["gs"][0, 0] == 1.0 and three["gs"][1, 0] == 2.0
