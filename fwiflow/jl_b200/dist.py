"""Shot sharding across GPUs: one process per GPU, one NCCL all-reduce per gradient.

The reference shards shots by hand -- one `fwi_op` graph node per GPU with its own
`gpu_id` and a shot range, summed by TensorFlow on the host in float64
(test/TestFWI.jl:58-69, docs/src/tutorials/fwi_lowlevel.md:142-152); it has no
collective.  Here every rank owns a true partition of the shots, runs them through its
device-resident plan, and the packed float32 buffer [grad_lambda|grad_mu|grad_den|misfit]
(3*nz*nx+1 values) is summed in place with a single all-reduce over NVLink
(`torch.distributed`, backend nccl; gloo on CPU for the host-logic tests).
grad_stf rows are per shot, so they are gathered, not reduced.
"""
from __future__ import annotations

import numpy as np

__all__ = ["shard_shots", "allreduce_result", "gather_stf_grads", "sharded_gradient"]


def shard_shots(shot_ids, rank, world_size):
    """Round-robin partition (rank r gets ids[r::world_size]).  Unlike the reference's inclusive ranges
    (test/TestFWI.jl:65) no shot is processed twice; the union over ranks is exactly `shot_ids`."""
    ids = np.asarray(shot_ids, dtype=np.int32).ravel()
    return np.ascontiguousarray(ids[rank::world_size])


def allreduce_result(buf):
    """Sum the packed result buffer over all ranks, in place.  `buf` is a torch tensor (CUDA for nccl)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    return buf


def gather_stf_grads(local_ids, local_rows, n_total, nSteps):
    """All-gather the per-shot stf gradients into a (n_total, nSteps) array indexed by GLOBAL shot id."""
    import torch
    import torch.distributed as dist
    out = np.zeros((n_total, nSteps), np.float64)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        out[np.asarray(local_ids)] = local_rows
        return out
    objs = [None] * dist.get_world_size()
    dist.all_gather_object(objs, (np.asarray(local_ids), np.asarray(local_rows)))
    for ids, rows in objs:
        if len(ids):
            out[ids] = rows
    return out


def sharded_gradient(plan_factory, shot_ids, rank, world_size, lam, mu, den, stf, device=None):
    """Gradient of the whole survey computed by `world_size` ranks.

    plan_factory(local_ids) -> object with set_model / set_stf / load_obs_files / run(1) /
    result_tensor() (a torch tensor of 3*nz*nx+1 float32 on `device`) / result() -- i.e.
    `fwiflow.jl_b200.ops.Plan` on a GPU box.  Returns (misfit, gl, gm, gd) as float64 arrays
    identical on every rank.
    """
    local = shard_shots(shot_ids, rank, world_size)
    if len(local):
        plan = plan_factory(local)
        plan.set_model(lam, mu, den)
        plan.set_stf(stf)
        plan.load_obs_files()
        plan.run(1)
        buf = plan.result_tensor()
        nz, nx = plan.nz, plan.nx
    else:
        # more ranks than shots: this rank owns nothing.  No plan is created (a plan needs at least one shot); it
        # contributes zeros of the right size to the all-reduce so that the other ranks do not block.
        import torch
        nz, nx = np.asarray(lam).shape
        buf = torch.zeros(3 * nz * nx + 1, dtype=torch.float32, device=device if device is not None else "cpu")
    allreduce_result(buf)
    h = buf.detach().cpu().numpy().astype(np.float64)
    n = nz * nx
    return float(h[3 * n]), h[:n].reshape(nz, nx), h[n:2 * n].reshape(nz, nx), h[2 * n:3 * n].reshape(nz, nx)
