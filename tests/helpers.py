"""Shared helpers of the parity tests (test infrastructure)."""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, GOLDEN)

from make_golden import golden_cases, rel, run_case  # noqa: E402,F401

# tolerances stated by BASELINE.json's north_star
TOL_TRACE = 1e-4
TOL_GRAD = 1e-3
TOL_MISFIT = 1e-4


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, f"{name}.npz")))


def b200_cufd(calc_id, lam, mu, den, stf, shot_ids, para_fname, gpu_id=0):
    """Same call shape as oracle_py.oracle_cufd / ref_cufd, routed through the C ABI of libfwi_b200.so."""
    from fwiflow.jl_b200 import ops
    from oracle import oracle_py as op
    para = op.read_para(para_fname)
    nSteps = para["nSteps"]
    out = {}
    if calc_id == 2:
        ops.fwi_obs_op(lam, mu, den, stf, gpu_id, shot_ids, para_fname)
        out["syn"] = [np.fromfile(os.path.join(para["data_dir_name"], f"Shot{int(s)}.bin"), np.float32).reshape(-1, nSteps)
                      for s in shot_ids]
        out["misfit"] = 0.0
    elif calc_id == 0:
        out["misfit"] = ops.fwi_op(lam, mu, den, stf, gpu_id, shot_ids, para_fname)
    else:
        gl, gm, gd, gs = ops.fwi_op_grad(lam, mu, den, stf, gpu_id, shot_ids, para_fname)
        out.update(grad_lambda=gl, grad_mu=gm, grad_den=gd, grad_stf=gs[np.asarray(shot_ids)])
    return out


def interior_mask(c, margin_top=10, margin=0):
    """The reference's gradient mask (src/FWI.jl:46-48): inside the PML, minus the 10 rows under the top PML."""
    m = np.zeros((c.nz_pad, c.nx_pad), bool)
    m[c.nPml + margin_top:c.nPml + c.nz - margin, c.nPml + margin:c.nPml + c.nx - margin] = True
    return m


def away_from_sources(c, radius=4):
    """True everywhere except within `radius` cells of a source: res = obs - syn cancels catastrophically at a
    receiver sitting on the source, so adjoint-dependent values there are float32 noise."""
    m = np.ones((c.nz_pad, c.nx_pad), bool)
    for zs, xs in zip(c.z_src + c.nPml, c.x_src + c.nPml):
        m[max(zs - radius, 0):zs + radius + 1, max(xs - radius, 0):xs + radius + 1] = False
    return m
