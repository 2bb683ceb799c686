"""Full-record-length parity of libfwi_b200.so against the reference's own op (oracle/_ref/libCUFD_ref.so) on a B200:

    python scripts/parity_full_length.py marmousi [nshots]   # real docs/data fixtures, 48-shot geometry of test/TestFWI.jl
    python scripts/parity_full_length.py c3 [nsteps]         # one shot on the C3 grid, 4000 steps
    python scripts/parity_full_length.py c5 [nsteps]         # one shot on the C5 grid, 8000 steps

Prints one JSON line per run: rel-L2 deviations of traces / misfit / gradients / grad_stf and the wall time of both
implementations for the gradient call (host buffers, their own file I/O).  TEST INFRASTRUCTURE (uses oracle/)."""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

from helpers import b200_cufd, interior_mask, rel
from oracle import oracle_py as op
from fwiflow.jl_b200 import synthetic
from fwiflow.jl_b200.utils import velocity_to_moduli


def marmousi_case(nshots=48):
    """test/TestFWI.jl:6-35 on the committed input fixtures (tests/golden/marmousi_inputs.npz)."""
    d = np.load(os.path.join(ROOT, "tests", "golden", "marmousi_inputs.npz"))
    xs = np.arange(4, 385, 8, dtype=np.int64)[:nshots]
    xr = np.arange(3, 382, dtype=np.int64)
    c = synthetic.Case(name="marmousi", nz=134, nx=384, dz=24.0, dx=24.0, dt=0.0025, nSteps=2000, f0=4.5,
                       z_src=np.full(xs.shape, 2, dtype=np.int64), x_src=xs, z_rec=np.full(xr.shape, 2, dtype=np.int64),
                       x_rec=xr)
    assert (c.nz_pad, c.nx_pad) == d["cp_true"].shape
    c.cp_true = d["cp_true"].astype(np.float64)
    c.cp_init = d["cp_init"].astype(np.float64)
    c.cs_true = np.zeros_like(c.cp_true); c.cs_init = np.zeros_like(c.cp_true)
    c.rho_true = np.full_like(c.cp_true, 2500.0); c.rho_init = np.full_like(c.cp_true, 2500.0)
    c.stf = np.repeat(d["stf"].astype(np.float64)[None, :], len(xs), axis=0)
    return c


def compare(c, ids, init, tag):
    lam, mu, rho = c.moduli("true")
    lam0, mu0, rho0 = init
    para_r = c.write_files(tempfile.mkdtemp(prefix=f"{tag}_ref_"))
    para_b = c.write_files(tempfile.mkdtemp(prefix=f"{tag}_b200_"))
    out = {"case": tag, "grid": [c.nz_pad, c.nx_pad], "shots": len(ids), "nSteps": c.nSteps}
    t0 = time.time(); ref_obs = op.ref_cufd(2, lam, mu, rho, c.stf, ids, para_r)["syn"]; out["ref_obs_s"] = time.time() - t0
    t0 = time.time(); b_obs = b200_cufd(2, lam, mu, rho, c.stf, ids, para_b)["syn"]; out["b200_obs_s"] = time.time() - t0
    out["traces"] = max(rel(a[:, 1:], b[:, 1:]) for a, b in zip(b_obs, ref_obs))
    out["traces_absmax"] = float(max(np.abs(b).max() for b in ref_obs))
    j_r = op.ref_cufd(0, lam0, mu0, rho0, c.stf, ids, para_r)["misfit"]
    j_b = b200_cufd(0, lam0, mu0, rho0, c.stf, ids, para_b)["misfit"]
    out["misfit_ref"] = j_r; out["misfit"] = abs(j_b - j_r) / j_r
    t0 = time.time(); g_r = op.ref_cufd(1, lam0, mu0, rho0, c.stf, ids, para_r); out["ref_grad_s"] = time.time() - t0
    t0 = time.time(); g_b = b200_cufd(1, lam0, mu0, rho0, c.stf, ids, para_b); out["b200_grad_s"] = time.time() - t0
    t0 = time.time(); g_b = b200_cufd(1, lam0, mu0, rho0, c.stf, ids, para_b); out["b200_grad_warm_s"] = time.time() - t0
    inner = interior_mask(c)
    for k in ("grad_lambda", "grad_mu", "grad_den"):
        out[k] = rel(g_b[k], g_r[k])
        out[k + "_mask"] = rel(g_b[k][inner], g_r[k][inner])
    out["grad_stf"] = rel(g_b["grad_stf"], g_r["grad_stf"])
    out["grad_stf_per_shot_max"] = max(rel(a, b) for a, b in zip(g_b["grad_stf"], g_r["grad_stf"]))
    # the reference run twice: its own run-to-run noise on the same quantities (atomics)
    if len(ids) * c.nz_pad * c.nx_pad * c.nSteps < 3e11:
        g_r2 = op.ref_cufd(1, lam0, mu0, rho0, c.stf, ids, para_r)
        out["ref_vs_ref"] = {k: rel(g_r2[k], g_r[k]) for k in ("grad_lambda", "grad_mu", "grad_den", "grad_stf")}
    print(json.dumps(out), flush=True)
    return out


def main():
    which = sys.argv[1]
    if which == "marmousi":
        n = int(sys.argv[2]) if len(sys.argv) > 2 else 48
        c = marmousi_case(n)
        ids = np.arange(n, dtype=np.int32)
        compare(c, ids, c.moduli("init"), "marmousi")
    else:
        full = {"c3": 4000, "c5": 8000}[which]
        n = int(sys.argv[2]) if len(sys.argv) > 2 else full
        c = {"c3": synthetic.case_c3, "c5": synthetic.case_c5}[which](nshots=1, nSteps=n)
        lam, mu, rho = c.moduli("true")
        compare(c, np.array([0], np.int32), (0.96 * lam, 0.97 * mu, rho), which)


if __name__ == "__main__":
    main()
