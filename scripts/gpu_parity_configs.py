"""Deviations of the CUDA path from the reference's own op (oracle/_ref) at the BASELINE grid sizes.
   python scripts/gpu_parity_configs.py     (GPU box)"""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import b200_cufd, interior_mask, rel
from oracle import oracle_py as op
from fwiflow.jl_b200 import synthetic
from fwiflow.jl_b200.utils import sourceGene

def run(name, c, ids, lam0, mu0, rho0):
    lam, mu, rho = c.moduli("true")
    pr = c.write_files(tempfile.mkdtemp(prefix=f"{name}_r_")); pb = c.write_files(tempfile.mkdtemp(prefix=f"{name}_b_"))
    ro = op.ref_cufd(2, lam, mu, rho, c.stf, ids, pr)["syn"]; bo = b200_cufd(2, lam, mu, rho, c.stf, ids, pb)["syn"]
    tr = max(rel(a[:, 1:], b[:, 1:]) for a, b in zip(bo, ro))
    jr = op.ref_cufd(0, lam0, mu0, rho0, c.stf, ids, pr)["misfit"]; jb = b200_cufd(0, lam0, mu0, rho0, c.stf, ids, pb)["misfit"]
    gr = op.ref_cufd(1, lam0, mu0, rho0, c.stf, ids, pr); gb = b200_cufd(1, lam0, mu0, rho0, c.stf, ids, pb)
    m = interior_mask(c)
    print(f"| {name} | {c.nz_pad}x{c.nx_pad}, {len(ids)} shot(s), {c.nSteps} steps | {tr:.1e} | {abs(jb-jr)/jr:.1e} | " +
          " / ".join(f"{rel(gb[k][m], gr[k][m]):.1e}" for k in ("grad_lambda", "grad_mu", "grad_den")) + " | " +
          " / ".join(f"{rel(gb[k], gr[k]):.1e}" for k in ("grad_lambda", "grad_mu", "grad_den")) + " |", flush=True)

print("| config | size | traces | misfit | gradients, reference mask (lambda / mu / rho) | gradients, whole grid |")
print("|---|---|---|---|---|---|")
c = synthetic.case_c2(nshots=30, nSteps=2000)
run("C2", c, np.array([0, 9, 17, 29], np.int32), *c.moduli("init"))
for which, n in (("c3", 900), ("c5", 260)):
    c = {"c3": synthetic.case_c3, "c5": synthetic.case_c5}[which](nshots=1, nSteps=n)
    c.stf = sourceGene(15.0, n, c.dt)
    lam, mu, rho = c.moduli("true")
    run(which.upper(), c, np.array([0], np.int32), 0.96 * lam, 0.97 * mu, rho)
