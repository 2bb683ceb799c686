"""Which of {LEAN reverse kernel, record length, grid size} makes the large-grid gradient drift from the reference?
   python scripts/c5_debug.py c5 2000 [4000 ...]     (GPU box; TEST INFRASTRUCTURE: uses oracle/_ref)"""
import json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import b200_cufd, interior_mask, rel
from oracle import oracle_py as op
from fwiflow.jl_b200 import ops, synthetic

which = sys.argv[1]
for n in [int(v) for v in sys.argv[2:]]:
    c = {"c3": synthetic.case_c3, "c5": synthetic.case_c5, "c2": synthetic.case_c2}[which](nshots=1, nSteps=n)
    ids = np.array([0], np.int32)
    lam, mu, rho = c.moduli("true")
    lam0, mu0, rho0 = 0.96 * lam, 0.97 * mu, rho
    para_r = c.write_files(tempfile.mkdtemp(prefix="dbg_ref_"))
    para_b = c.write_files(tempfile.mkdtemp(prefix="dbg_b200_"))
    op.ref_cufd(2, lam, mu, rho, c.stf, ids, para_r)
    b200_cufd(2, lam, mu, rho, c.stf, ids, para_b)
    t0 = time.time(); g_r = op.ref_cufd(1, lam0, mu0, rho0, c.stf, ids, para_r); t_ref = time.time() - t0
    inner = interior_mask(c)
    runs = {}
    for lean in (0, 1):
        ops.set_option("rev_lean", lean)
        for rep in (0, 1):
            runs[(lean, rep)] = b200_cufd(1, lam0, mu0, rho0, c.stf, ids, para_b)
    ops.set_option("rev_lean", -1)
    out = {"case": which, "nSteps": n, "ref_grad_s": t_ref}
    for lean in (0, 1):
        g = runs[(lean, 0)]
        out[f"lean{lean}"] = {k: [rel(g[k], g_r[k]), rel(g[k][inner], g_r[k][inner])] for k in ("grad_lambda", "grad_mu", "grad_den")}
        out[f"lean{lean}"]["grad_stf"] = rel(g["grad_stf"], g_r["grad_stf"])
        out[f"lean{lean}_rerun_identical"] = all(np.array_equal(runs[(lean, 0)][k], runs[(lean, 1)][k])
                                                 for k in ("grad_lambda", "grad_mu", "grad_den", "grad_stf"))
    out["lean0_vs_lean1"] = {k: rel(runs[(0, 0)][k], runs[(1, 0)][k]) for k in ("grad_lambda", "grad_mu", "grad_den")}
    # where does the deviation live?  rel-L2 per horizontal band of 1/8 of the rows (lambda gradient, lean auto choice)
    g = runs[(1, 0)]["grad_lambda"]; r = g_r["grad_lambda"]
    nb = 8; h = g.shape[0] // nb
    out["bands_lambda_lean1"] = [rel(g[i * h:(i + 1) * h], r[i * h:(i + 1) * h]) for i in range(nb)]
    g = runs[(0, 0)]["grad_lambda"]
    out["bands_lambda_lean0"] = [rel(g[i * h:(i + 1) * h], r[i * h:(i + 1) * h]) for i in range(nb)]
    print(json.dumps(out), flush=True)
    ops.release()
