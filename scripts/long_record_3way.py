"""Three-way comparison at long record lengths: reference op (GPU) vs CPU oracle vs libfwi_b200, one shot.
Answers "is a gradient deviation that grows with the record length OURS, or the float32 noise floor of the algorithm?":
the oracle is a loop-for-loop restatement of the reference (FP64 promotions included), so oracle-vs-reference is the
deviation any faithful float32 implementation shows.
   python scripts/long_record_3way.py c2 2000 8000 20000     (GPU box; TEST INFRASTRUCTURE: uses oracle/)"""
import json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import b200_cufd, interior_mask, rel
from oracle import oracle_py as op
from fwiflow.jl_b200 import ops, synthetic
from fwiflow.jl_b200.utils import sourceGene

which = sys.argv[1]
for n in [int(v) for v in sys.argv[2:]]:
    if which == "c2":
        c = synthetic.case_c2(nshots=30, nSteps=n); ids = np.array([14], np.int32)
    elif which == "small":
        c = synthetic.case_small("long", nSteps=n); ids = np.array([0], np.int32)
    else:
        c = synthetic.case_c3(nshots=1, nSteps=n); ids = np.array([0], np.int32)
    lam, mu, rho = c.moduli("true")
    lam0, mu0, rho0 = 0.96 * lam, 0.97 * mu, rho
    paras = {k: c.write_files(tempfile.mkdtemp(prefix=f"3w_{k}_")) for k in ("ref", "orc", "b200")}
    runs = {"ref": op.ref_cufd, "orc": op.oracle_cufd, "b200": b200_cufd}
    g, tr, tm = {}, {}, {}
    for k, f in runs.items():
        t0 = time.time()
        tr[k] = f(2, lam, mu, rho, c.stf, ids, paras[k])["syn"][0]
        g[k] = f(1, lam0, mu0, rho0, c.stf, ids, paras[k])
        tm[k] = time.time() - t0
    inner = interior_mask(c)
    out = {"case": which, "grid": [c.nz_pad, c.nx_pad], "nSteps": n, "seconds": tm}
    for a, b in (("b200", "ref"), ("orc", "ref"), ("b200", "orc")):
        d = {"traces": rel(tr[a][:, 1:], tr[b][:, 1:])}
        for q in ("grad_lambda", "grad_mu", "grad_den"):
            d[q] = [rel(g[a][q], g[b][q]), rel(g[a][q][inner], g[b][q][inner])]
        d["grad_stf"] = rel(g[a]["grad_stf"], g[b]["grad_stf"])
        # early / late halves of grad_stf (the adjoint field has run longest at early times)
        h = n // 2
        d["grad_stf_early_late"] = [rel(g[a]["grad_stf"][:, :h], g[b]["grad_stf"][:, :h]), rel(g[a]["grad_stf"][:, h:], g[b]["grad_stf"][:, h:])]
        out[f"{a}_vs_{b}"] = d
    out["grad_norms_ref"] = {q: float(np.linalg.norm(g["ref"][q][inner])) for q in ("grad_lambda", "grad_mu", "grad_den")}
    out["stf_grad_absmax_ref_quarters"] = [float(np.abs(g["ref"]["grad_stf"][0, i * n // 4:(i + 1) * n // 4]).max()) for i in range(4)]
    print(json.dumps(out), flush=True)
    ops.release()
