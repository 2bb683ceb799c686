"""BASELINE configs[0] (C1: 100x100 homogeneous model, 1 shot, 1000 steps, fwi_obs_op): wall time of the forward
modelling call through host buffers for libfwi_b200, the reference op and the CPU oracle.  TEST INFRASTRUCTURE."""
import json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import b200_cufd
from oracle import oracle_py as op
from fwiflow.jl_b200 import synthetic
c = synthetic.case_c1(nSteps=1000)
ids = np.array([0], np.int32)
lam, mu, rho = c.moduli("true")
out = {"config": "c1", "grid": [c.nz_pad, c.nx_pad], "nSteps": 1000}
for who, run in (("b200", b200_cufd), ("ref", op.ref_cufd), ("oracle_cpu", op.oracle_cufd)):
    para = c.write_files(tempfile.mkdtemp(prefix=f"c1_{who}_"))
    run(2, lam, mu, rho, c.stf, ids, para)
    t = []
    for _ in range(5):
        t0 = time.perf_counter(); run(2, lam, mu, rho, c.stf, ids, para); t.append(time.perf_counter() - t0)
    out[who + "_s"] = float(np.median(t))
out["cell_updates_per_s"] = {k[:-2]: c.nz_pad * c.nx_pad * 999 / out[k] for k in ("b200_s", "ref_s", "oracle_cpu_s")}
out["host_cores"] = os.cpu_count()
print(json.dumps(out))
