"""Debug helper: small C2-shaped run.  python scripts/dbg_run.py nshots nsteps calc_ids..."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from fwiflow.jl_b200 import ops, synthetic
nshots, nsteps = int(sys.argv[1]), int(sys.argv[2])
c = synthetic.case_c2(nshots=nshots, nSteps=nsteps)
para = c.write_files(tempfile.mkdtemp(prefix="dbg_"))
ids = np.arange(nshots, dtype=np.int32)
p = ops.Plan(para, ids)
p.set_stf(c.stf); p.set_model(*c.moduli("true"))
p.run(2); print("obs ok", flush=True)
p.write_obs_files(); p.set_model(*c.moduli("init")); p.load_obs_files()
p.run(1); print("grad ok", flush=True)
