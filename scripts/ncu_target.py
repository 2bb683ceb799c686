"""Short C2-shaped gradient run (30 shots, few time steps) for ncu captures.
   ncu ... python scripts/ncu_target.py [nsteps] [nshots]"""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from fwiflow.jl_b200 import ops, synthetic
nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 16
nshots = int(sys.argv[2]) if len(sys.argv) > 2 else 30
case = sys.argv[3] if len(sys.argv) > 3 else "c2"
c = synthetic.case_c2(nshots=nshots, nSteps=nsteps) if case == "c2" else synthetic.case_c3(nshots=nshots, nSteps=nsteps)
para = c.write_files(tempfile.mkdtemp(prefix="ncu_"))
ids = np.arange(nshots, dtype=np.int32)
p = ops.Plan(para, ids)
p.set_stf(c.stf); p.set_model(*c.moduli("true")); p.run(2); p.write_obs_files()
p.set_model(*c.moduli("init")); p.load_obs_files()
p.run(1); p.run(1)
print("done", p.launch_count())
