"""Time the hot kernels of a plan with CUDA events (C-ABI fwi_b200_plan_time_kernel).
   python scripts/kernel_times.py [c2|c3|c5] [nshots] [iters]     (FWI_ACC=k: shots per accumulator slot)"""
import os, sys, tempfile, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from fwiflow.jl_b200 import ops, synthetic
case = sys.argv[1] if len(sys.argv) > 1 else "c2"
nshots = int(sys.argv[2]) if len(sys.argv) > 2 else 30
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 200
c = {"c2": synthetic.case_c2, "c3": synthetic.case_c3, "c5": synthetic.case_c5}[case](nshots=nshots, nSteps=64)
para = c.write_files(tempfile.mkdtemp(prefix="kt_"))
ids = np.arange(nshots, dtype=np.int32)
if "FWI_DYN" in os.environ:         # reverse step with shot groups: units claimed dynamically (1) or dealt round-robin (0)
    ops.set_option("dyn_units", int(os.environ["FWI_DYN"]))
if "FWI_ACC" in os.environ:      # shots per accumulator slot of the reverse step (0 automatic, 1 a slot per shot)
    ops.set_option("acc_group", int(os.environ["FWI_ACC"]))
p = ops.Plan(para, ids)
p.set_stf(c.stf); p.set_model(*c.moduli("true")); p.run(2); print('obs ok', flush=True); p.write_obs_files()
p.set_model(*c.moduli("init")); p.load_obs_files(); p.run(1); print('grad ok', flush=True)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
names = {0: "fwd", 1: "fwd+save", 2: "rev_image", 3: "adj", 4: "bwd_merged"}
if "FWI_WHICH" in os.environ:
    names = {int(k): names[int(k)] for k in os.environ["FWI_WHICH"].split(",")}
for w in names:
    try:
        ms, b = p.time_kernel(w, iters=iters)
    except Exception as e:
        print(names[w], "n/a", e); continue
    print(f"{case} shots={nshots} batch={p.batch} {names[w]:10s} {ms*1e3:8.1f} us  alg {b/1e6:8.1f} MB  {b/ms/1e6:7.0f} GB/s  frac {b/ms/1e6/peak:.3f}")
