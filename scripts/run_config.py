"""One gradient of a BASELINE config on one GPU, timed on the device (CUDA events around plan.run(1)).
   python scripts/run_config.py c3 25 4000        # the per-rank share of C3 (200 shots) at 8 GPUs
   python scripts/run_config.py c2 30 2000
Prints one JSON line (shot-gradients/s, cell-updates/s, batch, frame memory)."""
import json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from fwiflow.jl_b200 import ops, synthetic

case, nshots, nsteps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
if "FWI_RING" in os.environ:        # depth of the saved boundary ring: 5 (reference) or 2 (thin)
    ops.set_option("frame_ring", int(os.environ["FWI_RING"]))
if "FWI_DYN" in os.environ:         # reverse step with shot groups: units claimed dynamically (1) or dealt round-robin (0)
    ops.set_option("dyn_units", int(os.environ["FWI_DYN"]))
if "FWI_ACC" in os.environ:         # shots per accumulator slot of the reverse step (0 automatic, 1 a slot per shot)
    ops.set_option("acc_group", int(os.environ["FWI_ACC"]))
if "FWI_MERGED" in os.environ:      # A/B: backward loop as one merged launch per time index (1) or two launches (0)
    ops.set_option("merged_bwd", int(os.environ["FWI_MERGED"]))
mk = {"c2": synthetic.case_c2, "c3": synthetic.case_c3, "c5": synthetic.case_c5}[case]
c = mk(nshots=nshots, nSteps=nsteps)
para = c.write_files(tempfile.mkdtemp(prefix=f"cfg_{case}_"))
ids = np.arange(nshots, dtype=np.int32)
p = ops.Plan(para, ids, max_batch=int(os.environ.get("FWI_BATCH", "0")))   # FWI_BATCH: shots per launch (0: chosen by the plan)
p.set_stf(c.stf); p.set_model(*c.moduli("true"))
t0 = time.time(); p.run(2); t_obs = time.time() - t0
p.write_obs_files(); p.set_model(*c.moduli("init")); p.load_obs_files()
free0, total = torch.cuda.mem_get_info()
s = torch.cuda.Stream()                    # a created stream: handle 0 (the legacy default stream) would mean "the plan's own"
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
p.run(1)                                   # warm-up (allocations)
torch.cuda.synchronize()
t0 = time.perf_counter()
e0.record(s); p.run(1, stream=s.cuda_stream, sync=False); e1.record(s); torch.cuda.synchronize()
wall = time.perf_counter() - t0
ms = e0.elapsed_time(e1)
assert abs(ms * 1e-3 - wall) < 0.05 * wall + 5e-3, (ms, wall)   # the events bracket the work
free1, _ = torch.cuda.mem_get_info()
j, gl, gm, gd, gs = p.result()
cells = c.nz_pad * c.nx_pad
print(json.dumps({"acc_group": os.environ.get("FWI_ACC", "default"), "merged_bwd": os.environ.get("FWI_MERGED", "default"), "frame_ring": os.environ.get("FWI_RING", "default"), "config": case, "grid": [c.nz_pad, c.nx_pad], "shots": nshots, "nSteps": nsteps, "batch": p.batch,
                  "gradient_s": ms / 1e3, "forward_only_s": t_obs, "shot_gradients_per_s": nshots / (ms / 1e3),
                  "cell_updates_per_s": 2.0 * nshots * cells * (nsteps - 1) / (ms / 1e3),
                  "hbm_used_gb": (total - free1) / 1e9, "misfit": j,
                  "grad_finite": bool(np.isfinite(gl).all() and np.isfinite(gm).all() and np.isfinite(gd).all())}))
