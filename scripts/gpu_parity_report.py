"""Run every golden case through libfwi_b200.so and print deviations from the reference's golden vectors
(and from the CPU oracle run on the same box).  Usage (GPU box): python scripts/gpu_parity_report.py"""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import golden_cases, run_case, rel, load_golden, b200_cufd, interior_mask, away_from_sources

for name, c in golden_cases().items():
    g = load_golden(name)
    t0 = time.time()
    mine = run_case(name, c, b200_cufd, tempfile.mkdtemp(prefix=f"b200_{name}_"))
    t1 = time.time()
    print(f"[{name}] {c.nz_pad}x{c.nx_pad} shots={c.nShots} steps={c.nSteps}  b200 {t1-t0:.2f}s")
    far = away_from_sources(c)
    for k in g:
        if np.ndim(g[k]) == 0:
            print(f"    {k:12s} ref={float(g[k]):.9g} b200={float(mine[k]):.9g}")
        elif k == "obs":
            print(f"    obs          relL2={rel(mine[k][..., 1:], g[k][..., 1:]):.3e}  t0 absmax={np.abs(mine[k][..., 0]).max():.3g}")
        elif k == "grad_stf":
            print(f"    {k:12s} relL2={rel(mine[k], g[k]):.3e}")
        else:
            print(f"    {k:12s} relL2={rel(mine[k], g[k]):.3e}  away-from-src={rel(mine[k][far], g[k][far]):.3e}  nan={np.isnan(mine[k]).sum()}")
