"""Generate golden vectors for the FWI hot path FROM THE REFERENCE ITSELF.

Runs on a GPU box (the reference has no CPU path):
    python tests/golden/make_golden.py [outdir]
It calls the unmodified reference `cufd()` (oracle/_ref/libCUFD_ref.so, built by
oracle/build_ref.sh from /root/reference) on the seeded synthetic cases of
fwiflow.jl_b200.synthetic and stores traces / misfit / gradients as small .npz
files.  The committed copies under tests/golden/ pin the CPU oracle
(tests/test_oracle_golden.py) and the CUDA path (tests/test_parity_gpu.py).
It also prints the oracle-vs-reference deviations measured on the same box.
"""
from __future__ import annotations

import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from fwiflow.jl_b200 import synthetic as syn  # noqa: E402
from oracle import oracle_py as op  # noqa: E402


def golden_cases():
    return {
        "c1": syn.case_c1(nSteps=1000),
        "small_elastic": syn.case_small("small_elastic", elastic=True),
        "small_acoustic": syn.case_small("small_acoustic", elastic=False, seed=7),
        "gradtest": syn.case_gradtest_small(n=112, nSteps=500),
        "small_windows": syn.case_small_windows(),
        "aniso": syn.case_aniso(),
    }


def rel(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    d = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / d) if d > 0 else float(np.linalg.norm(a))


def run_case(name, c, runner, workdir):
    """obs at the true model (calc 2); misfit (calc 0) and gradient (calc 1) at the initial model."""
    para = c.write_files(workdir)
    ids = np.arange(c.nShots, dtype=np.int32)
    lam, mu, rho = c.moduli("true")
    lam0, mu0, rho0 = c.moduli("init")
    out = {}
    o = runner(2, lam, mu, rho, c.stf, ids, para)
    out["obs"] = np.stack(o["syn"]).astype(np.float32)
    if name != "c1":
        out["misfit_true"] = np.float64(runner(0, lam, mu, rho, c.stf, ids, para)["misfit"])
        out["misfit_init"] = np.float64(runner(0, lam0, mu0, rho0, c.stf, ids, para)["misfit"])
        g = runner(1, lam0, mu0, rho0, c.stf, ids, para)
        for k in ("grad_lambda", "grad_mu", "grad_den", "grad_stf"):
            out[k] = g[k].astype(np.float32)
    return out


def main():
    outdir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(outdir, exist_ok=True)
    assert op.ref_available(), "oracle/_ref/libCUFD_ref.so missing: run oracle/build_ref.sh where /root/reference exists"
    for name, c in golden_cases().items():
        t0 = time.time()
        ref = run_case(name, c, op.ref_cufd, tempfile.mkdtemp(prefix=f"ref_{name}_"))
        t1 = time.time()
        np.savez_compressed(os.path.join(outdir, f"{name}.npz"), **ref)
        # second run of the reference: its own run-to-run noise (atomics)
        ref2 = run_case(name, c, op.ref_cufd, tempfile.mkdtemp(prefix=f"ref2_{name}_"))
        orc = run_case(name, c, op.oracle_cufd, tempfile.mkdtemp(prefix=f"orc_{name}_"))
        t2 = time.time()
        print(f"[{name}] nz_pad={c.nz_pad} nx_pad={c.nx_pad} shots={c.nShots} nrec={c.nrec} steps={c.nSteps} "
              f"ref {t1 - t0:.2f}s oracle {t2 - t1:.2f}s")
        for k in ref:
            if np.ndim(ref[k]) == 0:
                print(f"    {k:12s} ref={float(ref[k]):.9g} ref2={float(ref2[k]):.9g} oracle={float(orc[k]):.9g}")
            else:
                a, b = ref[k], orc[k]
                if k == "obs":  # sample 0 is never written by the reference
                    print(f"    obs[t=0] ref absmax={np.abs(a[..., 0]).max():.3g}")
                    a, b = a[..., 1:], b[..., 1:]
                    r2 = ref2[k][..., 1:]
                else:
                    r2 = ref2[k]
                print(f"    {k:12s} relL2(oracle,ref)={rel(b, a):.3e}  relL2(ref2,ref)={rel(r2, a):.3e}  "
                      f"absmax={np.abs(a).max():.4g}")
    print("golden written to", outdir)


if __name__ == "__main__":
    main()
