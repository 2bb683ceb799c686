"""CPU: host-side conventions and the C ABI surface (no compute call needs a GPU here)."""
import ctypes
import json
import os
import re
import tempfile

import numpy as np
import pytest

from fwiflow.jl_b200 import _lib, ops, synthetic, utils
from fwiflow.jl_b200.fwi import FWI, FWIExample
from helpers import ROOT


def test_library_exports_every_declared_symbol(lib_built):
    hdr = open(os.path.join(ROOT, "include", "fwi_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(fwi_b200_[a-z_]+)\s*\(", hdr)))
    assert declared, "no declarations found"
    assert sorted(_lib.SYMBOLS) == declared
    L = ctypes.CDLL(lib_built)
    for s in declared:
        assert hasattr(L, s), s
    assert b"sm_100a" in _lib.lib().fwi_b200_version()


def test_paragen_surveygen_roundtrip():
    wd = tempfile.mkdtemp()
    para = os.path.join(wd, "para_file.json")
    survey = os.path.join(wd, "survey_file.json")
    utils.paraGen(224, 448, 24.0, 24.0, 2000, 0.0025, 4.5, 32, 26, para, survey, os.path.join(wd, "Data"))
    utils.surveyGen([2, 2], [4, 12], [2, 2, 2], [3, 4, 5], survey)
    text = open(para).read()
    assert "\n" not in text                      # the reference parser reads ONE line (Parameter.cpp:28)
    p = json.loads(text)
    assert list(p)[:9] == ["nz", "nx", "dz", "dx", "nSteps", "dt", "f0", "nPoints_pml", "nPad"]
    assert p["nz"] == 224 and p["survey_fname"] == survey and os.path.isdir(p["data_dir_name"])
    s = json.loads(open(survey).read())
    assert s["nShots"] == 2 and s["shot1"]["x_src"] == 12 and s["shot0"]["nrec"] == 3
    assert s["shot1"]["x_rec"] == [3, 4, 5]


def test_sourcegene_matches_julia_definition():
    f, n, dt = 4.5, 2000, 0.0025
    s = utils.sourceGene(f, n, dt)
    e = np.pi ** 2 * f ** 2
    src = np.zeros(n)
    for it in range(n):
        t = dt * it - 1.2 / f
        src[it] = (1 - 2 * e * t * t) * np.exp(-e * t * t)
    for it in range(1, n):
        src[it] += src[it - 1]
    assert s.shape == (1, n)
    np.testing.assert_allclose(s[0], src * dt, rtol=1e-12, atol=1e-18)


def test_reference_source_fixture_is_this_wavelet_family():
    """docs/data/sourceF_4p5_2_high.bin peaks at sample 175 with 0.0287;
    sourceGene(4.5, 2000, 0.0025) is the same integrated Ricker up to the fixture's extra filtering."""
    s = utils.sourceGene(4.5, 2000, 0.0025)[0]
    assert 100 < int(np.argmax(s)) < 140 and 0.02 < s.max() < 0.06


def test_velocity_moduli_chain_rule():
    rng = np.random.default_rng(0)
    cp = 3000 + 100 * rng.random((5, 6)); cs = cp / 1.8; rho = 2000 + 50 * rng.random((5, 6))
    gl, gm, gd = rng.random((5, 6)), rng.random((5, 6)), rng.random((5, 6))
    g_cp, g_cs, g_rho = utils.moduli_to_velocity_grads(cp, cs, rho, gl, gm, gd)
    f = lambda cp, cs, rho: sum(np.sum(g * v) for g, v in zip((gl, gm), utils.velocity_to_moduli(cp, cs, rho))) + np.sum(gd * rho)
    h = 1e-3
    for arr, g in ((cp, g_cp), (cs, g_cs), (rho, g_rho)):
        d = np.zeros_like(arr); d[2, 3] = h
        args = [cp, cs, rho]
        idx = [i for i, a in enumerate(args) if a is arr][0]
        up = list(args); up[idx] = arr + d
        dn = list(args); dn[idx] = arr - d
        assert (f(*up) - f(*dn)) / (2 * h) == pytest.approx(g[2, 3], rel=1e-6)


def test_fwi_struct_matches_reference_conventions():
    fwi = FWIExample()
    assert (fwi.nPad, fwi.nz_pad, fwi.nx_pad) == (26, 224, 448)         # src/FWI.jl:12-14
    assert len(fwi.ind_src_x) == 48 and len(fwi.ind_rec_x) == 379        # src/FWI.jl:87-91
    assert fwi.mask.shape == (224, 448) and fwi.mask[32:42].sum() == 0 and fwi.mask[42, 32] == 1
    p = json.loads(open(fwi.para_path).read())
    assert (p["nz"], p["nx"], p["nPad"]) == (224, 448, 26)
    assert utils.nPad_rule(100) == 28 and utils.nPad_rule(64) == 32


def test_symmetric_padding_is_tf_symmetric():
    a = np.arange(12.0).reshape(3, 4)
    p = utils.symmetric_pad(a, 2, 1)
    assert p.shape == (3 + 2 + 3, 4 + 4)
    assert p[1, 2] == a[0, 0] and p[0, 2] == a[1, 0] and p[2 + 3, 2] == a[2, 0] and p[2 + 3 + 2, 2] == a[0, 0]


def _case_files():
    c = synthetic.case_small("host", nz=40, nx=56, nSteps=50, nshots=2)
    wd = tempfile.mkdtemp()
    return c, wd, c.write_files(wd)


def test_unsupported_keys_are_rejected_not_ignored(lib_built):
    c, wd, para = _case_files()
    lam, mu, rho = c.moduli("true")
    p = json.loads(open(para).read())
    for extra in ({"filter": [0.0, 0.1, 100.0, 200.0]}, {"if_src_update": True}):
        q = dict(p); q.update(extra)
        fn = os.path.join(wd, "para_bad.json")
        open(fn, "w").write(json.dumps(q))
        with pytest.raises(ops.FwiError) as ei:
            ops.fwi_obs_op(lam, mu, rho, c.stf, 0, [0], fn)
        assert ei.value.code == -6
    # per-trace windows ARE supported: if_win=true then needs win_start / win_end / weights in the survey file
    q = dict(p); q.update({"if_win": True})
    fn = os.path.join(wd, "para_win.json")
    open(fn, "w").write(json.dumps(q))
    with pytest.raises(ops.FwiError) as ei:
        ops.fwi_obs_op(lam, mu, rho, c.stf, 0, [0], fn)
    assert ei.value.code == -3 and "win_start" in str(ei.value)
    q = dict(p); q.update({"if_win": False, "if_src_update": False, "isAc": True})   # the reference's sample file
    fn = os.path.join(wd, "para_ok.json")
    open(fn, "w").write(json.dumps(q, indent=1))      # multi-line is tolerated
    with pytest.raises(ops.FwiError) as ei:
        ops.fwi_obs_op(lam, mu, rho, c.stf, 0, [0], fn)
    assert ei.value.code == -5                        # parsed fine; fails only because this box has no GPU


def test_error_codes_instead_of_exit(lib_built):
    c, wd, para = _case_files()
    lam, mu, rho = c.moduli("true")
    with pytest.raises(ops.FwiError) as ei:
        ops.fwi_op(lam, mu, rho, c.stf, 0, [0], os.path.join(wd, "nope.json"))
    assert ei.value.code == -2
    open(os.path.join(wd, "broken.json"), "w").write('{"nz": 12, ')
    with pytest.raises(ops.FwiError) as ei:
        ops.fwi_op(lam, mu, rho, c.stf, 0, [0], os.path.join(wd, "broken.json"))
    assert ei.value.code == -3
    p = json.loads(open(para).read()); del p["nSteps"]
    open(os.path.join(wd, "missing.json"), "w").write(json.dumps(p))
    with pytest.raises(ops.FwiError) as ei:
        ops.fwi_op(lam, mu, rho, c.stf, 0, [0], os.path.join(wd, "missing.json"))
    assert ei.value.code == -3
    with pytest.raises(ops.FwiError):
        ops.fwi_op(lam, mu, rho, c.stf, 0, [7], para)   # shot7 is not in the survey


def test_no_cpu_fallback(lib_built):
    """Without a CUDA device the product path must fail loudly (never route through the oracle)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    c, wd, para = _case_files()
    lam, mu, rho = c.moduli("true")
    with pytest.raises(ops.FwiError) as ei:
        ops.fwi_obs_op(lam, mu, rho, c.stf, 0, [0, 1], para)
    assert ei.value.code == -5 and "no CPU fallback" in str(ei.value)
    src = ""
    for root, _, files in os.walk(os.path.join(ROOT, "fwiflow")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".hpp", ".cuh", ".h")):
                src += open(os.path.join(root, f)).read()
    assert "oracle_py" not in src and "fwi_oracle" not in src and "libCUFD_ref" not in src


def test_synthetic_cases_have_baseline_shapes():
    c2 = synthetic.case_c2(nSteps=10)
    assert (c2.nz_pad, c2.nx_pad, c2.nShots, c2.nrec) == (224, 448, 30, 379)
    c1 = synthetic.case_c1(nSteps=10)
    assert (c1.nz_pad, c1.nx_pad, c1.nShots, c1.nrec) == (192, 164, 1, 94)


def test_klauder_bounds_and_resampled_padding():
    """Host helpers of src/Utils.jl that sit beside the op: klauderWave, cs_bounds_cloud, padding with resampling."""
    from fwiflow.jl_b200 import utils
    w = utils.klauderWave(2.0, 20.0, 4.0, 500, 100, 0.002)
    n = 500 - 100
    assert w.shape == (1, n + 100) and w[0, 100] == 1.0                      # centre sample after the delay
    assert np.allclose(w[0, 100 - 50:100], w[0, 101:151][::-1])              # symmetric about the centre
    t = 0.002 * 3
    K, f0 = (20.0 - 2.0) / 4.0, 11.0
    assert w[0, 103] == pytest.approx(np.sin(np.pi * K * t * (4.0 - t)) * np.cos(2 * np.pi * f0 * t) / (np.pi * K * t * 4.0))
    hi, lo = utils.cs_bounds_cloud(np.array([[1500.0, 2500.0], [3500.0, 9000.0]]),
                                   np.array([[2000.0, 3000.0, 4000.0], [1200.0, 1800.0, 2400.0], [900.0, 1300.0, 1700.0]]))
    assert hi[0, 1] == pytest.approx(1500.0) and lo[0, 1] == pytest.approx(1100.0)
    assert hi[0, 0] == 1200.0 and lo[1, 1] == 1700.0                         # held constant outside the cloud
    a = np.arange(12.0).reshape(3, 4)
    assert np.array_equal(utils.resize_bilinear(a, 3, 4), a)
    up = utils.resize_bilinear(a, 6, 8)
    assert up.shape == (6, 8) and up[0, 0] == 0.0 and up[2, 2] == pytest.approx(a[1, 1]) and up[1, 0] == pytest.approx(2.0)
    cp, cs, den = utils.padding(a, a / 2, a + 1, 3, 4, 6, 8, 4, 2)
    assert cp.shape == (6 + 2 * 4 + 2, 8 + 2 * 4) and np.array_equal(cp[4:10, 4:12], up)
    assert np.array_equal(cp[3, 4:12], up[0]) and np.array_equal(cp[2, 4:12], up[1])   # SYMMETRIC (edge repeated)


def test_julia_binding_matches_the_header():
    """julia/FwiB200.jl cannot be executed here (no Julia in the image): check statically that every `ccall` in it names
    a function the header declares, with as many argument types as the C prototype has parameters, doubles / ints /
    strings in the same positions."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "fwi_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = {}
    for m in re.finditer(r"(?:int|void|const char \*|float \*|size_t|long long)\s*\*?\s*(fwi_b200_\w+)\s*\(([^)]*)\)\s*;", hdr):
        args = [a.strip() for a in m.group(2).split(",") if a.strip() and a.strip() != "void"]
        kinds = []
        for a in args:
            if "char" in a:
                kinds.append("str")
            elif "double" in a:
                kinds.append("f64p")
            elif "*" in a and "int" in a:
                kinds.append("i32p")
            elif "*" in a:
                kinds.append("ptr")
            else:
                kinds.append("int")
        protos[m.group(1)] = kinds
    assert "fwi_b200_backward" in protos and len(protos["fwi_b200_backward"]) == 12
    jl = open(os.path.join(root, "julia", "FwiB200.jl")).read()
    calls = re.findall(r"ccall\(\(:(fwi_b200_\w+), LIBFWI\), (\w+),\s*\(([^)]*)\)", jl)
    assert len(calls) >= 6
    jmap = {"Ref{Cdouble}": "f64p", "Ptr{Cdouble}": "f64p", "Cint": "int", "Ptr{Cint}": "i32p", "Cstring": "str",
            "Ptr{Cvoid}": "ptr", "Ref{Ptr{Cvoid}}": "ptr"}
    assert {"fwi_b200_plan_create", "fwi_b200_plan_set_velocities", "fwi_b200_plan_get_velocity_gradients",
            "fwi_b200_plan_run"} <= {c[0] for c in calls}
    for name, ret, types in calls:
        assert name in protos, name
        jt = [t.strip() for t in types.split(",") if t.strip()]
        assert [jmap[t] for t in jt] == protos[name], (name, jt, protos[name])



def test_device_layout_invariants(lib_built):
    """fwi_b200_grid_info (host-only): tiles cover every stored row, the row offset keeps quads aligned and never costs
    a row of tiles, the never-stored rows are exactly the inactive nPad rows, and the quad-granular boundary frames
    hold every cell of the reference's 5-deep ring (Boundary.cu:17-27) at no more than 1.6x its size."""
    from fwiflow.jl_b200.utils import paraGen, nPad_rule
    rng = np.random.default_rng(0)
    sizes = [(100, 100, 32), (134, 384, 32), (1000, 3000, 32), (4000, 8000, 32), (44, 66, 20)]
    sizes += [(int(rng.integers(20, 400)), int(rng.integers(20, 500)), int(rng.choice([12, 20, 32]))) for _ in range(40)]
    for nz0, nx0, nPml in sizes:
        nPad = nPad_rule(nz0, nPml)
        nz, nx = nz0 + 2 * nPml + nPad, nx0 + 2 * nPml
        wd = tempfile.mkdtemp()
        para = os.path.join(wd, "p.json")
        paraGen(nz, nx, 10.0, 10.0, 100, 0.001, 5.0, nPml, nPad, para, os.path.join(wd, "s.json"), os.path.join(wd, "D"))
        ops.set_option("frame_ring", 5)          # the reference's ring depth; the default is the thin ring, below
        try:
            g = ops.grid_info(para)
        finally:
            ops.set_option("frame_ring", 2)
        assert (g["nz"], g["nx"]) == (nz, nx) and g["pitch"] % 32 == 0 and g["pitch"] >= nz
        az_hi = nz - nPad - 3
        assert g["zlive"] % 4 == 0 and az_hi < g["zlive"] <= min(nz, az_hi + 4)
        assert g["z_off"] <= 0 and g["z_off"] % 4 == 0 and g["z_off"] > -56
        assert g["tiles_z"] * 56 + g["z_off"] >= g["zlive"] and g["tiles_z"] == -(-(g["zlive"] - g["z_off"]) // 56)
        assert g["tiles_z"] <= -(-g["zlive"] // 56) and g["tiles_x"] == -(-nx // 28)
        assert (g["zlo"], g["zhi"], g["xlo"], g["xhi"]) == (nPml, nz - nPad - 1 - nPml, nPml, nx - 1 - nPml)
        len_bnd = 10 * ((nz - 2 * nPml - nPad + 4) + (nx - 2 * nPml + 4))          # Boundary.cu:19-23
        assert len_bnd <= g["frame_len"] <= 1.6 * len_bnd + 64, (nz0, nx0, nPml, len_bnd, g["frame_len"])
        # the thin ring (the two cells outside the box only; still whole quads) is at most 0.8x the 5-deep one -- half of
        # it where the two ring rows fall into one quad (C2 / C3 / C5: 0.49x)
        thin = ops.grid_info(para)["frame_len"]
        assert 0.2 * len_bnd <= thin <= 0.8 * g["frame_len"], (nz0, nx0, nPml, len_bnd, thin)
    c2 = ops.grid_info(os.path.join(_case_c2_para()))
    assert (c2["tiles_z"], c2["tiles_x"], c2["zlive"]) == (4, 16, 196)


def _case_c2_para():
    c = synthetic.case_c2(nshots=1, nSteps=10)
    return c.write_files(tempfile.mkdtemp())


def test_timelapse_driver_spreads_surveys_over_gpus(monkeypatch):
    """timelapse_misfit_and_gradients (flow-coupled FWI driver): survey i runs on gpu_ids[i % n], results come back in
    survey order, misfits are summed, and an error in one survey is raised in the caller.  The op is mocked: host logic
    only (the real thing is tests/test_configs_gpu.py::test_c4_timelapse_six_surveys)."""
    from fwiflow.jl_b200 import fwi as F
    calls = []

    def fake(fwi, cp, cs, rho, stf, shot_ids=None, gpu_id=0, **kw):
        calls.append((fwi, gpu_id, kw.get("is_masked")))
        if fwi == "boom":
            raise RuntimeError("survey failed")
        return float(fwi), np.full((2, 2), fwi), np.zeros((2, 2)), np.zeros((2, 2))

    monkeypatch.setattr(F, "compute_misfit_and_gradient", fake)
    surveys = [(k, None, None, None) for k in range(6)]
    total, per = F.timelapse_misfit_and_gradients(surveys, None, gpu_ids=(3, 5), is_masked=True)
    assert total == 15.0 and [p[0] for p in per] == [0.0, 1.0, 2.0, 3.0, 4.0, 5.0]
    assert sorted(calls) == sorted([(k, (3, 5)[k % 2], True) for k in range(6)])
    with pytest.raises(RuntimeError, match="survey failed"):
        F.timelapse_misfit_and_gradients([(1, None, None, None), ("boom", None, None, None)], None, gpu_ids=(0,))


def test_array_shapes_are_checked_against_the_para_file(lib_built):
    """The C ABI carries no sizes (like the reference's cufd): the binding refuses arrays whose shapes do not match the
    parameter file instead of letting the library read or write past them."""
    c, wd, para = _case_files()
    lam, mu, rho = c.moduli("true")
    info = ops.para_info(para)
    assert (info["nz"], info["nx"], info["nSteps"]) == (c.nz_pad, c.nx_pad, c.nSteps)
    P = c.nPml
    unpadded = lam[P:P + c.nz, P:P + c.nx]
    for call in (ops.fwi_op, ops.fwi_obs_op, ops.fwi_op_grad, ops.fwi_op_and_grad):
        with pytest.raises(ops.FwiError, match="parameter file says"):
            call(unpadded, unpadded, unpadded, c.stf, 0, [0], para)              # unpadded (nz, nx) model
        with pytest.raises(ops.FwiError, match="nSteps"):
            call(lam, mu, rho, c.stf[:, :-1], 0, [0], para)                       # wrong record length
        with pytest.raises(ops.FwiError, match="rows of stf"):
            call(lam, mu, rho, c.stf[0], 0, [1], para)                            # 1-D stf = one row, shot id 1
        with pytest.raises(ops.FwiError, match="empty"):
            call(lam, mu, rho, c.stf, 0, [], para)
    with pytest.raises(ops.FwiError):
        ops.fwi_op_and_grad_multi(lam[:-1], mu[:-1], rho[:-1], c.stf, [0, 1], [0, 1], para)
    with pytest.raises(ops.FwiError):
        ops.timelapse([(para, lam, mu, rho), (para, lam.T, mu.T, rho.T)], c.stf, [0], [0, 1])


def test_json_unicode_paths(lib_built):
    """Paths with non-ASCII characters: written raw (UTF-8) by paraGen, and decoded from \\uXXXX escapes -- including a
    surrogate pair -- when another writer (json.dumps' default, rapidjson) escaped them."""
    import torch
    from fwiflow.jl_b200 import synthetic
    c = synthetic.case_small("u", nz=40, nx=48, nSteps=50, nshots=1)
    base = tempfile.mkdtemp(prefix="fwi_unicode_")
    wd = os.path.join(base, "données_é_\U0001F30A")
    para = c.write_files(wd)
    raw = open(para, encoding="utf-8").read()
    assert "données" in raw                                    # written as UTF-8, one line
    esc = os.path.join(base, "para_escaped.json")
    open(esc, "w").write(json.dumps(json.loads(raw)))          # ensure_ascii=True: é, 🌊
    assert "\\u00e9" in open(esc).read() and "\\ud83c\\udf0a" in open(esc).read()
    lam, mu, rho = c.moduli("true")
    for fn in (para, esc):
        assert ops.para_info(fn)["nz"] == c.nz_pad
        if torch.cuda.is_available():
            ops.fwi_obs_op(lam, mu, rho, c.stf, 0, [0], fn)
            assert os.path.exists(os.path.join(wd, "Data", "Shot0.bin"))
        else:
            with pytest.raises(ops.FwiError) as ei:
                ops.fwi_obs_op(lam, mu, rho, c.stf, 0, [0], fn)
            assert ei.value.code == -5, str(ei.value)          # the survey file WAS found (else -2): only the GPU is missing
    bad = os.path.join(base, "para_lone.json")
    open(bad, "w").write(raw.replace("données", "donn\\ud83c"))
    with pytest.raises(ops.FwiError) as ei:
        ops.para_info(bad)
    assert ei.value.code == -3


def test_bench_reads_the_committed_ncu_traffic_file():
    """bench.py multiplies the entries of profiles/traffic.json; notes and shot counts in that file must not reach it."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    t = bench.ncu_traffic()
    assert set(t) >= {"c2", "c3"}
    for cfg in ("c2", "c3"):
        assert {"fwd_step_kernel<save_frames>", "adj_step_kernel", "rev_image_kernel"} <= set(t[cfg])
        assert all(isinstance(v, float) and v > 0 for v in t[cfg].values())
    # whole-gradient byte count of the C2 bench step: 60 + 32 f, 60 + 64 f per cell, 64 per box cell, per time index
    c = synthetic.case_c2(nshots=2, nSteps=8)
    b = bench.whole_gradient_alg_bytes(c, 30, 2000)
    assert 1.1e12 < b < 1.3e12
