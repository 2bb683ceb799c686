"""TEST INFRASTRUCTURE ONLY -- ctypes doors to the CPU oracle (oracle/fwi_oracle.cpp)
and to the rebuilt, unmodified reference CUDA library (oracle/_ref/libCUFD_ref.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product (fwiflow/jl_b200) never does.
"""
from __future__ import annotations

import ctypes
import json
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "libfwi_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libCUFD_ref.so")

_c_dp = ctypes.POINTER(ctypes.c_double)
_c_fp = ctypes.POINTER(ctypes.c_float)
_c_ip = ctypes.POINTER(ctypes.c_int)


def build_oracle(force=False):
    src = os.path.join(HERE, "fwi_oracle.cpp")
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, os.path.join(HERE, "_build", "libfwi_oracle.so")],
                              stdout=subprocess.DEVNULL)
    return ORACLE_SO


_oracle = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        build_oracle()
        _oracle = ctypes.CDLL(ORACLE_SO)
        _oracle.fwi_oracle_cufd.restype = ctypes.c_int
        _oracle.fwi_oracle_courant.restype = ctypes.c_float
    return _oracle


def _dp(a):
    return None if a is None else a.ctypes.data_as(_c_dp)


def _fp(a):
    return None if a is None else a.ctypes.data_as(_c_fp)


def _ip(a):
    return None if a is None else a.ctypes.data_as(_c_ip)


def read_para(para_fname):
    with open(para_fname) as f:
        return json.loads(f.readline())


def read_survey(survey_fname):
    with open(survey_fname) as f:
        return json.loads(f.readline())


def oracle_cufd(calc_id, lam, mu, den, stf, shot_ids, para_fname, obs=None, snap_it=-1, threads=None):
    """Run the CPU oracle with the reference's file conventions.

    lam, mu, den: (nz_pad, nx_pad) float64 row-major, lam/mu in MPa.  stf: (nShotsTotal, nSteps).
    obs: optional list of per-shot (nrec, nSteps) float32 arrays; when None and calc_id in (0, 1) the
    observed data are read from data_dir_name/Shot<id>.bin like the reference (libCUFD.cu:189-192).
    Returns a dict with misfit / grads / traces.
    """
    para = read_para(para_fname)
    survey = read_survey(para["survey_fname"])
    nz, nx, nSteps = para["nz"], para["nx"], para["nSteps"]
    nPml, nPad = para["nPoints_pml"], para["nPad"]
    shot_ids = np.ascontiguousarray(shot_ids, dtype=np.int32)
    G = len(shot_ids)
    zs = np.zeros(G, np.int32)
    xs = np.zeros(G, np.int32)
    rec_off = np.zeros(G + 1, np.int32)
    zr, xr = [], []
    for i, sid in enumerate(shot_ids):
        sh = survey[f"shot{int(sid)}"]
        zs[i] = sh["z_src"] + nPml
        xs[i] = sh["x_src"] + nPml
        zr.append(np.asarray(sh["z_rec"], np.int32) + nPml)
        xr.append(np.asarray(sh["x_rec"], np.int32) + nPml)
        rec_off[i + 1] = rec_off[i] + sh["nrec"]
    zr = np.ascontiguousarray(np.concatenate(zr), np.int32)
    xr = np.ascontiguousarray(np.concatenate(xr), np.int32)
    win = None
    if para.get("if_win", False):   # Src_Rec.cu:157-200: per-receiver windows (s) and trace weights
        win = [np.ascontiguousarray(np.concatenate([np.asarray(survey[f"shot{int(s)}"][k], np.float64) for s in shot_ids]),
                                    np.float32) for k in ("win_start", "win_end", "weights")]
        assert all(w.size == rec_off[-1] for w in win)
    if "filter" in para or para.get("if_src_update", False):
        raise NotImplementedError("oracle: band-pass filter / source update are not restated")
    ntr = int(rec_off[-1]) * nSteps
    obs_in = None
    if calc_id in (0, 1):
        if obs is None:
            obs = [np.fromfile(os.path.join(para["data_dir_name"], f"Shot{int(s)}.bin"), np.float32)
                   for s in shot_ids]
        obs_in = np.ascontiguousarray(np.concatenate([np.asarray(o, np.float32).ravel() for o in obs]))
        assert obs_in.size == ntr
    lam = np.ascontiguousarray(lam, np.float64)
    mu = np.ascontiguousarray(mu, np.float64)
    den = np.ascontiguousarray(den, np.float64)
    stf = np.ascontiguousarray(stf, np.float64)
    misfit = np.zeros(1, np.float64)
    gl = np.zeros((nz, nx), np.float64)
    gm = np.zeros((nz, nx), np.float64)
    gd = np.zeros((nz, nx), np.float64)
    gs = np.zeros((G, nSteps), np.float64)
    syn = np.zeros(ntr, np.float32)
    res = np.zeros(ntr, np.float32)
    snap_f = np.zeros(nz * nx, np.float32) if snap_it >= 0 else None
    snap_b = np.zeros(nz * nx, np.float32) if snap_it >= 0 else None
    if threads is not None:
        os.environ["OMP_NUM_THREADS"] = str(threads)
    args = [
        ctypes.c_int(nz), ctypes.c_int(nx), ctypes.c_int(nPml), ctypes.c_int(nPad), ctypes.c_int(nSteps),
        ctypes.c_float(para["dz"]), ctypes.c_float(para["dx"]), ctypes.c_float(para["dt"]),
        ctypes.c_float(para["f0"]), ctypes.c_int(calc_id), ctypes.c_int(G), _ip(shot_ids),
        _dp(lam), _dp(mu), _dp(den), _dp(stf), _ip(zs), _ip(xs), _ip(rec_off), _ip(zr), _ip(xr),
        _fp(obs_in), _dp(misfit), _dp(gl), _dp(gm), _dp(gd), _dp(gs), _fp(syn), _fp(res), None,
        ctypes.c_int(snap_it), _fp(snap_f), _fp(snap_b)]
    if win is None:
        rc = oracle_lib().fwi_oracle_cufd(*args)
    else:
        rc = oracle_lib().fwi_oracle_cufd_win(*args, _fp(win[0]), _fp(win[1]), _fp(win[2]))
    if rc != 0:
        raise RuntimeError(f"oracle: Courant limit violated (rc={rc})")
    out = {"misfit": float(misfit[0]), "grad_lambda": gl, "grad_mu": gm, "grad_den": gd, "grad_stf": gs}
    traces, residuals = [], []
    for i in range(G):
        a, b = int(rec_off[i]) * nSteps, int(rec_off[i + 1]) * nSteps
        traces.append(syn[a:b].reshape(-1, nSteps))
        residuals.append(res[a:b].reshape(-1, nSteps))
    out["syn"] = traces
    out["res"] = residuals
    if snap_it >= 0:
        out["snap_fwd"] = snap_f.reshape(nx, nz).T  # [z][x]
        out["snap_back"] = snap_b.reshape(nx, nz).T
    if calc_id == 2:  # libCUFD.cu:514-521
        for i, sid in enumerate(shot_ids):
            traces[i].tofile(os.path.join(para["data_dir_name"], f"Shot{int(sid)}.bin"))
    return out


# ------------------------------------------------------------------------------------------------
# the unmodified reference (needs a GPU)
# ------------------------------------------------------------------------------------------------
_ref = None


def ref_available():
    return os.path.exists(REF_SO)


def ref_lib():
    global _ref
    if _ref is None:
        _ref = ctypes.CDLL(REF_SO)
        _ref.ref_cufd.restype = ctypes.c_int
    return _ref


def ref_cufd(calc_id, lam, mu, den, stf, shot_ids, para_fname, gpu_id=0):
    """Call the reference cufd() (deps/CustomOps/FWI/Src/libCUFD.cu:34) through oracle/ref_shim.cu.
    Files are read / written by the reference itself (Data/Shot<id>.bin)."""
    para = read_para(para_fname)
    nz, nx, nSteps = para["nz"], para["nx"], para["nSteps"]
    shot_ids = np.ascontiguousarray(shot_ids, dtype=np.int32)
    G = len(shot_ids)
    lam = np.ascontiguousarray(lam, np.float64)
    mu = np.ascontiguousarray(mu, np.float64)
    den = np.ascontiguousarray(den, np.float64)
    stf = np.ascontiguousarray(stf, np.float64)
    misfit = np.zeros(1, np.float64)
    gl = np.zeros((nz, nx), np.float64)
    gm = np.zeros((nz, nx), np.float64)
    gd = np.zeros((nz, nx), np.float64)
    gs = np.zeros((G, nSteps), np.float64)
    rc = ref_lib().ref_cufd(_dp(misfit), _dp(gl), _dp(gm), _dp(gd), _dp(gs), _dp(lam), _dp(mu), _dp(den),
                            _dp(stf), ctypes.c_int(calc_id), ctypes.c_int(gpu_id), ctypes.c_int(G),
                            _ip(shot_ids), para_fname.encode())
    assert rc == 0
    out = {"misfit": float(misfit[0]), "grad_lambda": gl, "grad_mu": gm, "grad_den": gd, "grad_stf": gs}
    if calc_id == 2:
        out["syn"] = [np.fromfile(os.path.join(para["data_dir_name"], f"Shot{int(s)}.bin"),
                                  np.float32).reshape(-1, nSteps) for s in shot_ids]
    return out
