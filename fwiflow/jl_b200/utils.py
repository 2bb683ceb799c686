"""Host-side conventions of the FWI path: parameter / survey files, Ricker source,
velocity <-> moduli.  Python mirror of the Julia helpers of the reference
(`src/Utils.jl`), same names, same argument meaning, same on-disk formats:

* ``paraGen``            -- src/Utils.jl:21-63   (single-line para JSON)
* ``surveyGen``          -- src/Utils.jl:80-109  (single-line survey JSON)
* ``sourceGene``         -- src/Utils.jl:116-132 (integrated Ricker wavelet)
* ``velocity_to_moduli`` -- src/Utils.jl:221-227 (lambda, mu in MPa)

The JSON files are what `fwi_op` / `fwi_obs_op` consume through `para_fname`
(reference parser: deps/CustomOps/FWI/Src/Parameter.cpp:41-177,
deps/CustomOps/FWI/Src/Src_Rec.cu:71-115).
"""
from __future__ import annotations

import json
import os
from collections import OrderedDict

import numpy as np

__all__ = [
    "paraGen",
    "surveyGen",
    "sourceGene",
    "velocity_to_moduli",
    "moduli_to_velocity_grads",
    "symmetric_pad",
    "nPad_rule",
]


def _num(v):
    """JSON number with Julia's JSON.json formatting rules (Int stays Int, Real -> float)."""
    if isinstance(v, (bool, np.bool_)):
        return bool(v)
    if isinstance(v, (int, np.integer)):
        return int(v)
    return float(v)


def paraGen(nz, nx, dz, dx, nSteps, dt, f0, nPml, nPad, para_fname, survey_fname,
            data_dir_name, if_win=False, filter_para=None, if_src_update=False,
            scratch_dir_name=""):
    """Write the parameter file (src/Utils.jl:21-63).  `nz`, `nx` are the PADDED sizes."""
    para = OrderedDict()
    para["nz"] = int(nz)
    para["nx"] = int(nx)
    para["dz"] = _num(dz)
    para["dx"] = _num(dx)
    para["nSteps"] = int(nSteps)
    para["dt"] = float(dt)  # the reference parser asserts IsDouble (Parameter.cpp:88)
    para["f0"] = _num(f0)
    para["nPoints_pml"] = int(nPml)
    para["nPad"] = int(nPad)
    if if_win:
        para["if_win"] = True
    if filter_para is not None:
        para["filter"] = [float(v) for v in filter_para]
    if if_src_update:
        para["if_src_update"] = True
    para["survey_fname"] = str(survey_fname)
    para["data_dir_name"] = str(data_dir_name)
    if not os.path.isdir(data_dir_name):
        os.makedirs(data_dir_name, exist_ok=True)
    if scratch_dir_name != "":
        para["scratch_dir_name"] = str(scratch_dir_name)
        if not os.path.isdir(scratch_dir_name):
            os.makedirs(scratch_dir_name, exist_ok=True)
    with open(para_fname, "w", encoding="utf-8") as f:   # ONE line: the reference reads one getline; paths as raw UTF-8
        f.write(json.dumps(para, separators=(",", ":"), ensure_ascii=False))
    return para


def surveyGen(z_src, x_src, z_rec, x_rec, survey_fname, Windows=None, Weights=None):
    """Write the survey file (src/Utils.jl:80-109).  All shots share the receiver list.

    Coordinates are 0-based offsets into the UNPADDED grid; the op adds nPml
    (deps/CustomOps/FWI/Src/Src_Rec.cu:86-113).  Shot keys are ``shot0 .. shot{n-1}``.
    """
    z_src = [int(v) for v in np.asarray(z_src).ravel()]
    x_src = [int(v) for v in np.asarray(x_src).ravel()]
    z_rec = [int(v) for v in np.asarray(z_rec).ravel()]
    x_rec = [int(v) for v in np.asarray(x_rec).ravel()]
    assert len(z_src) == len(x_src) and len(z_rec) == len(x_rec)
    survey = OrderedDict()
    survey["nShots"] = len(x_src)
    for i in range(len(x_src)):
        shot = OrderedDict()
        shot["z_src"] = z_src[i]
        shot["x_src"] = x_src[i]
        shot["nrec"] = len(x_rec)
        shot["z_rec"] = z_rec
        shot["x_rec"] = x_rec
        if Windows is not None:
            shot["win_start"] = list(Windows[f"shot{i}"]["start"])
            shot["win_end"] = list(Windows[f"shot{i}"]["end"])
        if Weights is not None:
            shot["weights"] = list(Weights[f"shot{i}"]["weights"])
        survey[f"shot{i}"] = shot
    with open(survey_fname, "w", encoding="utf-8") as f:
        f.write(json.dumps(survey, separators=(",", ":"), ensure_ascii=False))
    return survey


def sourceGene(f, nStep, delta_t):
    """Integrated Ricker wavelet, shape (1, nStep) float64 (src/Utils.jl:116-132)."""
    e = np.pi * np.pi * f * f
    t_delay = 1.2 / f
    t = delta_t * np.arange(nStep, dtype=np.float64) - t_delay
    source = (1 - 2 * e * t ** 2) * np.exp(-e * t ** 2)
    # Julia accumulates sequentially source[it] += source[it-1]; cumsum has the same order
    source = np.cumsum(source)
    return (source * delta_t).reshape(1, nStep)


def velocity_to_moduli(cp, cs, den):
    """lambda = (cp^2 - 2 cs^2) rho / 1e6, mu = cs^2 rho / 1e6 [MPa] (src/Utils.jl:221-227)."""
    cp = np.asarray(cp, dtype=np.float64)
    cs = np.asarray(cs, dtype=np.float64)
    den = np.asarray(den, dtype=np.float64)
    lam = (cp * cp - 2.0 * cs * cs) * den / 1e6
    mu = cs * cs * den / 1e6
    return lam, mu


def moduli_to_velocity_grads(cp, cs, den, g_lam, g_mu, g_den):
    """Chain rule of `velocity_to_moduli` (what TF autodiff does in the reference,
    src/FWI.jl:178 + src/Utils.jl:224-225): gradients w.r.t. (cp, cs, rho)."""
    cp = np.asarray(cp, dtype=np.float64)
    cs = np.asarray(cs, dtype=np.float64)
    den = np.asarray(den, dtype=np.float64)
    g_cp = 2.0 * cp * den / 1e6 * g_lam
    g_cs = (-4.0 * g_lam + 2.0 * g_mu) * cs * den / 1e6
    g_rho = g_den + ((cp * cp - 2.0 * cs * cs) * g_lam + cs * cs * g_mu) / 1e6
    return g_cp, g_cs, g_rho


def nPad_rule(nz, nPml=32):
    """nPad = 32 - mod(nz + 2 nPml, 32), in [1, 32] (src/FWI.jl:12)."""
    return 32 - ((nz + 2 * nPml) % 32)


def symmetric_pad(a, nPml, nPad):
    """tf.pad(a, [nPml (nPml+nPad); nPml nPml], "SYMMETRIC") (src/FWI.jl:202)."""
    return np.pad(np.asarray(a, dtype=np.float64), ((nPml, nPml + nPad), (nPml, nPml)), mode="symmetric")


def klauderWave(fmin, fmax, t_sweep, nStepTotal, nStepDelay, delta_t):
    """Klauder wavelet, shape (1, nStepTotal) like the reference's crop (src/Utils.jl:163-183): the symmetric
    autocorrelation of a linear sweep, centre sample 1.0, cropped to start `nStepDelay` samples before the centre."""
    nStep = int(nStepTotal) - int(nStepDelay)
    K = (fmax - fmin) / t_sweep
    f0 = (fmin + fmax) / 2.0
    t = delta_t * np.arange(1, nStep, dtype=np.float64)
    half = np.sin(np.pi * K * t * (t_sweep - t)) * np.cos(2.0 * np.pi * f0 * t) / (np.pi * K * t * t_sweep)
    source = np.empty(2 * nStep - 1, dtype=np.float64)
    source[:nStep - 1] = half[::-1]
    source[nStep - 1] = 1.0
    source[nStep:] = half
    return source[nStep - int(nStepDelay) - 1:].reshape(1, -1)      # Julia 1-based source[:, nStep-nStepDelay:end]


def cs_bounds_cloud(cpImg, Bounds):
    """Upper / lower cs bounds from a (vp, vs_high, vs_low) reference cloud by piecewise-linear interpolation, held
    constant outside the cloud like Dierckx `Spline1D(...; k=1)` with its default "nearest" boundary
    (src/Utils.jl:144-156)."""
    cp = np.asarray(cpImg, dtype=np.float64)
    B = np.asarray(Bounds, dtype=np.float64)
    order = np.argsort(B[0])
    return np.interp(cp, B[0][order], B[1][order]), np.interp(cp, B[0][order], B[2][order])


def resize_bilinear(a, nz, nx):
    """tf.image.resize_bilinear(a, (nz, nx)) with its default align_corners=False / legacy sampling
    (source coordinate = destination index * in/out), as src/Utils.jl:192-200 applies it."""
    a = np.asarray(a, dtype=np.float64)

    def axis_weights(n_in, n_out):
        pos = np.arange(n_out, dtype=np.float64) * (n_in / float(n_out))
        lo = np.minimum(np.floor(pos).astype(np.int64), n_in - 1)
        hi = np.minimum(lo + 1, n_in - 1)
        return lo, hi, pos - lo

    zl, zh, zw = axis_weights(a.shape[0], nz)
    xl, xh, xw = axis_weights(a.shape[1], nx)
    top = a[zl][:, xl] * (1.0 - xw) + a[zl][:, xh] * xw
    bot = a[zh][:, xl] * (1.0 - xw) + a[zh][:, xh] * xw
    return top * (1.0 - zw)[:, None] + bot * zw[:, None]


def padding(cp, cs, den, nz_orig, nx_orig, nz, nx, nPml, nPad):
    """Resample the (nz_orig, nx_orig) model to (nz, nx), then pad symmetrically with the PML and the nPad rows
    (src/Utils.jl:192-214)."""
    out = []
    for a in (cp, cs, den):
        a = np.asarray(a, dtype=np.float64).reshape(nz_orig, nx_orig)
        out.append(symmetric_pad(resize_bilinear(a, nz, nx), nPml, nPad))
    return tuple(out)
