"""High-level FWI API, Python mirror of the reference's `src/FWI.jl`.

* ``FWI``                  -- struct holding geometry + file locations          (src/FWI.jl:3-25, 38-60)
* ``compute_observation``  -- forward modelling -> (nShots, nSteps, nrec) array  (src/FWI.jl:109-135)
* ``compute_misfit``       -- masked model -> moduli -> fwi_op                   (src/FWI.jl:156-189)
* ``compute_misfit_and_gradient`` -- same, plus gradients w.r.t. (cp, cs, rho) through the chain rule
  that TensorFlow autodiff applies in the reference (velocity_to_moduli + mask blend)
* ``padding`` / ``try_pad`` -- symmetric PML padding                              (src/FWI.jl:193-232)

Shot ids are 1-based here, exactly like the Julia API, and converted to 0-based before the op.
"""
from __future__ import annotations

import os
import tempfile
from dataclasses import dataclass, field

import numpy as np

from . import ops
from .utils import (moduli_to_velocity_grads, paraGen, surveyGen, symmetric_pad, velocity_to_moduli)

__all__ = ["FWI", "FWIExample", "compute_observation", "compute_misfit", "compute_misfit_and_gradient", "padding",
           "try_pad", "timelapse_misfit_and_gradients", "timelapse_misfit_and_gradients_batched"]


@dataclass
class FWI:
    nz: int = 134
    nx: int = 384
    dz: float = 24.0
    dx: float = 24.0
    nSteps: int = 2000
    dt: float = 0.0025
    f0: float = 4.5
    nPml: int = 32
    nPad: int | None = None
    para_fname: str = "para_file.json"
    survey_fname: str = "survey_file.json"
    data_dir_name: str = "Data"
    WORKSPACE: str = field(default_factory=lambda: tempfile.mkdtemp(prefix="fwi_b200_"))
    ind_src_x: np.ndarray | None = None
    ind_src_z: np.ndarray | None = None
    ind_rec_x: np.ndarray | None = None
    ind_rec_z: np.ndarray | None = None
    mask: np.ndarray | None = None
    mask_neg: np.ndarray | None = None

    def __post_init__(self):
        if self.nPad is None:
            self.nPad = 32 - ((self.nz + 2 * self.nPml) % 32)       # src/FWI.jl:12
        self.nz_pad = self.nz + 2 * self.nPml + self.nPad           # src/FWI.jl:13
        self.nx_pad = self.nx + 2 * self.nPml                       # src/FWI.jl:14
        for k in ("ind_src_x", "ind_src_z", "ind_rec_x", "ind_rec_z"):
            v = getattr(self, k)
            if v is None:
                raise ValueError(f"FWI: {k} is required")
            setattr(self, k, np.asarray(v, dtype=np.int64).ravel())
        assert len(self.ind_rec_x) == len(self.ind_rec_z)
        assert len(self.ind_src_x) == len(self.ind_src_z)
        P, nz, nx = self.nPml, self.nz, self.nx
        Mask = np.zeros((self.nz_pad, self.nx_pad))                 # src/FWI.jl:45-49
        Mask[P:P + nz, P:P + nx] = 1.0
        Mask[P:P + 10, :] = 0.0
        self.mask = Mask
        self.mask_neg = 1.0 - Mask
        os.makedirs(self.WORKSPACE, exist_ok=True)
        paraGen(self.nz_pad, self.nx_pad, self.dz, self.dx, self.nSteps, self.dt, self.f0, self.nPml, self.nPad,
                self.para_path, os.path.join(self.WORKSPACE, self.survey_fname),
                os.path.join(self.WORKSPACE, self.data_dir_name))
        surveyGen(self.ind_src_z, self.ind_src_x, self.ind_rec_z, self.ind_rec_x,
                  os.path.join(self.WORKSPACE, self.survey_fname))

    @property
    def para_path(self):
        return os.path.join(self.WORKSPACE, self.para_fname)


def FWIExample(**kw):
    """The Marmousi geometry of the reference (src/FWI.jl:87-97)."""
    ind_src_x = np.arange(4, 385, 8)
    ind_rec_x = np.arange(3, 382)
    return FWI(nz=134, nx=384, dz=24.0, dx=24.0, nSteps=2000, dt=0.0025, ind_src_x=ind_src_x,
               ind_src_z=2 * np.ones_like(ind_src_x), ind_rec_x=ind_rec_x, ind_rec_z=2 * np.ones_like(ind_rec_x), **kw)


def padding(fwi: FWI, *arrays):
    """tf.pad(cp, [nPml (nPml+nPad); nPml nPml], "SYMMETRIC") (src/FWI.jl:193-205).  Inputs must already
    have the (nz, nx) shape (the reference's bilinear resize of mismatching inputs is not reproduced)."""
    out = []
    for a in arrays:
        a = np.asarray(a, dtype=np.float64)
        if a.shape != (fwi.nz, fwi.nx):
            raise ValueError(f"padding: expected shape {(fwi.nz, fwi.nx)}, got {a.shape}")
        out.append(symmetric_pad(a, fwi.nPml, fwi.nPad))
    return out[0] if len(out) == 1 else out


def try_pad(fwi: FWI, *arrays):
    out = []
    for a in arrays:
        a = np.asarray(a, dtype=np.float64)
        out.append(a if a.shape == (fwi.nz_pad, fwi.nx_pad) else padding(fwi, a))
    return out[0] if len(out) == 1 else out


def _stf_rows(fwi, stf_array):
    stf = np.asarray(stf_array, dtype=np.float64)
    if stf.ndim == 1 or (stf.ndim == 2 and stf.shape[0] == 1 and len(fwi.ind_src_z) > 1):
        stf = np.repeat(stf.reshape(1, -1), len(fwi.ind_src_z), axis=0)   # src/FWI.jl:118-120
    return stf


def compute_observation(fwi: FWI, cp, cs, rho, stf_array, shot_ids=None, gpu_id=0):
    """Writes WORKSPACE/Data/Shot<id>.bin and returns data[i, :, :] = (nSteps, nrec) per shot (src/FWI.jl:109-135)."""
    cp_pad, cs_pad, rho_pad = try_pad(fwi, cp, cs, rho)
    stf = _stf_rows(fwi, stf_array)
    lam, mu = velocity_to_moduli(cp_pad, cs_pad, rho_pad)
    if shot_ids is None:
        shot_ids = np.arange(1, len(fwi.ind_src_x) + 1)
    ids0 = np.asarray(shot_ids, dtype=np.int32) - 1
    ops.fwi_obs_op(lam, mu, rho_pad, stf, gpu_id, ids0, fwi.para_path)
    nrec = len(fwi.ind_rec_z)
    data = np.zeros((len(ids0), fwi.nSteps, nrec))
    for i, sid in enumerate(ids0):
        a = np.fromfile(os.path.join(fwi.WORKSPACE, fwi.data_dir_name, f"Shot{int(sid)}.bin"), np.float32)
        data[i] = a.reshape(nrec, fwi.nSteps).T      # Julia reshape (nSteps, nrec) column-major
    return data


def _masked_models(fwi, cp, cs, rho, is_masked, cp_ref, cs_ref, rho_ref):
    cp_pad, cs_pad, rho_pad = try_pad(fwi, cp, cs, rho)
    if is_masked:
        return cp_pad, cs_pad, rho_pad
    if cp_ref is None or cs_ref is None or rho_ref is None:
        raise ValueError("compute_misfit: cp_ref, cs_ref, rho_ref are required when is_masked is False")
    cp_r, cs_r, rho_r = try_pad(fwi, cp_ref, cs_ref, rho_ref)
    return (cp_pad * fwi.mask + cp_r * fwi.mask_neg, cs_pad * fwi.mask + cs_r * fwi.mask_neg,
            rho_pad * fwi.mask + rho_r * fwi.mask_neg)                  # src/FWI.jl:174-176


def compute_misfit(fwi: FWI, cp, cs, rho, stf_array, shot_ids=None, gpu_id=0, is_masked=False, cp_ref=None,
                   cs_ref=None, rho_ref=None):
    """Misfit of the (masked) model against the data in WORKSPACE/Data (src/FWI.jl:156-189)."""
    cp_m, cs_m, rho_m = _masked_models(fwi, cp, cs, rho, is_masked, cp_ref, cs_ref, rho_ref)
    lam, mu = velocity_to_moduli(cp_m, cs_m, rho_m)
    stf = _stf_rows(fwi, stf_array)
    if shot_ids is None:
        shot_ids = np.arange(1, len(fwi.ind_src_x) + 1)
    ids0 = np.asarray(shot_ids, dtype=np.int32) - 1
    return ops.fwi_op(lam, mu, rho_m, stf, gpu_id, ids0, fwi.para_path)


def compute_misfit_and_gradient(fwi: FWI, cp, cs, rho, stf_array, shot_ids=None, gpu_id=0, is_masked=False,
                                cp_ref=None, cs_ref=None, rho_ref=None):
    """(misfit, d/dcp, d/dcs, d/drho) on the padded grid: what `gradients(loss, [cp, cs, rho])` yields in the
    reference (mask blend -> velocity_to_moduli -> fwi_op), from a single propagation pair."""
    cp_m, cs_m, rho_m = _masked_models(fwi, cp, cs, rho, is_masked, cp_ref, cs_ref, rho_ref)
    lam, mu = velocity_to_moduli(cp_m, cs_m, rho_m)
    stf = _stf_rows(fwi, stf_array)
    if shot_ids is None:
        shot_ids = np.arange(1, len(fwi.ind_src_x) + 1)
    ids0 = np.asarray(shot_ids, dtype=np.int32) - 1
    misfit, gl, gm, gd, _ = ops.fwi_op_and_grad(lam, mu, rho_m, stf, gpu_id, ids0, fwi.para_path)
    g_cp, g_cs, g_rho = moduli_to_velocity_grads(cp_m, cs_m, rho_m, gl, gm, gd)
    if not is_masked:
        g_cp, g_cs, g_rho = g_cp * fwi.mask, g_cs * fwi.mask, g_rho * fwi.mask
    return misfit, g_cp, g_cs, g_rho


def compute_misfit_and_gradient_resident(fwi: FWI, cp, cs, rho, stf_array, shot_ids=None, gpu_id=0, is_masked=False,
                                         cp_ref=None, cs_ref=None, rho_ref=None, reload_obs=False):
    """Same value as compute_misfit_and_gradient with everything between the caller's (cp, cs, rho) and the gradients
    w.r.t. them on the device: a plan per (gpu, shot group) is kept on `fwi`, the observations
    are read from WORKSPACE/Data once (`reload_obs=True` after they change), the source time functions are uploaded
    when they change, and symmetric padding, mask blend, velocity_to_moduli and the chain rule back run as kernels
    (fwi_b200_plan_set_velocities / _get_velocity_gradients).  Per evaluation the host moves 3 model grids down and 3
    gradient grids up, nothing else."""
    if shot_ids is None:
        shot_ids = np.arange(1, len(fwi.ind_src_x) + 1)
    ids0 = np.asarray(shot_ids, dtype=np.int32) - 1
    cache = fwi.__dict__.setdefault("_resident_plans", {})
    key = (int(gpu_id), tuple(int(i) for i in ids0))
    ent = cache.get(key)
    if ent is None:
        ent = cache[key] = {"plan": ops.Plan(fwi.para_path, ids0, gpu_id=gpu_id), "stf": None, "obs": False}
    plan = ent["plan"]
    if reload_obs or not ent["obs"]:
        plan.load_obs_files()
        ent["obs"] = True
    stf = _stf_rows(fwi, stf_array)                       # rows indexed by global shot id (Src_Rec.cu:135)
    if ent["stf"] is None or ent["stf"].shape != stf.shape or not np.array_equal(ent["stf"], stf):
        plan.set_stf(stf)
        ent["stf"] = stf.copy()
    models = [np.asarray(a, dtype=np.float64) for a in (cp, cs, rho)]
    refs = None
    if not is_masked:
        if cp_ref is None or cs_ref is None or rho_ref is None:
            raise ValueError("compute_misfit: cp_ref, cs_ref, rho_ref are required when is_masked is False")
        refs = [np.asarray(a, dtype=np.float64) for a in (cp_ref, cs_ref, rho_ref)]
    if len({a.shape for a in models + (refs or [])}) > 1:      # mixed padded / unpadded inputs: pad on the host
        models = list(try_pad(fwi, *models))
        refs = list(try_pad(fwi, *refs)) if refs else None
    plan.set_velocities(*models, refs=refs, is_masked=is_masked)
    plan.run(1)
    return plan.velocity_gradients()


def timelapse_misfit_and_gradients(surveys, stf_array, shot_ids=None, gpu_ids=(0,), **kw):
    """Time-lapse (flow-coupled) FWI: one misfit + gradient per survey, baseline and monitors, as the reference's
    coupled inversion evaluates them (docs/codes/src_fwi_coupled/main_two_phase_flow_inversion.jl:50-62,84-93: one
    `fwi_op` per survey with its own para file / Data directory, survey i on GPU `i % nGpus`, misfits summed).

    surveys : sequence of (fwi, cp, cs, rho) -- each `fwi` has its own WORKSPACE holding that survey's observations
    gpu_ids : GPUs to spread the surveys over; surveys mapped to the same GPU run one after the other, different
              GPUs run concurrently (the C ABI releases the GIL; one host thread per GPU like TF's inter-op threads)
    returns : (sum of misfits, [(misfit, d/dcp, d/dcs, d/drho) per survey])
    """
    import threading

    gpu_ids = list(gpu_ids)
    out = [None] * len(surveys)
    errs = []

    def worker(g):
        try:
            for i in range(g, len(surveys), len(gpu_ids)):
                fwi, cp, cs, rho = surveys[i]
                out[i] = compute_misfit_and_gradient(fwi, cp, cs, rho, stf_array, shot_ids=shot_ids,
                                                     gpu_id=gpu_ids[g], **kw)
        except Exception as e:  # re-raised in the caller's thread
            errs.append(e)

    threads = [threading.Thread(target=worker, args=(g,)) for g in range(len(gpu_ids))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errs:
        raise errs[0]
    return float(sum(o[0] for o in out)), out


def timelapse_misfit_and_gradients_batched(surveys, stf_array, shot_ids=None, gpu_ids=(0,), is_masked=False, cp_ref=None,
                                           cs_ref=None, rho_ref=None):
    """The same evaluation through ONE C-ABI call (fwi_b200_timelapse): the mask blend and the velocity -> moduli map
    run here, all surveys go down together (survey i on gpu_ids[i % len(gpu_ids)], one host thread per device inside
    the library, plans and observations resident between calls), and the chain rule back to (cp, cs, rho) is applied to
    what comes back.  Same return value as timelapse_misfit_and_gradients."""
    items, models = [], []
    for fwi, cp, cs, rho in surveys:
        cp_m, cs_m, rho_m = _masked_models(fwi, cp, cs, rho, is_masked, cp_ref, cs_ref, rho_ref)
        lam, mu = velocity_to_moduli(cp_m, cs_m, rho_m)
        items.append((fwi.para_path, lam, mu, rho_m))
        models.append((fwi, cp_m, cs_m, rho_m))
    fwi0 = surveys[0][0]
    stf = _stf_rows(fwi0, stf_array)
    if shot_ids is None:
        shot_ids = np.arange(1, len(fwi0.ind_src_x) + 1)
    ids0 = np.asarray(shot_ids, dtype=np.int32) - 1
    res = ops.timelapse(items, stf, list(gpu_ids), ids0)
    out = []
    for (fwi, cp_m, cs_m, rho_m), (j, gl, gm, gd) in zip(models, res):
        g_cp, g_cs, g_rho = moduli_to_velocity_grads(cp_m, cs_m, rho_m, gl, gm, gd)
        if not is_masked:
            g_cp, g_cs, g_rho = g_cp * fwi.mask, g_cs * fwi.mask, g_rho * fwi.mask
        out.append((j, g_cp, g_cs, g_rho))
    return float(sum(o[0] for o in out)), out
