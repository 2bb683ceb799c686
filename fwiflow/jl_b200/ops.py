"""Operator surface of the FWI path, Python mirror of the reference's Julia wrappers.

* ``fwi_op(lambda, mu, den, stf, gpu_id, shot_ids, para_fname)``      -- src/Core.jl:20-31
* ``fwi_obs_op(lambda, mu, den, stf, gpu_id, shot_ids, para_fname)``  -- src/Core.jl:42-53
* ``fwi_op_grad(...)``  -- the FwiOpGrad kernel (deps/CustomOps/FWI/FwiOp.cpp:130-223)
* ``FwiOp``             -- torch.autograd.Function with the same forward/backward pairing as
                           ADCME's ``load_op_and_grad`` (loss from calc_id 0, gradient from calc_id 1)
* ``Plan``              -- device-resident plan (include/fwi_b200.h) for callers that keep
                           inputs on the GPU / all-reduce gradients across GPUs

Same argument meaning as the reference: lambda/mu in MPa, (nz_pad, nx_pad) float64 arrays
(row-major like the TF tensors), stf (nShots, nSteps), shot_ids 0-based int32.
All compute happens in libfwi_b200.so (CUDA, sm_100a); nothing here falls back to the CPU.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from ._lib import FwiError, c_dp, c_fp, c_ip, check

__all__ = ["fwi_op", "fwi_obs_op", "fwi_op_grad", "fwi_op_and_grad", "fwi_op_and_grad_multi", "timelapse", "para_info",
           "grid_info", "FwiOp", "Plan", "release", "FwiError"]


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _dp(a):
    return None if a is None else a.ctypes.data_as(c_dp)


def para_info(para_fname):
    """What the parameter file says (host-only; fwi_b200_para_info): the sizes the C ABI will read and write."""
    out = np.zeros(8, np.int32)
    check(_lib.lib().fwi_b200_para_info(str(para_fname).encode(), out.ctypes.data_as(c_ip)))
    keys = ("nz", "nx", "nSteps", "nPml", "nPad", "if_win", "save_scratch")
    return dict(zip(keys, (int(v) for v in out)))


def _check_shapes(nz, nx, nSteps, lam, mu, den, stf, ids):
    """The C ABI carries no sizes (like the reference's): it reads nz*nx doubles per model array, row shot_id of stf
    for every shot, and writes nz*nx doubles per gradient.  Refuse anything else here, before pointers are handed over."""
    for name, a in (("lambda", lam), ("mu", mu), ("den", den)):
        if a is not None and a.shape != (nz, nx):
            raise FwiError(-1, f"{name} has shape {a.shape}; the parameter file says (nz, nx) = ({nz}, {nx}) "
                               "(padded sizes: src/FWI.jl:13-14)")
    if stf is not None:
        if stf.ndim != 2 or stf.shape[1] != nSteps:
            raise FwiError(-1, f"stf has shape {stf.shape}; expected (nShotsTotal, nSteps = {nSteps})")
        if len(ids) and (int(ids.min()) < 0 or int(ids.max()) >= stf.shape[0]):
            raise FwiError(-1, f"shot ids {int(ids.min())}..{int(ids.max())} do not index the {stf.shape[0]} rows of stf "
                               "(row = global shot id, Src_Rec.cu:135)")


def _prep(lam, mu, den, stf, shot_ids, para_fname):
    lam, mu, den, stf = _f64(lam), _f64(mu), _f64(den), _f64(stf)
    if stf.ndim == 1:
        stf = stf.reshape(1, -1)
    ids = np.ascontiguousarray(np.asarray(shot_ids, dtype=np.int32).ravel())
    if len(ids) == 0:
        raise FwiError(-1, "shot_ids is empty")
    p = para_info(para_fname)
    _check_shapes(p["nz"], p["nx"], p["nSteps"], lam, mu, den, stf, ids)
    return lam, mu, den, stf, ids


def fwi_op(lam, mu, den, stf, gpu_id, shot_ids, para_fname):
    """FWI loss 0.5 * sum(residual^2) over the shots in `shot_ids` (calc_id 0)."""
    lam, mu, den, stf, ids = _prep(lam, mu, den, stf, shot_ids, para_fname)
    misfit = ctypes.c_double(0.0)
    check(_lib.lib().fwi_b200_forward(ctypes.cast(ctypes.byref(misfit), c_dp), _dp(lam), _dp(mu), _dp(den), _dp(stf),
                                      int(gpu_id), len(ids), ids.ctypes.data_as(c_ip), str(para_fname).encode()))
    return float(misfit.value)


def fwi_obs_op(lam, mu, den, stf, gpu_id, shot_ids, para_fname):
    """Forward modelling: writes data_dir_name/Shot<id>.bin, returns 0.0 (calc_id 2)."""
    lam, mu, den, stf, ids = _prep(lam, mu, den, stf, shot_ids, para_fname)
    misfit = ctypes.c_double(0.0)
    check(_lib.lib().fwi_b200_obscalc(ctypes.cast(ctypes.byref(misfit), c_dp), _dp(lam), _dp(mu), _dp(den), _dp(stf),
                                      int(gpu_id), len(ids), ids.ctypes.data_as(c_ip), str(para_fname).encode()))
    return float(misfit.value)


def fwi_op_grad(lam, mu, den, stf, gpu_id, shot_ids, para_fname):
    """Gradients (d/dlambda, d/dmu [per MPa], d/dden, d/dstf) of the loss (calc_id 1).

    grad_stf has shape (nShotsTotal, nSteps): the reference fills row k = position in the group
    and leaves the rest uninitialised; here the rows of the group's GLOBAL shot ids
    are filled and every other row is zero, which is what an optimiser over stf needs.
    """
    lam, mu, den, stf, ids = _prep(lam, mu, den, stf, shot_ids, para_fname)
    gl, gm, gd = np.zeros_like(lam), np.zeros_like(lam), np.zeros_like(lam)
    gs_group = np.zeros((len(ids), stf.shape[1]), np.float64)
    check(_lib.lib().fwi_b200_backward(_dp(gl), _dp(gm), _dp(gd), _dp(gs_group), _dp(lam), _dp(mu), _dp(den), _dp(stf),
                                       int(gpu_id), len(ids), ids.ctypes.data_as(c_ip), str(para_fname).encode()))
    gs = np.zeros_like(stf)
    gs[ids] = gs_group
    return gl, gm, gd, gs


def fwi_op_and_grad(lam, mu, den, stf, gpu_id, shot_ids, para_fname):
    """Loss AND gradients from one forward propagation (the reference propagates twice)."""
    lam, mu, den, stf, ids = _prep(lam, mu, den, stf, shot_ids, para_fname)
    gl, gm, gd = np.zeros_like(lam), np.zeros_like(lam), np.zeros_like(lam)
    gs_group = np.zeros((len(ids), stf.shape[1]), np.float64)
    misfit = ctypes.c_double(0.0)
    check(_lib.lib().fwi_b200_misfit_and_gradient(
        ctypes.cast(ctypes.byref(misfit), c_dp), _dp(gl), _dp(gm), _dp(gd), _dp(gs_group), _dp(lam), _dp(mu), _dp(den),
        _dp(stf), int(gpu_id), len(ids), ids.ctypes.data_as(c_ip), str(para_fname).encode()))
    gs = np.zeros_like(stf)
    gs[ids] = gs_group
    return float(misfit.value), gl, gm, gd, gs


def fwi_cufd_column_major(calc_id, lam, mu, den, stf, gpu_id, shot_ids, para_fname, with_misfit=True):
    """fwi_b200_cufd_ex with layout = 1: lam / mu / den are handed over as COLUMN-major (nz, nx) arrays (Fortran / Julia
    order -- the order the device keeps) and the gradients come back the same way; no transpose on either side.
    Returns (misfit, gl, gm, gd, gs) with Fortran-ordered (nz, nx) gradient arrays."""
    lam, mu, den, stf, ids = _prep(lam, mu, den, stf, shot_ids, para_fname)
    F = lambda a: np.asfortranarray(a)                      # element (z, x) at x * nz + z
    lam, mu, den = F(lam), F(mu), F(den)
    gl, gm, gd = (np.zeros(lam.shape, np.float64, order="F") for _ in range(3))
    gs_group = np.zeros((len(ids), stf.shape[1]), np.float64)
    misfit = ctypes.c_double(0.0)
    fp = lambda a: a.ctypes.data_as(c_dp)
    check(_lib.lib().fwi_b200_cufd_ex(ctypes.cast(ctypes.byref(misfit), c_dp), fp(gl), fp(gm), fp(gd), _dp(gs_group), fp(lam),
                                      fp(mu), fp(den), _dp(stf), int(calc_id), int(gpu_id), len(ids),
                                      ids.ctypes.data_as(c_ip), str(para_fname).encode(), 1, 1 if with_misfit else 0))
    gs = np.zeros_like(stf)
    gs[ids] = gs_group
    return float(misfit.value), gl, gm, gd, gs


def fwi_op_and_grad_multi(lam, mu, den, stf, gpu_ids, shot_ids, para_fname):
    """Loss and gradients with the shots of the group sharded over several GPUs of THIS process
    (fwi_b200_gradient_multi): shot k goes to gpu_ids[k % len(gpu_ids)], the devices run concurrently."""
    lam, mu, den, stf, ids = _prep(lam, mu, den, stf, shot_ids, para_fname)
    gpus = np.ascontiguousarray(np.asarray(gpu_ids, dtype=np.int32).ravel())
    gl, gm, gd = np.zeros_like(lam), np.zeros_like(lam), np.zeros_like(lam)
    gs_group = np.zeros((len(ids), stf.shape[1]), np.float64)
    misfit = ctypes.c_double(0.0)
    check(_lib.lib().fwi_b200_gradient_multi(
        ctypes.cast(ctypes.byref(misfit), c_dp), _dp(gl), _dp(gm), _dp(gd), _dp(gs_group), _dp(lam), _dp(mu), _dp(den),
        _dp(stf), len(gpus), gpus.ctypes.data_as(c_ip), len(ids), ids.ctypes.data_as(c_ip), str(para_fname).encode()))
    gs = np.zeros_like(stf)
    gs[ids] = gs_group
    return float(misfit.value), gl, gm, gd, gs


def timelapse(surveys, stf, gpu_ids, shot_ids):
    """fwi_b200_timelapse: misfit + gradients of several surveys (baseline + monitors) in one call.
    surveys: sequence of (para_fname, lam, mu, den).  Returns [(misfit, gl, gm, gd)] in survey order."""
    S = len(surveys)
    if S == 0:
        return []
    ids = np.ascontiguousarray(np.asarray(shot_ids, dtype=np.int32).ravel())
    gpus = np.ascontiguousarray(np.asarray(gpu_ids, dtype=np.int32).ravel())
    stf = _f64(stf)
    if stf.ndim == 1:
        stf = stf.reshape(1, -1)
    models, paras = [], []
    for para, lam, mu, den in surveys:
        lam, mu, den = _f64(lam), _f64(mu), _f64(den)
        p = para_info(para)
        _check_shapes(p["nz"], p["nx"], p["nSteps"], lam, mu, den, stf, ids)
        models.append((lam, mu, den))
        paras.append(str(para).encode())
    grads = [tuple(np.zeros_like(m[0]) for _ in range(3)) for m in models]
    misfit = np.zeros(S, np.float64)
    arr = lambda items: (c_dp * S)(*[_dp(a) for a in items])
    check(_lib.lib().fwi_b200_timelapse(
        S, (ctypes.c_char_p * S)(*paras), arr([m[0] for m in models]), arr([m[1] for m in models]),
        arr([m[2] for m in models]), _dp(stf), len(gpus), gpus.ctypes.data_as(c_ip), len(ids), ids.ctypes.data_as(c_ip),
        _dp(misfit), arr([g[0] for g in grads]), arr([g[1] for g in grads]), arr([g[2] for g in grads])))
    return [(float(misfit[i]),) + grads[i] for i in range(S)]


def set_option(name, value):
    """Developer A/B switches of the library (fwi_b200_set_option; include/fwi_b200.h has the details):
    "rev_lean" (-1 auto / 0 / 1), "merged_bwd" (0 / 1), "frame_ring" (2 / 5, plans created afterwards),
    "acc_group" (0 auto, k shots per imaging-accumulator slot of the reverse step), "dyn_units" (1 / 0: the reverse
    step's work units claimed from a device counter / dealt round-robin).  All of them leave the results unchanged up
    to the order of float sums."""
    check(_lib.lib().fwi_b200_set_option(str(name).encode(), int(value)))


def grid_info(para_fname):
    """The device layout derived from a parameter file (host-only; fwi_b200_grid_info)."""
    out = np.zeros(12, np.int32)
    check(_lib.lib().fwi_b200_grid_info(str(para_fname).encode(), out.ctypes.data_as(c_ip)))
    keys = ("nz", "nx", "pitch", "zlive", "z_off", "tiles_z", "tiles_x", "frame_len", "zlo", "zhi", "xlo", "xhi")
    return dict(zip(keys, (int(v) for v in out)))


def release():
    """Free the cached device contexts behind the host-buffer entry points."""
    _lib.lib().fwi_b200_release()


try:  # torch is plumbing: only needed for the autograd wrapper and device-tensor interop
    import torch

    class FwiOp(torch.autograd.Function):
        """loss = FwiOp.apply(lam, mu, den, stf, gpu_id, shot_ids, para_fname) with autograd.

        Like the reference's FwiOpGrad (FwiOp.cpp:220-222) the upstream gradient is NOT ignored here:
        the returned gradients are scaled by grad_output (the reference drops it).
        """

        @staticmethod
        def forward(ctx, lam, mu, den, stf, gpu_id, shot_ids, para_fname):
            ctx.save_for_backward(lam, mu, den, stf)
            ctx.meta = (int(gpu_id), np.asarray(shot_ids, dtype=np.int32), str(para_fname))
            v = fwi_op(lam.detach().cpu().numpy(), mu.detach().cpu().numpy(), den.detach().cpu().numpy(),
                       stf.detach().cpu().numpy(), *ctx.meta)
            return torch.tensor(v, dtype=torch.float64)

        @staticmethod
        def backward(ctx, gout):
            lam, mu, den, stf = ctx.saved_tensors
            gl, gm, gd, gs = fwi_op_grad(lam.detach().cpu().numpy(), mu.detach().cpu().numpy(),
                                         den.detach().cpu().numpy(), stf.detach().cpu().numpy(), *ctx.meta)
            s = float(gout)
            mk = lambda a, ref: torch.from_numpy(a * s).to(dtype=ref.dtype, device=ref.device).reshape(ref.shape)
            return mk(gl, lam), mk(gm, mu), mk(gd, den), mk(gs, stf), None, None, None
except Exception:  # pragma: no cover
    torch = None
    FwiOp = None


class Plan:
    """Device-resident FWI plan (fwi_b200_plan_* in include/fwi_b200.h)."""

    def __init__(self, para_fname, shot_ids, gpu_id=0, max_batch=0):
        self._L = _lib.lib()
        self._h = ctypes.c_void_p()
        ids = np.ascontiguousarray(np.asarray(shot_ids, dtype=np.int32).ravel())
        self.shot_ids = ids
        check(self._L.fwi_b200_plan_create(ctypes.byref(self._h), str(para_fname).encode(), int(gpu_id), len(ids),
                                           ids.ctypes.data_as(c_ip), int(max_batch)))
        v = [ctypes.c_int() for _ in range(8)]
        check(self._L.fwi_b200_plan_info(self._h, *[ctypes.byref(x) for x in v]))
        (self.nz, self.nx, self.nSteps, self.nPml, self.nPad, self.group_size, self.batch,
         self.max_nrec) = [x.value for x in v]
        self.gpu_id = int(gpu_id)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.fwi_b200_plan_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_model(self, lam, mu, den):
        lam, mu, den = _f64(lam), _f64(mu), _f64(den)
        _check_shapes(self.nz, self.nx, self.nSteps, lam, mu, den, None, self.shot_ids)
        check(self._L.fwi_b200_plan_set_model(self._h, _dp(lam), _dp(mu), _dp(den)))

    def set_velocities(self, cp, cs, rho, refs=None, is_masked=False):
        """cp, cs, rho -> symmetric padding, mask blend with `refs` = (cp_ref, cs_ref, rho_ref) unless is_masked,
        velocity_to_moduli and the model planes, all on the device (src/FWI.jl:156-205).  The grids are either
        unpadded (nz - 2 nPml - nPad, nx - 2 nPml) or padded (nz, nx), refs like the models."""
        arrs = [_f64(a) for a in (cp, cs, rho)]
        inner = (self.nz - 2 * self.nPml - self.nPad, self.nx - 2 * self.nPml)
        if arrs[0].shape not in (inner, (self.nz, self.nx)):
            raise FwiError(-1, f"set_velocities: cp has shape {arrs[0].shape}; the para file gives {inner} unpadded or "
                               f"{(self.nz, self.nx)} padded")
        if not is_masked:
            if refs is None:
                raise FwiError(-1, "set_velocities: refs=(cp_ref, cs_ref, rho_ref) is required unless is_masked")
            arrs += [_f64(a) for a in refs]
        for a in arrs:
            if a.shape != arrs[0].shape:
                raise FwiError(-1, f"set_velocities: shapes differ: {a.shape} vs {arrs[0].shape}")
        ptrs = [_dp(a) for a in arrs] + [None] * (6 - len(arrs))
        check(self._L.fwi_b200_plan_set_velocities(self._h, *ptrs, 1 if is_masked else 0,
                                                   1 if arrs[0].shape == (self.nz, self.nx) else 0))

    def velocity_gradients(self):
        """(misfit, g_cp, g_cs, g_rho) on the padded grid after run(1) on a model given by set_velocities."""
        misfit = ctypes.c_double(0.0)
        g = [np.zeros((self.nz, self.nx)) for _ in range(3)]
        check(self._L.fwi_b200_plan_get_velocity_gradients(self._h, ctypes.cast(ctypes.byref(misfit), c_dp),
                                                           *[_dp(a) for a in g]))
        return (float(misfit.value), *g)

    def set_stf(self, stf):
        stf = _f64(stf)
        if stf.ndim == 1:
            stf = stf.reshape(1, -1)
        _check_shapes(self.nz, self.nx, self.nSteps, None, None, None, stf, self.shot_ids)
        check(self._L.fwi_b200_plan_set_stf(self._h, _dp(stf)))

    def set_obs(self, ishot, obs):
        """Observed data of the i-th shot of the group, (nrec, nSteps) float32 with time fastest (the Shot<id>.bin
        layout), kept resident on the device: no disk round trip per evaluation."""
        obs = np.ascontiguousarray(np.asarray(obs, dtype=np.float32))
        if not 0 <= int(ishot) < self.group_size:
            raise FwiError(-1, f"set_obs: shot position {ishot} outside the group of {self.group_size}")
        nrec = self.shot_geometry(ishot)[2]
        if obs.shape != (nrec, self.nSteps):
            raise FwiError(-1, f"set_obs: obs has shape {obs.shape}; shot {int(self.shot_ids[ishot])} has "
                               f"(nrec, nSteps) = ({nrec}, {self.nSteps})")
        check(self._L.fwi_b200_plan_set_obs(self._h, int(ishot), obs.ctypes.data_as(c_fp)))

    def load_obs_files(self):
        check(self._L.fwi_b200_plan_load_obs_files(self._h))

    def run(self, calc_id, stream=None, sync=True):
        """Enqueue one evaluation.  `stream`: a cudaStream_t handle (int) of a stream created by the caller, e.g.
        `torch.cuda.Stream().cuda_stream`; None = the plan's own non-blocking stream.  The handle 0 (CUDA's legacy
        default stream, what `torch.cuda.current_stream().cuda_stream` is unless a stream context is active) cannot be
        told apart from "none" by the C ABI and is refused here: work enqueued on the plan's own stream would NOT be
        ordered with anything the caller does on the default stream (events, collectives)."""
        if stream is not None and int(stream) == 0:
            raise FwiError(-1, "Plan.run: stream handle 0 (legacy default stream) -- pass a created stream's handle, or "
                               "None for the plan's own stream (then synchronise with sync=True)")
        check(self._L.fwi_b200_plan_run(self._h, int(calc_id), ctypes.c_void_p(int(stream)) if stream is not None else None,
                                        1 if sync else 0))

    def result(self, with_grad=True):
        misfit = ctypes.c_double(0.0)
        mp = ctypes.cast(ctypes.byref(misfit), c_dp)
        if not with_grad:
            check(self._L.fwi_b200_plan_get_result(self._h, mp, None, None, None, None))
            return float(misfit.value)
        gl = np.zeros((self.nz, self.nx)); gm = np.zeros_like(gl); gd = np.zeros_like(gl)
        gs = np.zeros((self.group_size, self.nSteps))
        check(self._L.fwi_b200_plan_get_result(self._h, mp, _dp(gl), _dp(gm), _dp(gd), _dp(gs)))
        return float(misfit.value), gl, gm, gd, gs

    def result_device_ptr(self):
        return int(self._L.fwi_b200_plan_result_device(self._h)), int(self._L.fwi_b200_plan_result_count(self._h))

    def result_tensor(self):
        """The packed [grad_lambda|grad_mu|grad_den|misfit] float32 buffer as a torch CUDA tensor (no copy)."""
        ptr, n = self.result_device_ptr()

        class _Wrap:  # __cuda_array_interface__ view of memory owned by the plan
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}

        return torch.as_tensor(_Wrap(), device=f"cuda:{self.gpu_id}")

    def traces(self, ishot, which=0):
        nrec = self.shot_geometry(ishot)[2]
        out = np.zeros((nrec, self.nSteps), np.float32)
        check(self._L.fwi_b200_plan_get_traces(self._h, int(ishot), int(which), out.ctypes.data_as(c_fp)))
        return out

    def write_obs_files(self):
        check(self._L.fwi_b200_plan_write_obs_files(self._h))

    def shot_geometry(self, ishot):
        zs, xs, nr = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        check(self._L.fwi_b200_plan_shot_geometry(self._h, int(ishot), ctypes.byref(zs), ctypes.byref(xs),
                                                  ctypes.byref(nr), None, None))
        zr = np.zeros(nr.value, np.int32); xr = np.zeros(nr.value, np.int32)
        check(self._L.fwi_b200_plan_shot_geometry(self._h, int(ishot), None, None, None, zr.ctypes.data_as(c_ip),
                                                  xr.ctypes.data_as(c_ip)))
        return zs.value, xs.value, nr.value, zr, xr

    def launch_count(self):
        return int(self._L.fwi_b200_plan_launch_count(self._h))

    def field(self, ishot, field):
        out = np.zeros((self.nz, self.nx), np.float32)
        check(self._L.fwi_b200_plan_get_field(self._h, int(ishot), int(field), out.ctypes.data_as(c_fp)))
        return out

    def time_kernel(self, which, iters=20, stream=None):
        ms = ctypes.c_float(0.0)
        nbytes = ctypes.c_double(0.0)
        check(self._L.fwi_b200_plan_time_kernel(self._h, int(which), int(iters),
                                                ctypes.c_void_p(int(stream)) if stream else None,
                                                ctypes.byref(ms), ctypes.cast(ctypes.byref(nbytes), c_dp)))
        return float(ms.value), float(nbytes.value)
