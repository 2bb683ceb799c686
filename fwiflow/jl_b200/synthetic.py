"""Seeded synthetic workloads for the FWI hot path (SURVEY.md section 8d / BASELINE.md).

There is no network for datasets, so every benchmark / parity input is generated
here: layered elastic models (Gardner density, cs = cp/sqrt(3)), survey geometries
and Ricker sources with the shapes of BASELINE.json's configs C1..C5.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

from .utils import (nPad_rule, paraGen, sourceGene, surveyGen, symmetric_pad,
                    velocity_to_moduli)

SEED = 20191207


@dataclass
class Case:
    """One FWI problem: unpadded grid, survey, source, and the (padded) models."""
    name: str
    nz: int
    nx: int
    dz: float
    dx: float
    dt: float
    nSteps: int
    f0: float
    z_src: np.ndarray
    x_src: np.ndarray
    z_rec: np.ndarray
    x_rec: np.ndarray
    nPml: int = 32
    nPad: int | None = None
    cp_true: np.ndarray | None = None   # padded (nz_pad, nx_pad) float64
    cs_true: np.ndarray | None = None
    rho_true: np.ndarray | None = None
    cp_init: np.ndarray | None = None
    cs_init: np.ndarray | None = None
    rho_init: np.ndarray | None = None
    stf: np.ndarray | None = None       # (nShots, nSteps) float64
    extra: dict = field(default_factory=dict)

    def __post_init__(self):
        if self.nPad is None:
            self.nPad = nPad_rule(self.nz, self.nPml)

    @property
    def nz_pad(self):
        return self.nz + 2 * self.nPml + self.nPad

    @property
    def nx_pad(self):
        return self.nx + 2 * self.nPml

    @property
    def nShots(self):
        return len(self.x_src)

    @property
    def nrec(self):
        return len(self.x_rec)

    def moduli(self, which="true"):
        cp, cs, rho = (self.cp_true, self.cs_true, self.rho_true) if which == "true" else (
            self.cp_init, self.cs_init, self.rho_init)
        lam, mu = velocity_to_moduli(cp, cs, rho)
        return np.ascontiguousarray(lam), np.ascontiguousarray(mu), np.ascontiguousarray(rho, dtype=np.float64)

    def write_files(self, workdir, scratch=False):
        """para_file.json, survey_file.json, Data/ under `workdir` (src/FWI.jl:53-58)."""
        os.makedirs(workdir, exist_ok=True)
        para = os.path.join(workdir, "para_file.json")
        survey = os.path.join(workdir, "survey_file.json")
        data = os.path.join(workdir, "Data")
        win = self.extra.get("windows")          # {"shot<i>": {"start": [...], "end": [...]}} -> para "if_win"
        paraGen(self.nz_pad, self.nx_pad, self.dz, self.dx, self.nSteps, self.dt, self.f0, self.nPml,
                self.nPad, para, survey, data, if_win=win is not None,
                scratch_dir_name=os.path.join(workdir, "Scratch") if scratch else "")
        surveyGen(self.z_src, self.x_src, self.z_rec, self.x_rec, survey, Windows=win,
                  Weights=self.extra.get("weights"))
        return para


def layered_cp(nz, nx, nlayers, rng, jitter=0.02, vmin=1500.0, vmax=4000.0):
    """cp(z) = vmin + (vmax-vmin) z/(nz-1) quantised into `nlayers` equal layers, +-jitter per layer."""
    z = np.arange(nz, dtype=np.float64)
    layer = np.minimum((z * nlayers / nz).astype(int), nlayers - 1)
    centres = (np.arange(nlayers) + 0.5) * nz / nlayers
    v = vmin + (vmax - vmin) * centres / max(nz - 1, 1)
    v = v * (1.0 + jitter * (2.0 * rng.random(nlayers) - 1.0))
    return np.repeat(v[layer][:, None], nx, axis=1)


def smooth_1d(cp, sigma):
    """Gaussian smoothing along z (sigma in cells) with edge replication -- the 'initial model'."""
    nz = cp.shape[0]
    r = int(4 * sigma)
    k = np.exp(-0.5 * (np.arange(-r, r + 1) / sigma) ** 2)
    k /= k.sum()
    padded = np.pad(cp, ((r, r), (0, 0)), mode="edge")
    out = np.zeros_like(cp)
    for i, w in enumerate(k):
        out += w * padded[i:i + nz]
    return out


def _elastic_from_cp(cp):
    cs = cp / np.sqrt(3.0)
    rho = 310.0 * cp ** 0.25  # Gardner
    return cs, rho


def make_layered_case(name, nz, nx, dz, dt, nSteps, f0, nlayers, nshots, rec_margin=3, src_z=2, rec_z=2,
                      smooth_sigma=10.0, seed=SEED, nPml=32, elastic=True, vmax=4000.0):
    rng = np.random.default_rng(seed)
    cp = layered_cp(nz, nx, nlayers, rng, vmax=vmax)
    cp0 = smooth_1d(cp, smooth_sigma)
    if elastic:
        cs, rho = _elastic_from_cp(cp)
        cs0, rho0 = _elastic_from_cp(cp0)
    else:
        cs, rho = np.zeros_like(cp), np.full_like(cp, 2500.0)
        cs0, rho0 = np.zeros_like(cp0), np.full_like(cp0, 2500.0)
    x_src = np.round(np.linspace(4, nx - 5, nshots)).astype(np.int64)
    z_srcs = np.full(nshots, src_z, dtype=np.int64)
    x_rec = np.arange(rec_margin, nx - rec_margin, dtype=np.int64)
    z_rec = np.full(x_rec.shape, rec_z, dtype=np.int64)
    c = Case(name=name, nz=nz, nx=nx, dz=dz, dx=dz, dt=dt, nSteps=nSteps, f0=f0, z_src=z_srcs, x_src=x_src,
             z_rec=z_rec, x_rec=x_rec, nPml=nPml)
    pad = lambda a: symmetric_pad(a, c.nPml, c.nPad)
    c.cp_true, c.cs_true, c.rho_true = pad(cp), pad(cs), pad(rho)
    c.cp_init, c.cs_init, c.rho_init = pad(cp0), pad(cs0), pad(rho0)
    c.stf = np.repeat(sourceGene(f0, nSteps, dt), nshots, axis=0)
    return c


def case_c1(nSteps=1000):
    """C1: 100x100 homogeneous (padded 192x164), 1 shot, forward modelling (gradtest.jl:57-61 values)."""
    nz = nx = 100
    x_rec = np.arange(3, 97, dtype=np.int64)
    c = Case(name="C1", nz=nz, nx=nx, dz=20.0, dx=20.0, dt=0.0025, nSteps=nSteps, f0=4.5,
             z_src=np.array([50]), x_src=np.array([50]), z_rec=np.full(x_rec.shape, 2, dtype=np.int64),
             x_rec=x_rec)
    shape = (c.nz_pad, c.nx_pad)
    c.cp_true = np.full(shape, 3000.0)
    c.cs_true = np.full(shape, 3000.0 / np.sqrt(3.0))
    c.rho_true = np.full(shape, 2000.0)
    c.cp_init, c.cs_init, c.rho_init = c.cp_true * 1.03, c.cs_true * 1.03, c.rho_true.copy()
    c.stf = sourceGene(4.5, nSteps, 0.0025)
    return c


def case_c2(nshots=30, nSteps=2000):
    """C2: Marmousi-sized 134x384 (padded 224x448), 30 shots, 379 receivers, gradient."""
    c = make_layered_case("C2", 134, 384, 24.0, 0.0025, nSteps, 4.5, nlayers=8, nshots=nshots)
    # sources x = 4:13:384 (30 shots) as in SURVEY.md section 8d
    xs = np.arange(4, 384, 13, dtype=np.int64)[:nshots] if nshots <= 30 else c.x_src
    c.x_src = xs
    c.z_src = np.full(xs.shape, 2, dtype=np.int64)
    c.x_rec = np.arange(3, 382, dtype=np.int64)          # x = 3..381 at z = 2: the 379 receivers of src/FWI.jl:90
    c.z_rec = np.full(c.x_rec.shape, 2, dtype=np.int64)
    c.stf = np.repeat(sourceGene(4.5, nSteps, 0.0025), len(xs), axis=0)
    return c


def case_c3(nshots=200, nSteps=4000):
    """C3: 1000x3000 (padded 1088x3064), 200 shots, 2994 receivers."""
    return make_layered_case("C3", 1000, 3000, 10.0, 0.001, nSteps, 5.0, nlayers=16, nshots=nshots)


def case_c5(nshots=512, nSteps=8000):
    """C5: 4000x8000 (padded 4096x8064), 512 shots."""
    return make_layered_case("C5", 4000, 8000, 10.0, 0.001, nSteps, 5.0, nlayers=32, nshots=nshots)


def case_small(name="S", nz=60, nx=80, nSteps=600, nshots=2, elastic=True, nlayers=4, dz=20.0, dt=0.002,
               f0=6.0, seed=SEED):
    """Small layered case for CPU-sized parity tests (padded 128x144)."""
    return make_layered_case(name, nz, nx, dz, dt, nSteps, f0, nlayers=nlayers, nshots=nshots,
                             smooth_sigma=4.0, seed=seed, elastic=elastic, vmax=3500.0)


def case_small_windows(name="small_windows", seed=11):
    """case_small with per-trace time windows and trace weights (para "if_win", Src_Rec.cu:157-200): a moveout-like
    window per receiver, random weights, and the edge cases of cuda_window (utilities.cu:654-706) -- limits outside
    the record (clamped), an empty window (the reference then leaves the trace untouched) and a zero weight."""
    c = case_small(name, elastic=True, seed=seed)
    rng = np.random.default_rng(seed)
    t_max = c.nSteps * c.dt
    windows, weights = {}, {}
    for i in range(c.nShots):
        off = np.abs((c.x_rec - c.x_src[i]) * c.dx)
        start = 0.08 + off / 3500.0 + 0.02 * rng.random(c.nrec)
        end = np.minimum(start + 0.45 + 0.1 * rng.random(c.nrec), 1.4 * t_max)
        w = 0.5 + rng.random(c.nrec)
        start[1], end[1] = -0.3, 2.0 * t_max       # clamped to [0, t_max]
        start[3], end[3] = 0.7, 0.2                # empty window: trace untouched, weight ignored
        w[5] = 0.0                                 # muted trace
        windows[f"shot{i}"] = {"start": [float(v) for v in start], "end": [float(v) for v in end]}
        weights[f"shot{i}"] = {"weights": [float(v) for v in w]}
    c.extra["windows"] = windows
    c.extra["weights"] = weights
    return c


def case_aniso(name="aniso", seed=21):
    """Unequal grid spacings (dz = 16 m, dx = 24 m), a 20-cell absorbing layer, receivers on a dipping line and two
    sources at different depths: everything the other golden cases keep equal or default."""
    nz, nx, nPml, nSteps, dt, f0 = 44, 66, 20, 520, 0.0016, 7.0
    rng = np.random.default_rng(seed)
    cp = layered_cp(nz, nx, 3, rng, vmax=3200.0)
    cp0 = smooth_1d(cp, 3.0)
    cs, rho = _elastic_from_cp(cp)
    cs0, rho0 = _elastic_from_cp(cp0)
    x_rec = np.arange(3, nx - 3, dtype=np.int64)
    z_rec = 2 + (x_rec // 12)
    c = Case(name=name, nz=nz, nx=nx, dz=16.0, dx=24.0, dt=dt, nSteps=nSteps, f0=f0, z_src=np.array([3, 9]),
             x_src=np.array([12, 50]), z_rec=z_rec, x_rec=x_rec, nPml=nPml)
    pad = lambda a: symmetric_pad(a, c.nPml, c.nPad)
    c.cp_true, c.cs_true, c.rho_true = pad(cp), pad(cs), pad(rho)
    c.cp_init, c.cs_init, c.rho_init = pad(cp0), pad(cs0), pad(rho0)
    c.stf = np.repeat(sourceGene(f0, nSteps, dt), 2, axis=0)
    return c


def case_gradtest_small(n=110, nSteps=500):
    """gradtest.jl-like: grid given INCLUDING the PML, nPad = 0, nz not a multiple of 32,
    one centre source, a lattice of receivers throughout the volume
    (deps/CustomOps/FWI/gradtest.jl:16-61, shrunk)."""
    nPml = 32
    inner = n - 2 * nPml
    pts = np.arange(5, inner - 4, 10, dtype=np.int64)
    xr, zr = np.meshgrid(pts, pts)
    c = Case(name="GT", nz=inner, nx=inner, dz=20.0, dx=20.0, dt=0.0025, nSteps=nSteps, f0=4.5,
             z_src=np.array([inner // 2]), x_src=np.array([inner // 2]), z_rec=zr.ravel(), x_rec=xr.ravel(),
             nPml=nPml, nPad=0)
    shape = (c.nz_pad, c.nx_pad)
    c.cp_true = np.full(shape, 3000.0)
    c.cs_true = np.full(shape, 3000.0 / np.sqrt(3.0))
    c.rho_true = np.full(shape, 2000.0)
    rng = np.random.default_rng(233)
    bump = np.zeros(shape)
    bump[nPml + 5:n - nPml - 5, nPml + 5:n - nPml - 5] = 1.0
    c.cp_init = c.cp_true * (1.0 + 0.04 * bump * rng.random(shape))
    c.cs_init = c.cp_init / np.sqrt(3.0)
    c.rho_init = c.rho_true * (1.0 + 0.03 * bump * rng.random(shape))
    c.stf = sourceGene(4.5, nSteps, 0.0025)
    return c
