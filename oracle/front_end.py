"""TEST INFRASTRUCTURE ONLY -- numpy restatement of what FwiFlow.jl does in TensorFlow AROUND the op for the
velocity-space interface, and of the derivative TensorFlow's autodiff takes through it.  Checker for
fwi_b200_plan_set_velocities / fwi_b200_plan_get_velocity_gradients; the product never imports this.

PARITY UNPINNED against the reference itself for this file: the reference's front end is Julia + TensorFlow, neither of
which runs in this image, and the reference's tests hold no golden vectors for it.  What pins it instead: np.pad's
"symmetric" mode as a second statement of tf.pad SYMMETRIC, the mask's definition, and central differences of its own
forward maps for the chain rule (tests/test_oracle_golden.py).  (The propagator oracle, oracle/fwi_oracle.cpp, IS pinned
against the reference's own op: tests/golden/.)

  padding            /root/reference/src/FWI.jl:193-205   tf.pad(cp, [nPml (nPml+nPad); nPml nPml], "SYMMETRIC")
  mask               /root/reference/src/FWI.jl:45-49     ones inside the absorbing layers, minus 10 rows under the top one
  mask blend         /root/reference/src/FWI.jl:174-176   cp .* mask + cp_ref .* mask_neg
  velocity_to_moduli /root/reference/src/Utils.jl:221-227 lambda = (cp.*cp - 2.0 * cs.*cs) .* den / 1e6, mu = cs.*cs .* den / 1e6
"""
from __future__ import annotations

import numpy as np


def symmetric_index(i, pad, n):
    """Source index of padded index i for SYMMETRIC padding of an axis of length n by `pad` cells in front: the edge
    cell is repeated ([c b a | a b c d | d c b]), and the reflection goes on periodically for pads longer than n."""
    t = i - pad
    m = 2 * n
    r = t % m                      # Python's % is non-negative for m > 0
    return r if r < n else m - 1 - r


def padding(a, nPml, nPad):
    """FWI.jl:193-205 by explicit index arithmetic (checked against np.pad(mode="symmetric") in tests/)."""
    a = np.asarray(a, dtype=np.float64)
    nz0, nx0 = a.shape
    zi = np.array([symmetric_index(i, nPml, nz0) for i in range(nz0 + 2 * nPml + nPad)])
    xi = np.array([symmetric_index(i, nPml, nx0) for i in range(nx0 + 2 * nPml)])
    return a[np.ix_(zi, xi)]


def mask(nz0, nx0, nPml, nPad):
    """FWI.jl:45-49 (1-based Julia ranges nPml+1 : nPml+nz etc. restated 0-based)."""
    m = np.zeros((nz0 + 2 * nPml + nPad, nx0 + 2 * nPml))
    m[nPml:nPml + nz0, nPml:nPml + nx0] = 1.0
    m[nPml:nPml + 10, :] = 0.0
    return m


def velocity_to_moduli(cp, cs, den):
    """Utils.jl:221-227, same operation order."""
    lam = (cp * cp - 2.0 * cs * cs) * den / 1e6
    mu = cs * cs * den / 1e6
    return lam, mu


def front_end(cp, cs, rho, nPml, nPad, is_masked, refs=None, shape_padded=None):
    """compute_misfit's graph up to the op (FWI.jl:165-177): pad what is not padded yet, blend with the padded
    reference models outside the mask unless is_masked, map to (lambda, mu, rho).  Returns the masked padded
    velocities too (the chain rule needs them) and the mask."""
    arrs = [np.asarray(a, dtype=np.float64) for a in (cp, cs, rho)]
    if shape_padded is None or arrs[0].shape != tuple(shape_padded):
        nz0, nx0 = arrs[0].shape
        arrs = [padding(a, nPml, nPad) for a in arrs]
    else:
        nz0, nx0 = shape_padded[0] - 2 * nPml - nPad, shape_padded[1] - 2 * nPml
    m = mask(nz0, nx0, nPml, nPad)
    if not is_masked:
        r = [np.asarray(a, dtype=np.float64) for a in refs]
        r = [a if a.shape == m.shape else padding(a, nPml, nPad) for a in r]
        arrs = [a * m + b * (1.0 - m) for a, b in zip(arrs, r)]
    lam, mu = velocity_to_moduli(*arrs)
    return lam, mu, arrs[2], arrs, m


def chain_rule(vel, g_lam, g_mu, g_den, m, is_masked):
    """d misfit / d (cp_pad, cs_pad, rho_pad) from d misfit / d (lambda, mu, rho): the derivative of
    velocity_to_moduli (Utils.jl:221-227) and of the blend `x .* mask + ref .* mask_neg` (FWI.jl:174-176)."""
    cp, cs, den = vel
    g_cp = 2.0 * cp * den / 1e6 * g_lam
    g_cs = (-4.0 * g_lam + 2.0 * g_mu) * cs * den / 1e6
    g_rho = g_den + ((cp * cp - 2.0 * cs * cs) * g_lam + cs * cs * g_mu) / 1e6
    if not is_masked:
        g_cp, g_cs, g_rho = g_cp * m, g_cs * m, g_rho * m
    return g_cp, g_cs, g_rho
