"""CPU oracle for SURVEY.md 8f-4: the Q1 strain-energy evaluator of the reference's deep-energy
back-end.  TEST INFRASTRUCTURE ONLY (who may import ``oracle/``: see fem_oracle.py header).

Restates in numpy (reference file:line):
* ``ObjectiveCalculator`` ........ DEM_src/objective_calculator.py:15-143: 2x2 Gauss rule on a uniform
  grid of Q1 cells, shape-function derivatives with the reference's 9-digit constants
  a = 0.394337567, b = 0.105662433 (``get_shape_derivatives``, :38-49), Jacobian diag(dx/2, dy/2)
  (:26-36), node order N1=(iy,ix), N2=(iy+1,ix), N3=(iy,ix+1), N4=(iy+1,ix+1) (``get_gauss_points``
  :51-70), displacement stored flattened with index ix*(Ny+1)+iy (``evaluate`` :126-131).
* ``StrainEnergy`` ............... DEM_src/elasisity_problem.py:64-129: sigma:eps with
  eps = (grad u + grad u^T)/2, sigma = lambda div u I + 2 mu eps, summed over the 4 Gauss points and
  multiplied by det J; objective sum r(rho) e, gradient -r'(rho) e; internal energy 1/2 sum r(rho) e.
* ``ElasticPenalizer`` ........... src/penalizers.py:29-46 (m = 1e-6).

The reference computes in torch float32; ``dtype=np.float32`` mirrors its operation order (sums over
cells then run in float32 pairwise order of numpy, not torch's: the totals agree to float32 rounding,
the per-cell values to 1 ulp or better).  ``energy_gradient_u`` (d energy / d u, which the reference
gets from autograd) is written out analytically and evaluated in float64.

Pinned by the reference's fixtures tests/test_data/DEM/{short_cantilever,bridge}/problem_data.dat
(tests/test_DEM_problem.py:16-47) and by outputs of the reference code itself on random fields,
generated in the build container (tests/golden/make_golden.py -> dem_strain_energy_reference.json).
"""
from __future__ import annotations

import numpy as np

A = 0.394337567
B = 0.105662433
# [node][d/ds | d/dt][gauss point]   (DEM_src/objective_calculator.py:44-49)
SHAPE_DERIVATIVES = np.array([
    [[-A, -B, -B, -A], [-A, -A, -B, -B]],
    [[-B, -A, -A, -B], [A, A, B, B]],
    [[A, B, B, A], [-B, -B, -A, -A]],
    [[B, A, A, B], [B, B, A, A]],
])


def lame(E, nu):
    mu = E / (2 * (1 + nu))
    return mu * nu / (0.5 - nu), mu  # lambda, mu  (DEM_src/elasisity_problem.py:79-80)


def node_values(u, shape):
    """u: (Nx+1)*(Ny+1) x 2 flattened as the reference does -> four (2, Ny, Nx) corner arrays."""
    ny1, nx1 = shape
    U = np.transpose(u.reshape(nx1, ny1, 2), (1, 0, 2))  # (Ny+1, Nx+1, 2)
    corners = [U[:-1, :-1], U[1:, :-1], U[:-1, 1:], U[1:, 1:]]
    return [np.moveaxis(c, 2, 0) for c in corners]


def strain_energy_density(u, shape, dxdy, lam, mu, dtype=np.float32):
    """sum over the 4 Gauss points of sigma:eps, times det J -> (Ny, Nx)."""
    u = np.asarray(u, dtype=dtype)
    dx, dy = dxdy
    jinv = np.array([1.0 / (dx / 2), 1.0 / (dy / 2)])  # np.linalg.inv of the diagonal Jacobian
    detj = (dx / 2) * (dy / 2)
    corners = node_values(u, shape)
    total = None
    for g in range(4):
        grad = np.zeros((2, 2) + corners[0].shape[1:], dtype=dtype)  # [component][d/dx | d/dy]
        for i in range(4):
            ddx = dtype(SHAPE_DERIVATIVES[i, 0, g] * jinv[0])
            ddy = dtype(SHAPE_DERIVATIVES[i, 1, g] * jinv[1])
            grad[:, 0] += corners[i] * ddx
            grad[:, 1] += corners[i] * ddy
        eps = dtype(0.5) * (grad + np.transpose(grad, (1, 0, 2, 3)))
        div = grad[0, 0] + grad[1, 1]
        sig = dtype(2 * mu) * eps
        sig[0, 0] = dtype(lam) * div + sig[0, 0]
        sig[1, 1] = dtype(lam) * div + sig[1, 1]
        val = np.sum(sig * eps, axis=(0, 1), dtype=dtype)
        total = val if total is None else total + val
    return total * dtype(detj)


def penalize(rho, p, m=1e-6):
    return m + rho ** p * (1 - m)


def penalize_derivative(rho, p, m=1e-6):
    return p * rho ** (p - 1) * (1 - m)


def objective_and_gradient(u, shape, density, dxdy, lam, mu, p=3.0, dtype=np.float32):
    """DEM_src/elasisity_problem.py:118-129."""
    e = strain_energy_density(u, shape, dxdy, lam, mu, dtype)
    rho = np.asarray(density, dtype=dtype)
    r = (dtype(1e-6) + rho ** dtype(p) * dtype(1 - 1e-6)).astype(dtype)
    dr = (dtype(p) * rho ** dtype(p - 1) * dtype(1 - 1e-6)).astype(dtype)
    objective = float(np.sum((r * e).astype(np.float64)))
    return objective, (-dr * e).astype(dtype)


def internal_energy(u, shape, density, dxdy, lam, mu, p=3.0, dtype=np.float64):
    """1/2 sum r(rho) e   (the first term of DEM_src/elasisity_problem.py:99-103)."""
    e = strain_energy_density(u, shape, dxdy, lam, mu, dtype)
    rho = np.asarray(density, dtype=dtype)
    return 0.5 * float(np.sum(penalize(rho, p) * e))


def energy_gradient_u(u, shape, density, dxdy, lam, mu, p=3.0):
    """d internal_energy / d u in the flattened layout of u (float64, analytic)."""
    u = np.asarray(u, dtype=np.float64)
    ny1, nx1 = shape
    dx, dy = dxdy
    jinv = np.array([1.0 / (dx / 2), 1.0 / (dy / 2)])  # np.linalg.inv of the diagonal Jacobian
    detj = (dx / 2) * (dy / 2)
    corners = node_values(u, shape)
    r = penalize(np.asarray(density, dtype=np.float64), p)
    G = np.zeros((ny1, nx1, 2))
    offs = [(0, 0), (1, 0), (0, 1), (1, 1)]
    for g in range(4):
        d = SHAPE_DERIVATIVES[:, :, g] * jinv[None, :]  # (node, 2)
        grad = np.zeros((2, 2) + corners[0].shape[1:])
        for i in range(4):
            grad[:, 0] += corners[i] * d[i, 0]
            grad[:, 1] += corners[i] * d[i, 1]
        eps = 0.5 * (grad + np.transpose(grad, (1, 0, 2, 3)))
        div = grad[0, 0] + grad[1, 1]
        sig = 2 * mu * eps
        sig[0, 0] += lam * div
        sig[1, 1] += lam * div
        # d(1/2 sigma:eps)/d grad = sigma
        for i, (oy, ox) in enumerate(offs):
            for c in range(2):
                contrib = r * detj * (sig[c, 0] * d[i, 0] + sig[c, 1] * d[i, 1])
                G[oy:oy + ny1 - 1, ox:ox + nx1 - 1, c] += contrib
    return np.transpose(G, (1, 0, 2)).reshape(-1, 2)
