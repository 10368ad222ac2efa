"""The C + OpenMP CPU kernel baseline (oracle/c/elast_omp.c) against the numpy/scipy oracle it
restates (CPU only).  Operator: FEM_src/elasisity_problem.py:112-118 with the Dirichlet rows of
FEM_src/elasisity_problem.py:183-192; solve: the role of FEM_src/pde_solver.py:130-131."""
import numpy as np
import pytest

from oracle.fem_oracle import StructuredMesh, lame, solve_spd
from oracle.omp_kernels import OmpElasticity

LAM, MU = lame(5.0 / 3.0, 0.3)


@pytest.mark.parametrize("fixed", [["Left"], ["Left", "Right"], ["Bottom", "Top"], ["Top", "Right", "Left", "Bottom"]])
@pytest.mark.parametrize("shape", [(3.0, 1.0, 12, 5), (1.0, 2.0, 1, 1), (1.0, 1.0, 7, 8)])
def test_operator_and_diagonal_match_assembled_matrix(fixed, shape):
    W, H, nx, ny = shape
    mesh = StructuredMesh(W, H, nx, ny)
    rng = np.random.default_rng(nx * 100 + ny)
    xi = rng.uniform(0.0, 1.0, mesh.n1)
    x = rng.standard_normal(mesh.nu)
    K = mesh.elasticity_matrix(xi, LAM, MU)
    mask = mesh.dirichlet_mask(fixed)
    y = K @ np.where(mask, 0.0, x)
    y[mask] = x[mask]
    d = K.diagonal().copy()
    d[mask] = 1.0
    op = OmpElasticity(W, H, nx, ny, LAM, MU, fixed)
    assert np.abs(op.apply(xi, x) - y).max() <= 1e-13 * np.abs(y).max()
    assert np.abs(op.diagonal(xi) - d).max() <= 1e-13 * d.max()


def test_general_penalty_exponent():
    mesh = StructuredMesh(2.0, 1.0, 6, 3)
    rng = np.random.default_rng(5)
    xi = rng.uniform(0.1, 1.0, mesh.n1)
    x = rng.standard_normal(mesh.nu)
    K = mesh.elasticity_matrix(xi, LAM, MU, p=2.5, m=1e-3, nq=6)
    op = OmpElasticity(2.0, 1.0, 6, 3, LAM, MU, [], p=2.5, m=1e-3, nq=6)
    y = K @ x
    assert np.abs(op.apply(xi, x) - y).max() <= 1e-13 * np.abs(y).max()


def test_jacobi_pcg_matches_direct_solve_and_is_thread_independent_to_rounding():
    W, H, nx, ny = 3.0, 1.0, 12, 5
    mesh = StructuredMesh(W, H, nx, ny)
    rng = np.random.default_rng(1)
    xi = rng.uniform(0.05, 1.0, mesh.n1)
    b = mesh.load_vector(None, [("Right", 0.5, 0.2, 0.0, -1.0)])
    mask = mesh.dirichlet_mask(["Left"])
    u_ref = solve_spd(mesh.elasticity_matrix(xi, LAM, MU), np.where(mask, 0.0, b), free=~mask)
    op = OmpElasticity(W, H, nx, ny, LAM, MU, ["Left"])
    u, its, rel, sec = op.jacobi_pcg(xi, b, rtol=1e-12)
    assert 0 < its < 5000 and rel <= 1e-12 and sec >= 0.0
    assert np.abs(u - u_ref).max() <= 1e-9 * np.abs(u_ref).max()
    assert np.all(u[mask] == 0.0)
    # an iteration cap is honoured
    _, its2, rel2, _ = op.jacobi_pcg(xi, b, rtol=1e-12, maxit=7)
    assert its2 == 7 and rel2 > 1e-12
    # zero right-hand side: no iterations, zero solution
    u0, its0, rel0, _ = op.jacobi_pcg(xi, np.zeros(mesh.nu))
    assert its0 == 0 and rel0 == 0.0 and not u0.any()


def test_malformed_side_raises():
    with pytest.raises(ValueError, match="Malformed side"):
        OmpElasticity(1.0, 1.0, 2, 2, LAM, MU, ["Front"])
