"""SURVEY.md 8f-4 on the GPU: ``tm_dem_strain_energy`` through the C ABI and the ``StrainEnergy``
mirror, against the reference's fixtures, the reference code's own outputs on random fields
(tests/golden/dem_strain_energy_reference.json) and the numpy oracle.  float32 path; tolerances as in
tests/test_dem_oracle.py: per-cell values 4 ulps of the field maximum, sums 1e-6, displacement
gradient 2e-6 of its maximum."""
import json
import os

import numpy as np
import pytest

from oracle import dem_oracle as do

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
ULP = float(np.finfo(np.float32).eps)


@pytest.fixture(scope="module")
def vectors(golden_dir):
    return json.load(open(os.path.join(golden_dir, "dem_strain_energy_reference.json")))


def make(case):
    from DEM_src.elasisity_problem import StrainEnergy
    from DEM_src.utils import Mesh

    mesh = Mesh(case["Nx"], case["Ny"], case["width"], case["height"])
    se = StrainEnergy(mesh, None, case["young_modulus"], case["poisson_ratio"], None)
    se.set_penalization(case["penalty"])
    return mesh, se


def test_reference_fixtures(vectors):
    """reference tests/test_DEM_problem.py:16-47 (u = the node coordinates, uniform density)."""
    for fx in vectors["fixtures"]:
        mesh, se = make(fx)
        x = torch.from_numpy(np.array([mesh.x_grid.T.flat, mesh.y_grid.T.flat]).T).float().cuda()
        rho = torch.full(mesh.intervals, fx["volume_fraction"], dtype=torch.float32, device="cuda")
        obj, grad = se.calculate_objective_and_gradient(x, mesh.shape, rho)
        ref = np.array(fx["gradient"], dtype=np.float32)
        assert grad.shape == ref.shape
        assert np.abs(grad.cpu().numpy() - ref).max() <= 2 * ULP * np.abs(ref).max()
        assert abs(float(obj) - fx["objective"]) < 1e-6 * fx["objective"]


def test_random_fields_match_the_reference_code(vectors):
    for case in vectors["random_cases"]:
        mesh, se = make(case)
        u = torch.tensor(case["u"], dtype=torch.float32, device="cuda", requires_grad=True)
        rho = torch.tensor(case["density"], dtype=torch.float32, device="cuda", requires_grad=True)
        obj, grad = se.calculate_objective_and_gradient(u, mesh.shape, rho)
        ref = np.array(case["gradient"], np.float32)
        assert np.abs(grad.cpu().numpy() - ref).max() <= 4 * ULP * np.abs(ref).max()
        assert abs(float(obj) - case["objective"]) < 1e-6 * abs(case["objective"])
        energy = se.calculate_energy(u, mesh.shape, rho)
        assert abs(float(energy) - case["energy"]) < 1e-6 * abs(case["energy"])
        (3.0 * energy).backward()
        ref_u = 3.0 * np.array(case["energy_gradient_u"])
        assert np.abs(u.grad.cpu().numpy() - ref_u).max() < 2e-6 * np.abs(ref_u).max()
        # d energy / d rho = -1/2 of the objective's density gradient
        assert np.abs(rho.grad.cpu().numpy() + 1.5 * ref).max() <= 8 * ULP * np.abs(ref).max() * 1.5


@pytest.mark.parametrize("nx,ny", [(1, 1), (257, 3), (96, 200), (640, 320)])
def test_against_the_oracle_at_other_sizes(nx, ny):
    """ragged sizes, a single cell, and the largest mesh the reference's scripts use (N = 320)."""
    rng = np.random.default_rng(nx * 1000 + ny)
    W, H, E, nu, p = 2.0, 1.0, 3.0, 0.3, 3.0
    from DEM_src.elasisity_problem import StrainEnergy
    from DEM_src.utils import Mesh

    mesh = Mesh(nx, ny, W, H)
    se = StrainEnergy(mesh, None, E, nu, None)
    se.set_penalization(p)
    u = rng.standard_normal(((nx + 1) * (ny + 1), 2)).astype(np.float32)
    rho = (0.05 + 0.9 * rng.random((ny, nx))).astype(np.float32)
    lam, mu = do.lame(E, nu)
    obj_o, grad_o = do.objective_and_gradient(u, mesh.shape, rho, mesh.dxdy, lam, mu, p)
    ut = torch.from_numpy(u).cuda().requires_grad_(True)
    obj, grad = se.calculate_objective_and_gradient(ut, mesh.shape, torch.from_numpy(rho).cuda())
    assert np.abs(grad.cpu().numpy() - grad_o).max() <= 4 * ULP * np.abs(grad_o).max()
    assert abs(float(obj) - obj_o) < 2e-6 * abs(obj_o)
    e = se.strain_energy_at_element(ut, mesh.shape).cpu().numpy()
    e_o = do.strain_energy_density(u, mesh.shape, mesh.dxdy, lam, mu)
    assert np.abs(e - e_o).max() <= 4 * ULP * np.abs(e_o).max()
    se.calculate_energy(ut, mesh.shape, torch.from_numpy(rho).cuda()).backward()
    gu_o = do.energy_gradient_u(u, mesh.shape, rho, mesh.dxdy, lam, mu, p)
    assert np.abs(ut.grad.cpu().numpy() - gu_o).max() < 5e-6 * np.abs(gu_o).max()


def test_errors_like_the_reference():
    from DEM_src.elasisity_problem import StrainEnergy
    from DEM_src.utils import Mesh

    mesh = Mesh(4, 3, 1.0, 1.0)
    se = StrainEnergy(mesh, None, 1.0, 0.3, None)
    u = torch.zeros((20, 2), device="cuda")
    rho = torch.ones((3, 4), device="cuda")
    with pytest.raises(ValueError):      # src/penalizers.py:14-21: penalisation not set
        se.calculate_objective_and_gradient(u, mesh.shape, rho)
    se.set_penalization(3.0)
    with pytest.raises(RuntimeError):    # CUDA tensors only
        se.calculate_objective_and_gradient(u.cpu(), mesh.shape, rho.cpu())
    with pytest.raises(ValueError):
        se.calculate_objective_and_gradient(u[:-1], mesh.shape, rho)
    obj, grad = se.calculate_objective_and_gradient(u, mesh.shape, rho)
    assert float(obj) == 0.0 and float(grad.abs().max()) == 0.0
