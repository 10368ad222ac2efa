"""Q1 strain-energy evaluator of the deep-energy back-end on the CUDA library.

Mirrors the reference's ``StrainEnergy`` (reference: DEM_src/elasisity_problem.py:64-129) over
``ObjectiveCalculator`` (DEM_src/objective_calculator.py:15-143) and its ``Mesh``
(DEM_src/utils.py:11-34): same constructor arguments, ``set_penalization``,
``calculate_objective_and_gradient(u, shape, density)`` and ``calculate_energy(u, shape, density)``,
same array layouts (``u`` flattened with x as the slow index, ``density`` as ``(Ny, Nx)``), float32.
One fused kernel computes the Gauss-point strains, sigma:eps, the SIMP weights, the density
gradient and the objective; ``calculate_energy`` is differentiable with respect to ``u`` and
``density`` through a hand-written backward kernel instead of autograd through ~200 small tensor
ops.  Only the internal energy is evaluated here: the boundary/body-force integrals of the
reference (``calculate_traction_integral``, ``calculate_body_force_integral``) stay outside the
scope, so ``body_force`` and ``traction_points_list`` must be None.

CUDA tensors only: there is no CPU path (``_lib.load_library`` raises without the built library).
"""
from __future__ import annotations

from ctypes import byref, c_double, c_void_p

import numpy as np
import torch

from . import _lib
from .penalizers import ElasticPenalizer


class Mesh:
    """reference: DEM_src/utils.py:11-34."""

    def __init__(self, Nx: int, Ny: int, width: float, height: float):
        self.Nx, self.Ny = int(Nx), int(Ny)
        self.width, self.height = float(width), float(height)
        x_ray = np.linspace(0.0, self.width, self.Nx + 1)
        y_ray = np.linspace(0.0, self.height, self.Ny + 1)
        self.x_grid, self.y_grid = np.meshgrid(x_ray, y_ray)
        self.intervals = (self.Ny, self.Nx)
        self.shape = (self.Ny + 1, self.Nx + 1)
        self.dxdy = (self.width / self.Nx, self.height / self.Ny)


def _evaluate(mesh: Mesh, lam: float, mu: float, minimum: float, penalty: float, u: torch.Tensor,
              density: torch.Tensor, want_cells: bool, want_grad_density: bool, want_grad_u: bool):
    if not (u.is_cuda and density.is_cuda):
        raise RuntimeError("topomax_b200.dem_energy runs on CUDA tensors only (no CPU fallback)")
    nodes, cells = (mesh.Nx + 1) * (mesh.Ny + 1), mesh.Nx * mesh.Ny
    u32 = u.detach().to(torch.float32).contiguous()
    rho32 = density.detach().to(torch.float32).contiguous()
    if u32.numel() != 2 * nodes or rho32.numel() != cells:
        raise ValueError(f"expected u with {nodes} x 2 entries and density with {cells}, got "
                         f"{tuple(u.shape)} and {tuple(density.shape)}")
    lib = _lib.load_library()
    mk = lambda n, want: torch.empty(n, dtype=torch.float32, device=u32.device) if want else None
    cell, gd, gu = mk(cells, want_cells), mk(cells, want_grad_density), mk(2 * nodes, want_grad_u)
    ptr = lambda t: c_void_p(t.data_ptr()) if t is not None else c_void_p(0)
    objective = c_double(0.0)
    with torch.cuda.device(u32.device):
        stream = torch.cuda.current_stream(u32.device).cuda_stream
        _lib.check(lib.tm_dem_strain_energy(
            mesh.Nx, mesh.Ny, mesh.width, mesh.height, lam, mu, minimum, penalty,
            ptr(u32), ptr(rho32), ptr(cell), ptr(gd), ptr(gu), byref(objective), c_void_p(stream)))
    shape2 = lambda t: t.reshape(mesh.Ny, mesh.Nx) if t is not None else None
    return objective.value, shape2(cell), shape2(gd), (gu.reshape(nodes, 2) if gu is not None else None)


class _InternalEnergy(torch.autograd.Function):
    """psi_int(u; rho) = 1/2 sum r(rho) e(u); backward from the library's gather kernel."""

    @staticmethod
    def forward(ctx, u, density, calc):
        need_u, need_rho = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        obj, _, gd, gu = _evaluate(calc.mesh, calc.lamé_lda, calc.lamé_mu, calc.penalizer.minimum,
                                   calc.penalizer.assert_has_penalization(), u, density, False, need_rho, need_u)
        ctx.save_for_backward(gu if need_u else None, gd if need_rho else None)
        ctx.shapes = (u.shape, density.shape, u.dtype, density.dtype)
        return torch.tensor(0.5 * obj, dtype=torch.float32, device=u.device)

    @staticmethod
    def backward(ctx, grad_out):
        gu, gd = ctx.saved_tensors
        ushape, rshape, udt, rdt = ctx.shapes
        du = (grad_out * gu).reshape(ushape).to(udt) if gu is not None else None
        # d(1/2 r e)/d rho = 1/2 r'(rho) e = -1/2 * (the objective's density gradient)
        drho = (grad_out * (-0.5) * gd).reshape(rshape).to(rdt) if gd is not None else None
        return du, drho, None


class StrainEnergy:
    """reference: DEM_src/elasisity_problem.py:64-129."""

    def __init__(self, mesh: Mesh, body_force, Young_modulus: float, Poisson_ratio: float, traction_points_list):
        if body_force is not None or traction_points_list is not None:
            raise NotImplementedError(
                "the external-energy integrals of the deep-energy back-end are outside the scope of "
                "topomax_b200 (SURVEY.md section 2, #14): pass None and add them on the torch side")
        self.mesh = mesh
        self.penalizer = ElasticPenalizer()
        self.lamé_mu = Young_modulus / (2 * (1 + Poisson_ratio))
        self.lamé_lda = self.lamé_mu * Poisson_ratio / (0.5 - Poisson_ratio)
        dx, dy = mesh.dxdy
        self.detJ = (dx / 2) * (dy / 2)

    def set_penalization(self, penalization: float):
        self.penalizer.set_penalization(penalization)

    def _check_shape(self, shape):
        if tuple(shape) != tuple(self.mesh.shape):
            raise ValueError(f"shape {tuple(shape)} does not match the mesh {self.mesh.shape}")

    def strain_energy_at_element(self, u: torch.Tensor, shape) -> torch.Tensor:
        """det J * sum over the Gauss points of sigma:eps, ``(Ny, Nx)``."""
        self._check_shape(shape)
        penalty = self.penalizer.penalization if self.penalizer.penalization is not None else 1.0
        dummy = torch.ones(self.mesh.intervals, dtype=torch.float32, device=u.device)
        return _evaluate(self.mesh, self.lamé_lda, self.lamé_mu, self.penalizer.minimum, penalty, u, dummy,
                         True, False, False)[1]

    def calculate_energy(self, u: torch.Tensor, shape, density: torch.Tensor) -> torch.Tensor:
        """psi(u; rho) = 1/2 sum r(rho) sigma:eps det J (internal part), differentiable."""
        self._check_shape(shape)
        return _InternalEnergy.apply(u, density, self)

    def calculate_objective_and_gradient(self, u: torch.Tensor, shape, density: torch.Tensor):
        """phi(rho) = sum r(rho) e  and  grad phi = -r'(rho) e  (reference :118-129)."""
        self._check_shape(shape)
        obj, _, gd, _ = _evaluate(self.mesh, self.lamé_lda, self.lamé_mu, self.penalizer.minimum,
                                  self.penalizer.assert_has_penalization(), u, density, False, True, False)
        return torch.tensor(obj, dtype=torch.float32, device=u.device), gd.reshape(density.shape)
