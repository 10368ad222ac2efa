"""Helmholtz density filter  -eps^2 lap(xi) + xi = rho,  natural boundary conditions
(reference: FEM_src/filter.py:8-41).  The reference factorises eps^2 K1 + M1 with MUMPS on
every call; here it is a matrix-free Jacobi-PCG on the vertex grid (csrc/tm_p1.cuh)."""
from __future__ import annotations

import torch

from .engine import Engine
from .mesh import Function, FunctionSpace


class AssembledP1Form:
    """An already assembled P1 right-hand side  b_i = int g phi_i  (what the reference gets by
    passing a UFL expression to ``HelmholtzFilter.apply``)."""

    def __init__(self, space: FunctionSpace, tensor: torch.Tensor):
        self.space = space
        self.tensor = tensor


class HelmholtzFilter:
    def __init__(self, epsilon: float, function_space: FunctionSpace, *, engine: Engine | None = None,
                 rtol: float = 1e-11, max_iterations: int = 50000):
        if function_space.degree != 1:
            raise ValueError("the filter acts on the P1 control space")
        self.epsilon = float(epsilon)
        self.function_space = function_space
        self.rtol = rtol
        self.max_iterations = max_iterations
        if engine is None:
            mesh = function_space.mesh()
            engine = Engine(mesh.nx, mesh.ny, mesh.width, mesh.height, filter_radius=self.epsilon,
                            dtype=function_space.dtype_name, device=function_space.device)
        self.engine = engine
        self.last_info = None

    def apply(self, input_function) -> Function:
        if isinstance(input_function, AssembledP1Form):
            out, info = self.engine.filter_apply(input_function.tensor, assembled=True,
                                                 rtol=self.rtol, maxit=self.max_iterations)
        elif isinstance(input_function, Function):
            out, info = self.engine.filter_apply(input_function.tensor, assembled=False,
                                                 rtol=self.rtol, maxit=self.max_iterations)
        else:
            raise TypeError("HelmholtzFilter.apply expects a Function or an AssembledP1Form")
        self.last_info = info
        return Function(self.function_space, out)
