"""Fluid topology-optimisation problem on the GPU (reference: FEM_src/fluid_problem.py:12-155,
FEM_src/problem.py:13-36, src/penalizers.py:49-68) -- SURVEY.md 8f-3.

Same public surface as the reference class -- ``FluidProblem(mesh, fluid_parameters,
domain_parameters)``, ``viscosity``, ``penalizer``, ``u``, ``rho``, ``boundary_flows``,
``set_penalization`` / ``calculate_objective`` / ``calculate_objective_gradient`` / ``forward`` --
with dolfin's assemble + MUMPS replaced by ``tm_fluid_*`` (Taylor-Hood Stokes-Brinkman operator,
preconditioned MINRES) and the L2 projection of the sensitivity by the engine's P1 mass solve.

The reference's linear system is singular (pressure constant) and, when the discrete boundary flux
does not vanish, inconsistent; its answer then depends on MUMPS' round-off (DESIGN.md section 7,
"8f-3").  Here the continuity right-hand side is projected onto the range (the
least-squares solution), which equals the reference's velocity whenever the system is consistent.
"""
from __future__ import annotations

import ctypes
from ctypes import byref, c_double, c_int, c_void_p

import numpy as np
import torch

from . import _lib
from .designs.definitions import DomainParameters, FluidParameters, Side
from .engine import Engine
from .mesh import Function, FunctionSpace, RectangleMesh
from .penalizers import Penalizer
from .problem import Problem


class FluidPenalizer(Penalizer):
    """reference: src/penalizers.py:49-68."""

    def __init__(self):
        super().__init__()
        self.minimum = 2.5 / 100**2
        self.maximum = 2.5 / 0.01**2

    def __call__(self, rho):
        q, mini, maxi = self.assert_has_penalization(), self.minimum, self.maximum
        return maxi + (mini - maxi) * rho * (1 + q) / (rho + q)

    def derivative(self, rho):
        q, mini, maxi = self.assert_has_penalization(), self.minimum, self.maximum
        return (mini - maxi) * q * (1 + q) / (rho + q) ** 2


class BoundaryFlows:
    """Nodal values of the reference's ``BoundaryFlows`` expression (FEM_src/fluid_problem.py:12-45)
    on the P2 lattice, combined the way its two ``DirichletBC`` objects are applied (:127-148):
    parabolic profiles on the flow sides, then no-slip on the other sides (which therefore win at
    the corners).  Coordinates follow dolfin's mesh generator: vertex ``(i*W)/nx``, edge midpoints
    the mean of their end vertices; the side tests are the reference's exact comparisons."""

    def __init__(self, domain_size, flows, mesh: RectangleMesh):
        self.domain_size = domain_size
        self.flows = flows
        self.mesh = mesh

    @staticmethod
    def get_flow(position, center, length, rate):
        t = position - center
        return np.where((-length / 2 < t) & (t < length / 2), rate * (1 - (2 * t / length) ** 2), 0.0)

    def lattice_coordinates(self):
        m = self.mesh

        def axis(n, size):
            v = (np.arange(n + 1, dtype=np.float64) * size) / n
            out = np.empty(2 * n + 1)
            out[0::2] = v
            out[1::2] = 0.5 * v[:-1] + 0.5 * v[1:]
            return out

        return np.meshgrid(axis(m.nx, m.width), axis(m.ny, m.height), indexing="xy")  # [Ly, Lx] each

    def nodal_values(self) -> np.ndarray:
        """[Ly, Lx, 2] array: prescribed velocity on boundary nodes, zero inside."""
        X, Y = self.lattice_coordinates()
        W, H = self.domain_size
        ux, uy = np.zeros_like(X), np.zeros_like(X)
        on_side = {Side.LEFT: X == 0.0, Side.RIGHT: X == W, Side.TOP: Y == H, Side.BOTTOM: Y == 0}
        flow_sides = set()
        for flow in self.flows:
            side, center, length, rate = flow.to_tuple()
            flow_sides.add(side)
            if side == Side.LEFT:
                ux += np.where(on_side[side], self.get_flow(Y, center, length, rate), 0.0)
            elif side == Side.RIGHT:
                ux -= np.where(on_side[side], self.get_flow(Y, center, length, rate), 0.0)
            elif side == Side.TOP:
                uy -= np.where(on_side[side], self.get_flow(X, center, length, rate), 0.0)
            elif side == Side.BOTTOM:
                uy += np.where(on_side[side], self.get_flow(X, center, length, rate), 0.0)
            else:
                raise ValueError(f"Malformed side: {side}")
        flow_mask = np.zeros_like(X, dtype=bool)
        for side in flow_sides:
            flow_mask |= on_side[side]
        no_slip = np.zeros_like(X, dtype=bool)
        for side in set(Side.get_all()).difference(flow_sides):
            no_slip |= on_side[side]
        keep = flow_mask & ~no_slip
        return np.stack([np.where(keep, ux, 0.0), np.where(keep, uy, 0.0)], axis=-1)


class FluidProblem(Problem):
    """Fluid power-dissipation topology optimization problem."""

    def __init__(self, mesh: RectangleMesh, fluid_parameters: FluidParameters,
                 domain_parameters: DomainParameters, *, control_space: FunctionSpace | None = None,
                 state_rtol: float = 1e-10, state_max_iterations: int = 200000, projection_rtol: float = 1e-12,
                 preconditioner: str = "auto", warm_start: bool = False,
                 device_scalars: bool = False, deterministic: bool | None = None, graph: bool | None = None,
                 device=None):
        """``preconditioner``: "diagonal", "multigrid" (V-cycles on per-triangle Galerkin matrices for the
        velocity block and the pressure's Darcy Laplacian; needs cell counts with enough factors of two) or
        "auto" (default): multigrid, its MINRES iterations replayed from a CUDA graph, on meshes with at least
        64 cells on the short side that can be coarsened far enough -- measured on the B200
        (profiles/r2d_fluid_bench.txt): diffuser N=128 104 ms against 403 ms per mirror-descent iteration
        (126 against ~5900 MINRES iterations), N=256 144 ms against 1211 ms -- and the diagonal preconditioner
        on small meshes, which stay launch bound.  ``graph`` None follows the preconditioner choice.
        ``deterministic`` None: the gather kernels (no atomics, bit-reproducible) on meshes below 64 cells per
        side, where they cost nothing (launch bound), the scatter-with-atomics kernels above (1.5x faster at
        N=256: 144 against 223 ms per iteration, summation order then varies at round-off level)."""
        self.parameters = fluid_parameters
        self.mesh = mesh
        self.domain_size = (domain_parameters.width, domain_parameters.height)
        self.viscosity = self.parameters.viscosity
        self.penalizer: FluidPenalizer = FluidPenalizer()
        self.state_rtol, self.state_max_iterations = state_rtol, state_max_iterations
        self.projection_rtol = projection_rtol

        if device is None and control_space is not None:
            device = control_space.device
        if device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("FluidProblem needs a CUDA device: topomax_b200 has no CPU fallback")
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        if control_space is None:
            control_space = FunctionSpace(mesh, "CG", 1, dtype="float64", device=self.device)
        if control_space.dtype_name != "float64":
            raise ValueError("the fluid path computes in float64")
        self.control_space = control_space
        self.lib = _lib.load_library()
        # P1 services (mass solve of the L2 projection, integrals, the mirror-descent kernels) come
        # from the elasticity engine on the same mesh; its material constants are not used here
        self.engine = Engine(mesh.nx, mesh.ny, mesh.width, mesh.height, lame_lambda=1.0, lame_mu=1.0,
                             simp_min=1e-6, filter_radius=0.0, fixed_sides=[], dtype="float64",
                             device=self.device)
        handle = c_void_p()
        _lib.check(self.lib.tm_fluid_create(mesh.nx, mesh.ny, mesh.width, mesh.height, self.viscosity,
                                            self.penalizer.minimum, self.penalizer.maximum,
                                            self.device.index or 0, byref(handle)))
        self._h = handle
        if preconditioner not in ("diagonal", "multigrid", "auto"):
            raise ValueError(f"preconditioner must be 'diagonal', 'multigrid' or 'auto', got {preconditioner!r}")
        self.auto_preconditioner = preconditioner == "auto"
        if self.auto_preconditioner:
            preconditioner = "multigrid" if self.multigrid_pays_off(mesh.nx, mesh.ny) else "diagonal"
            if graph is None:
                graph = preconditioner == "multigrid"
        self.preconditioner = preconditioner
        if preconditioner == "multigrid":
            _lib.check(self.lib.tm_fluid_set_option(self._h, 1, 1.0))
        self.warm_start = bool(warm_start)
        if self.warm_start:
            _lib.check(self.lib.tm_fluid_set_option(self._h, 4, 1.0))
        self.device_scalars = bool(device_scalars)  # MINRES recurrences on the device, no sync per iteration
        if self.device_scalars:
            _lib.check(self.lib.tm_fluid_set_option(self._h, 5, 1.0))
        if deterministic is None:
            deterministic = min(mesh.nx, mesh.ny) < 64
        self.deterministic = bool(deterministic)  # gather kernels instead of scatter + atomics
        if self.deterministic:
            _lib.check(self.lib.tm_fluid_set_option(self._h, 7, 1.0))
        self.graph = bool(graph)  # six MINRES iterations replayed from a captured CUDA graph
        if self.graph:
            _lib.check(self.lib.tm_fluid_set_option(self._h, 8, 1.0))
        self.solution_space = FunctionSpace(mesh, "CG", 2, dtype="float64", device=self.device)
        self.n1 = (mesh.nx + 1) * (mesh.ny + 1)
        self.nu = 2 * (2 * mesh.nx + 1) * (2 * mesh.ny + 1)

        self.boundary_flows = BoundaryFlows(self.domain_size, self.parameters.flows, mesh)
        self.boundary_velocity = torch.as_tensor(self.boundary_flows.nodal_values().reshape(-1),
                                                 dtype=torch.float64).to(self.device)
        self.u: Function | None = None
        self.p: torch.Tensor | None = None
        self.rho: Function | None = None
        self.solve_log: list[dict] = []

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                self.lib.tm_fluid_destroy(h)
            except Exception:
                pass
            self._h = None

    def _sync_stream(self):
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.tm_fluid_set_stream(self._h, c_void_p(stream)))

    @staticmethod
    def _ptr(t: torch.Tensor, n: int):
        if t.dtype != torch.float64 or not t.is_contiguous() or t.numel() != n or not t.is_cuda:
            raise ValueError(f"expected a contiguous float64 CUDA tensor with {n} entries, got "
                             f"{tuple(t.shape)} {t.dtype} on {t.device}")
        return c_void_p(t.data_ptr())

    # ------------------------------------------------------------------ Problem interface
    def set_penalization(self, penalization: float):
        if self.penalizer is None:
            raise ValueError("Classes deriving from Problem must set a penalizer in their initializer")
        self.penalizer.set_penalization(penalization)

    def set_density(self, rho: Function):
        """Weighted mass matrices + preconditioner for ``rho`` at the current penalisation."""
        q = self.penalizer.assert_has_penalization()
        self._sync_stream()
        _lib.check(self.lib.tm_fluid_set_density(self._h, self._ptr(rho.tensor, self.n1), q))

    @staticmethod
    def multigrid_pays_off(nx: int, ny: int) -> bool:
        """Mirror of ``CudaTriMG::level_list``: the mesh halves while both cell counts are even and more than
        four cells remain; the coarsest velocity level must fit the dense inverse (2048 scalar nodes)."""
        if min(nx, ny) < 64:
            return False
        while nx % 2 == 0 and ny % 2 == 0 and nx * ny > 4:
            nx, ny = nx // 2, ny // 2
        return (2 * nx + 1) * (2 * ny + 1) <= 2048

    def forward(self, rho: Function) -> torch.Tensor:
        """State solve; returns the combined ``[velocity | pressure]`` device vector."""
        self.set_density(rho)
        up = torch.empty(self.nu + self.n1, dtype=torch.float64, device=self.device)
        iters, relres = c_int(0), c_double(0.0)
        try:
            _lib.check(self.lib.tm_fluid_state_solve(self._h, self._ptr(self.boundary_velocity, self.nu),
                                                     self.state_rtol, self.state_max_iterations,
                                                     self._ptr(up, self.nu + self.n1), byref(iters), byref(relres)))
        except _lib.NotConverged:
            if not (self.auto_preconditioner and self.preconditioner == "multigrid"):
                raise
            # the automatic choice must never cost a run: back to the diagonal preconditioner for good
            self.preconditioner = "diagonal"
            _lib.check(self.lib.tm_fluid_set_option(self._h, 8, 0.0))
            _lib.check(self.lib.tm_fluid_set_option(self._h, 1, 0.0))
            self.graph = False
            return self.forward(rho)
        self.solve_log.append({"iterations": iters.value, "relative_residual": relres.value})
        return up

    def calculate_objective(self, rho: Function) -> float:
        """phi(rho) = 1/2 int r(rho)|u|^2 + mu |grad u|^2 dx."""
        self.rho = rho
        up = self.forward(rho)
        self.u = Function(self.solution_space, up[: self.nu])
        self.p = up[self.nu:]
        out = c_double(0.0)
        _lib.check(self.lib.tm_fluid_objective(self._h, self._ptr(self.u.tensor, self.nu), byref(out)))
        return float(out.value)

    def calculate_objective_gradient(self) -> Function:
        """L2 projection onto the control space of 1/2 r'(rho)|u|^2."""
        if self.rho is None or self.u is None:
            raise ValueError(
                "You must call calculate_objective before calling calculate_objective_gradient"
            )
        self._sync_stream()
        rhs = torch.empty(self.n1, dtype=torch.float64, device=self.device)
        _lib.check(self.lib.tm_fluid_sens_rhs(self._h, self._ptr(self.rho.tensor, self.n1),
                                              self._ptr(self.u.tensor, self.nu), self._ptr(rhs, self.n1)))
        grad, _ = self.engine.filter_apply(rhs, assembled=True, rtol=self.projection_rtol)
        return Function(self.control_space, grad)

    # ------------------------------------------------------------------ for the parity tests
    def apply_operator(self, x: torch.Tensor, mode: int = 0) -> torch.Tensor:
        self._sync_stream()
        y = torch.empty_like(x)
        _lib.check(self.lib.tm_fluid_apply(self._h, self._ptr(x, self.nu + self.n1),
                                           self._ptr(y, self.nu + self.n1), mode))
        return y
