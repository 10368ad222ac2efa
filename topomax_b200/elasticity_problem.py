"""Elastic compliance problem on the GPU engine (reference: FEM_src/elasisity_problem.py:76-196,
FEM_src/problem.py:13-36).

Same public surface as the reference class -- ``filter``, ``u``, ``filtered_rho``, ``penalizer``,
``lamé_mu``, ``lamé_lda``, ``body_force``, ``traction_term``, ``domain_size``,
``set_penalization`` / ``calculate_objective`` / ``calculate_objective_gradient`` / ``forward`` --
with dolfin's assemble + MUMPS replaced by the kernels behind ``Engine``.
"""
from __future__ import annotations

import torch  # noqa: F401  (tensor handoff only)

from .designs.definitions import DomainParameters, ElasticityParameters
from .engine import Engine
from .filter import AssembledP1Form, HelmholtzFilter
from .mesh import Function, FunctionSpace, RectangleMesh
from .penalizers import ElasticPenalizer
from .problem import Problem


class ElasticityProblem(Problem):
    """Elastic compliance topology optimization problem."""

    def __init__(self, mesh: RectangleMesh, control_space: FunctionSpace,
                 domain_parameters: DomainParameters, elasticity_parameters: ElasticityParameters,
                 *, state_rtol: float = 1e-10, state_max_iterations: int = 200000,
                 filter_rtol: float = 1e-11, preconditioner: str = "multigrid",
                 warm_start: bool = True, engine: Engine | None = None, mixed_precision: bool = False,
                 warm_start_extrapolation: bool = False, multigrid_cycle="auto",
                 attainable_accuracy_stop: float | None = None):
        self.parameters = elasticity_parameters
        self.domain_size = (domain_parameters.width, domain_parameters.height)
        self.mesh = mesh
        self.control_space = control_space

        self.Young_modulus = self.parameters.young_modulus
        self.Poisson_ratio = self.parameters.poisson_ratio
        self.penalizer = ElasticPenalizer()
        # Lamé parameters exactly as the reference computes them (plane-strain-type lambda)
        self.lamé_mu = self.Young_modulus / (2 * (1 + self.Poisson_ratio))
        self.lamé_lda = self.lamé_mu * self.Poisson_ratio / (0.5 - self.Poisson_ratio)

        # a sharded FEMSolver builds the engine first (it needs the partition to size its arrays)
        self.engine = engine if engine is not None else Engine(
            mesh.nx, mesh.ny, mesh.width, mesh.height,
            lame_lambda=self.lamé_lda, lame_mu=self.lamé_mu, simp_min=self.penalizer.minimum,
            filter_radius=self.parameters.filter_radius, fixed_sides=self.parameters.fixed_sides,
            dtype=control_space.dtype_name, device=control_space.device,
        )
        from . import _lib
        self.engine.set_option(_lib.OPT_PRECOND, _lib.PRECOND_MULTIGRID if preconditioner == "multigrid"
                               else _lib.PRECOND_JACOBI)
        self.preconditioner = preconditioner
        # optional: fp32 multigrid preconditioner inside the fp64 PCG (solution accuracy unchanged)
        self.mixed_precision = mixed_precision
        if mixed_precision:
            self.engine.set_option(117, 1)
        # shape of the multigrid cycle: "auto" (the library's window of small levels cycled twice), "v" (plain
        # V-cycle) or (first_level, last_level, cycles); None / "auto" leaves the library default untouched
        if multigrid_cycle not in (None, "auto"):
            if multigrid_cycle == "v":
                self.engine.set_option(_lib.OPT_CYCLE_GAMMA, 1)
            else:
                first, last, cycles = multigrid_cycle
                self.engine.set_option(_lib.OPT_CYCLE_FIRST, int(first))
                self.engine.set_option(_lib.OPT_CYCLE_LAST, int(last))
                self.engine.set_option(_lib.OPT_CYCLE_GAMMA, int(cycles))
        self.multigrid_cycle = multigrid_cycle
        # warm-started solves stop at max(state_rtol, factor x the residual level fp64 cannot resolve on this
        # mesh); None keeps the library default (0.5), 0 iterates until the recursive residual meets state_rtol
        if attainable_accuracy_stop is not None:
            self.engine.set_option(_lib.OPT_FP_FLOOR_FACTOR, float(attainable_accuracy_stop))
        self.state_rtol = state_rtol
        self.state_max_iterations = state_max_iterations
        self.warm_start = warm_start
        # optional: initial guess 2 u_k - u_{k-1} (linear extrapolation over the last two designs) instead of
        # u_k; costs one more lattice vector; the engine still keeps whichever of {guess, 0} has the smaller
        # residual, and the converged displacement meets the same tolerance
        self.warm_start_extrapolation = warm_start_extrapolation
        self._u_before: "torch.Tensor | None" = None

        self.solution_space = FunctionSpace(mesh, "CG", 2, dtype=control_space.dtype_name,
                                            device=self.engine.device, local_rows=control_space.local_rows)
        self.body_force = self.parameters.body_force
        self.traction_term = self.parameters.tractions or []
        # load vector, assembled once like SmartMumpsSolver(l_has_no_args=True)
        self.load = self.engine.load_vector(self.body_force, self.traction_term)

        self.filter = HelmholtzFilter(self.parameters.filter_radius, control_space,
                                      engine=self.engine, rtol=filter_rtol)
        self.u: Function | None = None
        self.filtered_rho: Function | None = None
        self.solve_log: list[dict] = []

    # ------------------------------------------------------------------ Problem interface
    def set_penalization(self, penalization: float):
        if self.penalizer is None:
            raise ValueError("Classes deriving from Problem must set a penalizer in their initializer")
        self.penalizer.set_penalization(penalization)

    def forward(self, rho: Function) -> Function:
        """State solve  K(rho) u = b, u = 0 on the fixed sides; ``rho`` is the filtered density."""
        p = self.penalizer.assert_has_penalization()
        warm = self.warm_start and self.u is not None
        # warm start: the previous displacement is the initial guess and is overwritten IN PLACE by the new
        # one (no clone: a lattice vector is 51 GB at N=16384); self.u is replaced by the result below
        u0 = self.u.tensor if warm else None
        if warm and self.warm_start_extrapolation:
            previous = u0.clone()
            if self._u_before is not None:
                u0.mul_(2.0).sub_(self._u_before)
            self._u_before = previous
        u, info = self.engine.state_solve(rho.tensor, self.load, p, rtol=self.state_rtol,
                                          maxit=self.state_max_iterations, u=u0, warm_start=warm)
        stats = self.engine.last_solve_stats()
        stats["relative_residual"] = info.relative_residual
        self.solve_log.append(stats)
        return Function(self.solution_space, u)

    def calculate_objective(self, rho: Function) -> float:
        """Compliance  phi(rho) = int u.f dx + int u.t ds = u . b."""
        self.filtered_rho = self.filter.apply(rho)
        self.u = self.forward(self.filtered_rho)
        return float(self.engine.dot_p2(self.u.tensor, self.load))

    def calculate_objective_gradient(self) -> Function:
        """Filtered  -r'(xi) (lambda |div u|^2 + 2 mu |eps(u)|^2)."""
        if self.filtered_rho is None or self.u is None:
            raise ValueError(
                "You must call calculate_objective before calling calculate_objective_gradient"
            )
        p = self.penalizer.assert_has_penalization()
        rhs = self.engine.sens_rhs(self.filtered_rho.tensor, self.u.tensor, p)
        return self.filter.apply(AssembledP1Form(self.control_space, rhs))
