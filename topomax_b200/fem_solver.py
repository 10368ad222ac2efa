"""``FEMSolver``: the reference's FEM back-end of the optimiser on the GPU engine
(reference: FEM_src/solver.py:15-89 driven by src/solver.py:208-302).

Two equivalent loops:
* ``solve()``          -- device resident: psi, rho, the gradient and the volume projection stay
                          in HBM; per iteration only scalars cross PCIe (rho is copied to the
                          host only when an iteration is saved).
* ``solve_generic()``  -- the reference's base-class loop over the numpy hooks
                          (``to_array`` / ``set_from_array`` / ``integrate``), for callers that
                          drive the hooks themselves.
"""
from __future__ import annotations

import os
import pickle

import numpy as np
import torch

from .designs.definitions import ElasticityParameters, FluidParameters
from .elasticity_problem import ElasticityProblem
from .mesh import Function, FunctionSpace, RectangleMesh
from .printer import Printer
from .solver import MAX_ITERATIONS, Solver, find_volume_shift
from .utils import Timer


def dolfin_p1_permutation(nx: int, ny: int) -> np.ndarray:
    """``perm[d]`` = row-major vertex index of dolfin's P1 dof ``d`` on ``RectangleMesh(nx, ny)`` (default
    "right" diagonal, serial): dofs sweep the diagonals from the top-left to the bottom-right corner,
    i.e. vertices sorted by (ix - iy, ix).  Verified against the reference's golden file
    tests/test_data/FEM/triangle/data/correct_rho.dat (11 x 11 vertices) -- the only
    reference-written design file there is; other mesh shapes follow the same rule unverified."""
    ix, iy = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), indexing="xy")
    ix, iy = ix.ravel(), iy.ravel()
    return np.lexsort((ix, ix - iy))


def pack_function_data(vector: np.ndarray, N: int, domain_size, problem: str, ordering: str = "row_major") -> dict:
    """The dict ``save_function`` pickles (reference: FEM_src/utils.py:63-70).  ``ordering``:
    ``"row_major"`` (default) stores this package's node order and says so in an extra ``"ordering"`` key
    (the reference's reader ignores unknown keys but would misplace the values);
    ``"dolfin"`` writes a file the REFERENCE's load_function reads correctly: dolfin's dof order, no extra
    key -- available for P1 ("design") vectors, whose dof map is known (``dolfin_p1_permutation``)."""
    vector = np.asarray(vector, dtype=np.float64)
    data = {"N": N, "domain_size": domain_size, "problem": problem}
    if ordering == "row_major":
        data["vector"] = vector
        data["ordering"] = "row_major"
    elif ordering == "dolfin":
        if problem != "design":
            raise ValueError("ordering='dolfin' is available for P1 'design' vectors only (dolfin's dof map of the "
                             "vector-P2 and Taylor-Hood spaces is not reconstructed here)")
        w, h = domain_size
        perm = dolfin_p1_permutation(int(w * N), int(h * N))
        data["vector"] = vector[perm]
    else:
        raise ValueError(f"unknown ordering {ordering!r}")
    return data


def unpack_function_data(data: dict):
    """(row-major vector, nx, ny, kind) of a pickled function dict, kind in {"P1", "P2", "TH"}.  A file
    WITHOUT the ``"ordering"`` key was written by the reference (dolfin dof order): P1 vectors are mapped
    through ``dolfin_p1_permutation``; vector-P2 / Taylor-Hood vectors in dolfin order are refused rather
    than loaded silently permuted."""
    problem = data["problem"]
    kinds = {"design": "P1", "elasticity": "P2", "fluid": "TH"}
    if problem not in kinds:
        raise ValueError(f"load_function got malformed problem: {problem}")
    w, h = data["domain_size"]
    nx, ny = int(w * data["N"]), int(h * data["N"])
    vector = np.asarray(data["vector"], dtype=np.float64)
    ordering = data.get("ordering")
    if ordering is None:
        if problem != "design":
            raise ValueError(f"{problem!r} file in dolfin dof order (written by the reference): only P1 'design' "
                             "files can be re-ordered to this package's node order")
        perm = dolfin_p1_permutation(nx, ny)
        if vector.size != perm.size:
            raise ValueError(f"design vector has {vector.size} entries, the {nx} x {ny} mesh {perm.size} vertices")
        lex = np.empty_like(vector)
        lex[perm] = vector
        vector = lex
    elif ordering != "row_major":
        raise ValueError(f"unknown ordering {ordering!r} in function file")
    return vector, nx, ny, kinds[problem]


def save_function(f: Function, filename: str, problem: str, N: int | None = None, domain_size=None,
                  *, ordering: str = "row_major"):
    """Pickle ``{N, domain_size, problem, vector}`` (reference: FEM_src/utils.py:47-70), see
    ``pack_function_data`` for the node order of ``vector``."""
    mesh = f.function_space().mesh()
    if N is None:
        N = int(round(1 / (mesh.hmin() / np.sqrt(2))))
    if domain_size is None:
        domain_size = mesh.domain_size
    with open(filename, "wb") as fh:
        pickle.dump(pack_function_data(f.vector()[:], N, domain_size, problem, ordering), fh)


def load_function(filename: str, mesh: RectangleMesh | None = None, function_space: FunctionSpace | None = None,
                  *, dtype: str = "float64", device=None):
    """Inverse of ``save_function`` with the reference's signature ``(filename, mesh=None,
    function_space=None)`` (FEM_src/utils.py:73-109): ``problem == 'design'`` gives a P1 function,
    ``'elasticity'`` a vector-P2 function, ``'fluid'`` the Taylor-Hood pair stored as ``[u | p]``.  Files
    written by the reference itself (dolfin dof order, no ``"ordering"`` key) load through
    ``unpack_function_data``."""
    with open(filename, "rb") as fh:
        data = pickle.load(fh)
    vector, nx, ny, kind = unpack_function_data(data)
    if mesh is None:
        w, h = data["domain_size"]
        mesh = RectangleMesh(w, h, nx, ny)
    if function_space is None:
        if kind == "TH":
            function_space = FunctionSpace(mesh, "TaylorHood", dtype=dtype, device=device)
        else:
            function_space = FunctionSpace(mesh, "CG", 1 if kind == "P1" else 2, dtype=dtype, device=device)
    if function_space.dim() != vector.size:
        raise ValueError(f"function file holds {vector.size} values, the function space has {function_space.dim()}")
    f = Function(function_space)
    f.vector()[:] = vector
    return f, mesh, function_space


class FEMSolver(Solver):
    def __init__(self, N: int, design_file: str, data_path: str = "output", skip_multiple: int = 1,
                 *, dtype: str = "float64", device=None, problem_options: dict | None = None,
                 verbose: bool = True, distributed: bool = False, dist_levels: int = 0):
        """``distributed=True``: one process per GPU under torch.distributed; the mesh is cut into
        strips of cell rows, every rank holds its strip + halo rows and the same
        optimisation runs collectively.  Rank 0 prints and writes the output files."""
        self.dtype_name = dtype
        self.device = torch.device(device) if device is not None else None
        self.problem_options = dict(problem_options or {})
        self.distributed = distributed
        self.dist_levels = dist_levels
        self.rank, self.world = 0, 1
        if distributed:
            import torch.distributed as dist
            if not dist.is_initialized():
                raise RuntimeError("distributed=True needs an initialised torch.distributed process group")
            self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self._engine = None
        super().__init__(N, design_file, data_path, skip_multiple)
        self.verbose = verbose and self.rank == 0
        self.problem: ElasticityProblem = self.problem

    # ------------------------------------------------------------------ hooks
    def get_name(self):
        return "FEM"

    def get_step_size(self):
        return self.parameters.fem_step_size

    def prepare_domain(self):
        if not torch.cuda.is_available():
            raise RuntimeError("FEMSolver needs a CUDA device: topomax_b200 has no CPU fallback")
        if self.device is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.mesh = RectangleMesh(self.width, self.height,
                                  int(self.width * self.N), int(self.height * self.N))
        local_rows = None
        if self.world > 1:
            # the partition comes from the library, so the engine is built before the spaces
            from .designs.design_parser import parse_design
            from .engine import Engine
            from .penalizers import ElasticPenalizer
            _, prm = parse_design(self.design_file)
            mu = prm.young_modulus / (2 * (1 + prm.poisson_ratio))
            lda = mu * prm.poisson_ratio / (0.5 - prm.poisson_ratio)
            self._engine = Engine(self.mesh.nx, self.mesh.ny, self.width, self.height, lame_lambda=lda,
                                  lame_mu=mu, simp_min=ElasticPenalizer().minimum,
                                  filter_radius=prm.filter_radius, fixed_sides=prm.fixed_sides,
                                  dtype=self.dtype_name, device=self.device, rank=self.rank,
                                  nranks=self.world, dist_levels=self.dist_levels)
            self._engine.init_comm()
            local_rows = (self._engine.cl0, self._engine.cl1)
        self.control_space = FunctionSpace(self.mesh, "CG", 1, dtype=self.dtype_name, device=self.device,
                                           local_rows=local_rows)

    def create_rho(self, volume_fraction: float):
        rho = Function(self.control_space)
        rho.tensor.fill_(volume_fraction)
        return rho

    def create_problem(self, problem_parameters):
        if isinstance(problem_parameters, ElasticityParameters):
            return ElasticityProblem(self.mesh, self.control_space, self.parameters,
                                     problem_parameters, engine=self._engine, **self.problem_options)
        if isinstance(problem_parameters, FluidParameters):
            if self.world > 1:
                raise NotImplementedError("the fluid problem runs on one GPU (SURVEY.md section 8f-3)")
            from .fluid_problem import FluidProblem
            # the elasticity options "preconditioner" / "warm_start" name other things: the fluid
            # solver's opt-in variants have their own keys and stay off unless asked for by name
            rename = {"state_rtol": "state_rtol", "state_max_iterations": "state_max_iterations",
                      "projection_rtol": "projection_rtol", "fluid_preconditioner": "preconditioner",
                      "fluid_warm_start": "warm_start", "fluid_device_scalars": "device_scalars",
                      "fluid_deterministic": "deterministic", "fluid_graph": "graph"}
            options = {rename[k]: v for k, v in self.problem_options.items() if k in rename}
            return FluidProblem(self.mesh, problem_parameters, self.parameters,
                                control_space=self.control_space, **options)
        raise ValueError(
            f"Got unknown problem '{self.parameters.problem}' "
            f"with problem parameters of type '{type(problem_parameters)}'"
        )

    # The numpy hooks always speak GLOBAL host arrays (row-major vertex grid).  Every host<->device
    # crossing goes through _h2d / _d2h (pinned staging, byte counters for bench.py's e2e leg); on a
    # sharded solver they slice / gather the rank-local strips.
    h2d_bytes = 0
    d2h_bytes = 0

    def _h2d(self, values: np.ndarray, local: bool = False) -> torch.Tensor:
        """``local``: ``values`` already is this rank's strip (owned + halo rows) of a sharded field."""
        if self.world > 1 and not local:
            from . import sharding
            t = sharding.local_p1(self.problem.engine, values)
        else:
            dtype = torch.float64 if self.dtype_name == "float64" else torch.float32
            n = int(np.size(values))
            if getattr(self, "_pinned", None) is None or self._pinned.numel() != n:
                self._pinned = torch.empty(n, dtype=dtype).pin_memory()
            self._pinned.copy_(torch.from_numpy(np.ascontiguousarray(values).reshape(-1)))
            t = self._pinned.to(self.device, non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()  # the staging buffer is reused
        self.h2d_bytes += t.numel() * t.element_size()
        return t

    def _d2h(self, tensor: torch.Tensor, local: bool = False) -> np.ndarray:
        self.d2h_bytes += tensor.numel() * tensor.element_size()
        if self.world > 1 and not local:
            from . import sharding
            return sharding.gather_p1(self.problem.engine, tensor)
        n = tensor.numel()
        if getattr(self, "_pinned_out", None) is None or self._pinned_out.numel() != n \
                or self._pinned_out.dtype != tensor.dtype:
            self._pinned_out = torch.empty(n, dtype=tensor.dtype).pin_memory()
        self._pinned_out.copy_(tensor.detach().reshape(-1), non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        # the staging buffer is reused: hand out a copy (torch's copy is threaded, numpy's is not)
        return torch.empty(n, dtype=tensor.dtype).copy_(self._pinned_out).numpy()

    def to_array(self, rho: Function) -> np.ndarray:
        return self._d2h(rho.tensor)

    def set_from_array(self, rho: Function, values: np.ndarray):
        rho.tensor.copy_(self._h2d(values))

    def integrate(self, values: np.ndarray) -> float:
        return self.problem.engine.integrate(self._h2d(values))

    def step(self, previous_psi: np.ndarray, step_size: float) -> np.ndarray:
        """``Solver.step`` (reference: src/solver.py:188-194) with host arrays in and out: the latent
        variable crosses PCIe once each way; filtered sensitivity, half step and the Newton/Brent
        volume projection run on the device (``step_device``).  ``Solver.project`` through the
        ``integrate`` hook stays available for callers that drive the projection themselves.
        Side effect: ``self.rho`` already holds expit(psi_new) on the device when this returns, so
        a caller that only needs the next objective can skip the host-side expit + upload the
        reference loop does next (doing it anyway, as ``Solver.solve`` does, is harmless)."""
        prev = self._h2d(previous_psi)
        psi = torch.empty_like(prev)
        self.step_device(prev, step_size, psi, self.rho.tensor)
        return self._d2h(psi)

    def host_array(self, tensor: torch.Tensor) -> np.ndarray:
        """Host copy of a rank-local device field (the whole field on an unsharded solver)."""
        return self._d2h(tensor, local=True)

    def step_local(self, previous_psi_local: np.ndarray, step_size: float) -> np.ndarray:
        """``step`` for a sharded solver whose HOST memory is distributed like the GPUs: every rank
        passes and receives its own strip of psi (stored rows, halo rows included); one pinned
        host-to-device and one device-to-host copy per rank, no gather.  On one GPU: ``step``."""
        prev = self._h2d(previous_psi_local, local=True)
        psi = torch.empty_like(prev)
        self.step_device(prev, step_size, psi, self.rho.tensor)
        return self._d2h(psi, local=True)

    def save_rho(self, rho: Function, file_root: str):
        rho_file = file_root + "_rho.dat"
        if self.world > 1:
            values = self.to_array(rho)  # collective
            if self.rank == 0:
                mesh = self.mesh
                data = pack_function_data(values, int(round(1 / (mesh.hmin() / np.sqrt(2)))), mesh.domain_size,
                                          "design")
                with open(rho_file, "wb") as fh:
                    pickle.dump(data, fh)
        else:
            save_function(rho, rho_file, "design")
        return os.path.basename(rho_file)

    def save_iteration(self, rho, objective, k, penalty):
        if self.world > 1 and self.rank != 0:
            self.save_rho(rho, "")  # take part in the gather, write nothing
            return
        super().save_iteration(rho, objective, k, penalty)

    def save_result(self, objectives, times, penalty, exit_condition):
        if self.rank == 0:
            super().save_result(objectives, times, penalty, exit_condition)

    # ------------------------------------------------------------------ device-resident step
    def step_device(self, previous_psi: torch.Tensor, step_size: float, psi_out: torch.Tensor,
                    rho_out: torch.Tensor):
        """One mirror-descent step on the device (reference: Solver.step/project,
        src/solver.py:149-194).  Returns sqrt(int (rho_new - expit(psi_prev))^2)."""
        engine = self.problem.engine
        gradient = self.problem.calculate_objective_gradient()
        half = engine.md_halfstep(previous_psi, gradient.tensor, step_size)
        # Newton with the iterate on the device (same iterates as scipy.optimize.newton from 0, tol 1e-12,
        # 50 iterations: src/solver.py:166-174); on failure the reference's own fallback, with the host
        # calls it makes (Newton again would fail the same way, so straight to Brent on [-r, r])
        c, _, status = engine.md_project(half, self.volume, 1e-12, 50)
        if status != 1:
            cache = {}

            def pair(c):  # one launch + one read-back serves error(c) and error'(c)
                if cache.get("c") != c:
                    cache["c"], cache["v"] = c, engine.md_volume(half, c)
                return cache["v"]

            c = find_volume_shift(lambda c: pair(c)[0] - self.volume, lambda c: pair(c)[1])
        delta_sq, _ = engine.md_apply(half, c, previous_psi, psi_out, rho_out)
        return float(np.sqrt(delta_sq))

    def solve_generic(self):
        """The base-class loop through the numpy hooks."""
        return Solver.solve(self)

    def solve(self, fixed_iterations: int | None = None):
        """Device-resident loop; same schedule, stop rules, prints and files as Solver.solve.
        ``fixed_iterations`` runs exactly that many steps without the stop rules (used to
        compare designs "after a fixed iteration count")."""
        total_timer, timer = Timer(), Timer()
        rho = self.rho.tensor
        psi = torch.log(rho / (1.0 - rho))
        previous_psi = torch.empty_like(psi)

        for penalty in self.parameters.penalties:
            self.problem.set_penalization(penalty)
            printer = Printer(self.verbose)
            if self.verbose:
                print(f"{'Penalty: ' + str(penalty):^{printer.title_length()}}")
            printer.print_title()

            timer.restart()
            objectives = [self.problem.calculate_objective(self.rho)]
            times = [timer.get_time_seconds()]
            printer.set(iteration=0, objective=objectives[0], seconds=times[0])
            deltas = []

            k = 0
            exit_condition = "Iteration did not converge"
            n_iterations = MAX_ITERATIONS if fixed_iterations is None else fixed_iterations
            for k in range(n_iterations):
                printer.print_values()
                if k % self.skip_multiple == 0 and fixed_iterations is None:
                    self.save_iteration(self.rho, objectives[-1], k, penalty)

                timer.restart()
                previous_psi.copy_(psi)
                try:
                    difference = self.step_device(previous_psi, self.step_size_at_iter(k), psi, rho)
                except ValueError as e:
                    exit_condition = str(e)
                    if self.verbose:
                        print(f"EXIT: {exit_condition}!")
                    break

                objectives.append(self.problem.calculate_objective(self.rho))
                times.append(timer.get_time_seconds())
                printer.set(iteration=k + 1, objective=objectives[-1], seconds=times[-1],
                            tolerance=self.tolerance(k), delta_rho=difference)
                deltas.append(difference)
                if fixed_iterations is not None:
                    continue

                stop = self.stop_condition(objectives, k)
                if stop is None and difference < self.tolerance(k):
                    stop = "Convergence treshold reached"
                if stop is not None:
                    exit_condition = stop
                    printer.exit(stop)
                    break
            else:
                if fixed_iterations is None:
                    printer.exit(exit_condition)
                else:
                    exit_condition = ""

            if fixed_iterations is None:
                self.save_iteration(self.rho, objectives[-1], k + 1, penalty)
                self.save_result(objectives, times, penalty, exit_condition)
            self.last_result = dict(objectives=objectives, times=times, k_final=k + 1,
                                    exit_condition=exit_condition, penalty=penalty, deltas=deltas)

        if self.verbose:
            print(f"\nTopology optimization took {total_timer.get_time_string()}")
        return self.last_result
