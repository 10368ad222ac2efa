"""CPU oracle for the optimisation loop: entropic mirror descent over the FEM oracle.

TEST INFRASTRUCTURE ONLY (see ``oracle/fem_oracle.py`` header for who may import it).

Restates, in plain numpy/scipy:
* ``Solver.__init__`` / ``solve`` / ``step`` / ``project`` / ``tolerance`` /
  ``step_size_at_iter`` ............ src/solver.py:53-78,149-302
* ``FEMSolver`` hooks .............. FEM_src/solver.py:38-89
* ``ElasticityProblem`` ............ FEM_src/elasisity_problem.py:79-169
* ``smart_brentq`` ................. src/utils.py:138-157

Pinned by the reference's golden run (tests/test_elasticity_solver.py:30-55): triangle.json,
N=10 must stop at saved iteration k=24 with objective 0.0018191324070894702.
"""
from __future__ import annotations

import json
import time

import numpy as np
from scipy import optimize

from .fem_oracle import StructuredMesh, lame, solve_spd


def expit(x):
    return 1.0 / (1.0 + np.exp(-x))


def expit_diff(x):
    e = expit(x)
    return e * (1 - e)


def logit(x):
    return np.log(x / (1.0 - x))


def smart_brentq(f, initial_radius, max_radius):
    r = initial_radius
    while True:
        if r > max_radius:
            raise ValueError("f(-max_radius) and f(max_radius) must have different signs!")
        try:
            return optimize.brentq(f, -r, r, full_output=True)
        except ValueError:
            r *= 2


def read_design(path):
    """Minimal reader of the design JSON (designs/design_parser.py:12-34) -> plain dict."""
    with open(path, "rb") as fh:
        root = json.load(fh)
    (kind,) = root.keys()
    if kind != "Elasticity":
        raise ValueError("oracle covers the elasticity path only")
    dom = root[kind]["domain_parameters"]
    prm = root[kind]["problem_parameters"]
    force = None
    if prm.get("body_force") is not None:
        f = prm["body_force"]
        force = (f["region"]["center"][0], f["region"]["center"][1], f["region"]["radius"],
                 f["value"][0], f["value"][1])
    tractions = None
    if prm.get("tractions") is not None:
        tractions = [(t["side"], t["center"], t["length"], t["value"][0], t["value"][1])
                     for t in prm["tractions"]]
    return dict(
        width=dom["width"], height=dom["height"], step=dom["fem_step_size"],
        penalties=dom["penalties"], volume_fraction=dom["volume_fraction"],
        fixed_sides=prm["fixed_sides"], body_force=force, tractions=tractions,
        filter_radius=prm["filter_radius"], E=prm["young_modulus"], nu=prm["poisson_ratio"],
    )


class OracleElasticityProblem:
    """FEM_src/elasisity_problem.py:76-169 on the scipy oracle."""

    def __init__(self, mesh: StructuredMesh, design: dict):
        self.mesh = mesh
        self.design = design
        self.lda, self.mu = lame(design["E"], design["nu"])
        self.minimum = 1e-6
        self.penalization = None
        self.nq = 4
        K1, M1 = mesh.p1_matrices()
        self.M1 = M1
        eps = design["filter_radius"]
        self.Af = (eps * eps) * K1 + M1
        self.b = mesh.load_vector(design["body_force"], design["tractions"])
        self.fixed = mesh.dirichlet_mask(design["fixed_sides"])
        self.b_bc = np.where(self.fixed, 0.0, self.b)
        self.u = None
        self.filtered_rho = None
        self.timings = {"filter": 0.0, "assemble": 0.0, "solve": 0.0, "sens": 0.0}

    def set_penalization(self, p):
        self.penalization = p
        # integer exponent: polynomial integrand of degree p + 2, integrated exactly (as FFC does);
        # otherwise the 16-point rule (UFL's degree estimate for a non-integer power is a heuristic;
        # PARITY UNPINNED for p != 3)
        self.nq = max(4, int(np.ceil((p + 4) / 2))) if float(p).is_integer() else 4

    def filter_nodal(self, rho):
        t0 = time.perf_counter()
        out = solve_spd(self.Af, self.M1 @ rho)
        self.timings["filter"] += time.perf_counter() - t0
        return out

    def filter_rhs(self, rhs):
        t0 = time.perf_counter()
        out = solve_spd(self.Af, rhs)
        self.timings["filter"] += time.perf_counter() - t0
        return out

    def forward(self, xi):
        if self.penalization is None:
            raise ValueError("You must set penalization before calling penalizer")
        t0 = time.perf_counter()
        K = self.mesh.elasticity_matrix(xi, self.lda, self.mu, self.penalization, self.minimum, nq=self.nq)
        t1 = time.perf_counter()
        u = solve_spd(K, self.b_bc, free=~self.fixed, lattice=(self.mesh.Lx, self.mesh.Ly))
        t2 = time.perf_counter()
        self.timings["assemble"] += t1 - t0
        self.timings["solve"] += t2 - t1
        return u

    def calculate_objective(self, rho):
        self.filtered_rho = self.filter_nodal(rho)
        self.u = self.forward(self.filtered_rho)
        return float(self.u @ self.b)

    def calculate_objective_gradient(self):
        if self.filtered_rho is None or self.u is None:
            raise ValueError(
                "You must call calculate_objective before calling calculate_objective_gradient"
            )
        t0 = time.perf_counter()
        rhs = self.mesh.sensitivity_rhs(
            self.u, self.filtered_rho, self.lda, self.mu, self.penalization, self.minimum, nq=self.nq
        )
        self.timings["sens"] += time.perf_counter() - t0
        return self.filter_rhs(rhs)


class OracleSolver:
    """src/solver.py:42-302 + FEM_src/solver.py, without the file output."""

    def __init__(self, N: int, design_file: str):
        d = read_design(design_file)
        self.design = d
        self.width, self.height = d["width"], d["height"]
        self.N = int(N / min(self.width, self.height))
        self.full_N = int(self.N * min(self.width, self.height))
        self.volume = self.width * self.height * d["volume_fraction"]
        self.step_size = d["step"]
        self.mesh = StructuredMesh(self.width, self.height,
                                   int(self.width * self.N), int(self.height * self.N))
        self.w = self.mesh.nodal_weights()
        self.rho = np.full(self.mesh.n1, d["volume_fraction"], dtype=np.float64)
        self.problem = OracleElasticityProblem(self.mesh, d)

    def integrate(self, values):
        return float(self.w @ values)

    def project(self, half_step, volume):
        def error(c):
            return self.integrate(expit(half_step + c)) - volume

        def error_derivative(c):
            return self.integrate(expit_diff(half_step + c))

        try:
            c, result = optimize.newton(error, 0, error_derivative, tol=1e-12, full_output=True)
            if result.converged:
                return half_step + c
        except RuntimeError:
            pass
        c, result = smart_brentq(error, 2, 2000)
        if not result.converged:
            raise ValueError("Projection failed to converge")
        return half_step + c

    def step(self, previous_psi, step_size):
        g = self.problem.calculate_objective_gradient()
        return self.project(previous_psi - step_size * g, self.volume)

    def tolerance(self, k):
        return min(25 * (k + 1) * 1e-5, 1e-2)

    def step_size_at_iter(self, k):
        if len(self.design["penalties"]) > 1:
            return self.step_size * min(k + 1, 10)
        return self.step_size * (k + 1)

    def solve(self, max_iterations=1000, fixed_iterations=None, history=False):
        """Returns dict(objectives, k_final, rho, exit_condition[, rhos, gradients]).

        ``fixed_iterations`` runs exactly that many steps ignoring the stop rules
        (north_star compares designs "after a fixed iteration count").
        """
        psi = logit(self.rho)
        out = {}
        for penalty in self.design["penalties"]:
            self.problem.set_penalization(penalty)
            objectives = [self.problem.calculate_objective(self.rho)]
            rhos, deltas = [self.rho.copy()], []
            k = 0
            exit_condition = ""
            n_it = max_iterations if fixed_iterations is None else fixed_iterations
            for k in range(n_it):
                previous_psi = psi.copy()
                try:
                    psi = self.step(previous_psi, self.step_size_at_iter(k))
                except ValueError as e:
                    exit_condition = str(e)
                    break
                self.rho = expit(psi)
                objectives.append(self.problem.calculate_objective(self.rho))
                if history:
                    rhos.append(self.rho.copy())
                if fixed_iterations is not None:
                    diff = np.sqrt(self.integrate((self.rho - expit(previous_psi)) ** 2))
                    deltas.append(diff)
                    continue
                if np.isnan(objectives[-1]):
                    exit_condition = "Objective is NaN"
                    break
                min_index = int(np.argmin(objectives))
                if objectives[-1] > 2 * objectives[min_index]:
                    exit_condition = "Objective is increasing"
                    break
                if k >= min_index + 50:
                    exit_condition = "Objective is not decreasing"
                    break
                diff = np.sqrt(self.integrate((self.rho - expit(previous_psi)) ** 2))
                deltas.append(diff)
                if diff < self.tolerance(k):
                    exit_condition = "Convergence treshold reached"
                    break
            else:
                if fixed_iterations is None:
                    exit_condition = "Iteration did not converge"
            out = dict(objectives=objectives, k_final=k + 1, rho=self.rho.copy(),
                       exit_condition=exit_condition, deltas=deltas, penalty=penalty)
            if history:
                out["rhos"] = rhos
        return out
