"""Entropic mirror descent driver (reference: src/solver.py:42-349; algorithm by B. Keith and
T. M. Surowiec).  ``Solver`` keeps the reference's constructor, its nine abstract hooks, the
stop rules with their exit strings, and the on-disk output tree; it makes no assumption about
what the design object is, exactly like the reference's base class (numpy in, numpy out).
``topomax_b200.fem_solver.FEMSolver`` overrides the loop with a device-resident version.
"""
from __future__ import annotations

import os
import pickle
from abc import ABC, abstractmethod
from typing import Any

import numpy as np
from scipy import optimize

from .designs.design_parser import parse_design
from .printer import Printer
from .problem import Problem
from .utils import IterationData, SolverResult, Timer, smart_brentq

MAX_ITERATIONS = 1000
OBJECTIVE_INCREASING_FACTOR = 2
MAX_ITERATIONS_WITHOUT_IMPROVEMENT = 50


def expit(x):
    """Sigmoid function."""
    return 1.0 / (1.0 + np.exp(-x))


def expit_diff(x):
    """Derivative of the sigmoid function."""
    s = expit(x)
    return s * (1 - s)


def logit(x):
    """Inverse sigmoid function."""
    return np.log(x / (1.0 - x))


def find_volume_shift(error, error_derivative):
    """Root c of ``error`` as the reference finds it (src/solver.py:164-186): Newton from 0 with
    an absolute step tolerance of 1e-12 (scipy semantics, 50 iterations), then Brent on
    [-r, r], r = 2, 4, ... 2000."""
    try:
        c, result = optimize.newton(error, 0, error_derivative, tol=1e-12, full_output=True)
        if result.converged:
            return float(c)
    except RuntimeError:
        pass
    c, result = smart_brentq(error, 2, 2000)
    if not result.converged:
        raise ValueError("Projection failed to converge")
    return float(c)


class Solver(ABC):
    def __init__(self, N: int, design_file: str, data_path="output", skip_multiple=1):
        self.full_N = N
        self.design_file = design_file
        self.design_str = os.path.splitext(os.path.basename(design_file))[0]
        self.output_folder = f"{data_path}/{self.get_name()}/{self.design_str}/data"
        self.skip_multiple = skip_multiple

        self.parameters, problem_parameters = parse_design(design_file)
        self.width = self.parameters.width
        self.height = self.parameters.height

        # N = elements per unit length; full_N is re-derived from it (it names the files)
        shortest = min(self.width, self.height)
        self.N = int(self.full_N / shortest)
        self.full_N = int(self.N * shortest)

        volume_fraction = self.parameters.volume_fraction
        self.volume = self.width * self.height * volume_fraction
        self.step_size = self.get_step_size()

        self.prepare_domain()
        self.rho = self.create_rho(volume_fraction)
        self.problem = self.create_problem(problem_parameters)
        self.penalty_formatter = self.get_penalty_formatter(self.parameters.penalties)
        self.verbose = True

    # ------------------------------------------------------------------ hooks
    @abstractmethod
    def get_name(self) -> str: ...

    @abstractmethod
    def get_step_size(self) -> float: ...

    @abstractmethod
    def prepare_domain(self) -> None: ...

    @abstractmethod
    def create_rho(self, volume_fraction: float) -> Any: ...

    @abstractmethod
    def create_problem(self, problem_parameters) -> Problem: ...

    @abstractmethod
    def to_array(self, rho: Any) -> np.ndarray: ...

    @abstractmethod
    def set_from_array(self, rho: Any, values: np.ndarray) -> None: ...

    @abstractmethod
    def integrate(self, values: np.ndarray) -> float: ...

    @abstractmethod
    def save_rho(self, rho: Any, file_root: str) -> str: ...

    # ------------------------------------------------------------------ schedule
    def get_penalty_formatter(self, penalties: list[float]):
        """Zero-pad penalties so that file names sort (reference: src/solver.py:130-147)."""
        split = [str(p).split(".") for p in penalties]
        left_pad = max(len(s[0]) for s in split)
        right_pad = max(len(s[1]) for s in split)

        def penalty_formatter(penalty: float):
            left = len(str(penalty).split(".", maxsplit=1)[0])
            return "0" * (left_pad - left) + f"{penalty:.{right_pad}f}"

        return penalty_formatter

    def tolerance(self, k: int):
        return min(25 * (k + 1) * 1e-5, 1e-2)

    def step_size_at_iter(self, k: int):
        if len(self.parameters.penalties) > 1:
            return self.step_size * min(k + 1, 10)
        return self.step_size * (k + 1)

    # ------------------------------------------------------------------ one step
    def project(self, half_step: np.ndarray, volume: float):
        """half_step + c with  int expit(half_step + c) dx = volume."""
        c = find_volume_shift(
            lambda c: self.integrate(expit(half_step + c)) - volume,
            lambda c: self.integrate(expit_diff(half_step + c)),
        )
        return half_step + c

    def step(self, previous_psi: np.ndarray, step_size: float):
        gradient = self.to_array(self.problem.calculate_objective_gradient())
        return self.project(previous_psi - step_size * gradient, self.volume)

    # ------------------------------------------------------------------ loop
    def solve(self):
        total_timer, timer = Timer(), Timer()
        psi = logit(self.to_array(self.rho))

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

            k = 0
            exit_condition = "Iteration did not converge"
            for k in range(MAX_ITERATIONS):
                printer.print_values()
                if k % self.skip_multiple == 0:
                    self.save_iteration(self.rho, objectives[-1], k, penalty)

                timer.restart()
                previous_psi = psi.copy()
                try:
                    psi = self.step(previous_psi, self.step_size_at_iter(k))
                except ValueError as e:
                    exit_condition = str(e)
                    if self.verbose:
                        print(f"EXIT: {exit_condition}!")
                    break
                self.set_from_array(self.rho, expit(psi))

                objectives.append(self.problem.calculate_objective(self.rho))
                times.append(timer.get_time_seconds())
                printer.set(iteration=k + 1, objective=objectives[-1], seconds=times[-1],
                            tolerance=self.tolerance(k))

                stop = self.stop_condition(objectives, k)
                if stop is None:
                    difference = np.sqrt(
                        self.integrate((self.to_array(self.rho) - expit(previous_psi)) ** 2)
                    )
                    printer.set_delta_rho(difference)
                    if difference < self.tolerance(k):
                        stop = "Convergence treshold reached"
                if stop is not None:
                    exit_condition = stop
                    printer.exit(stop)
                    break
            else:
                printer.exit(exit_condition)

            self.save_iteration(self.rho, objectives[-1], k + 1, penalty)
            self.save_result(objectives, times, penalty, exit_condition)
            self.last_result = dict(objectives=objectives, times=times, k_final=k + 1,
                                    exit_condition=exit_condition, penalty=penalty)

        if self.verbose:
            print(f"\nTopology optimization took {total_timer.get_time_string()}")

    @staticmethod
    def stop_condition(objectives: list[float], k: int):
        """The objective-based exits (reference: src/solver.py:268-284), in their order."""
        if np.isnan(objectives[-1]):
            return "Objective is NaN"
        min_index = int(np.argmin(objectives))
        if objectives[-1] > OBJECTIVE_INCREASING_FACTOR * objectives[min_index]:
            return "Objective is increasing"
        if k >= min_index + MAX_ITERATIONS_WITHOUT_IMPROVEMENT:
            return "Objective is not decreasing"
        return None

    # ------------------------------------------------------------------ output tree
    def save_iteration(self, rho, objective: float, k: int, penalty: float):
        file_root = f"{self.output_folder}/N={self.full_N}_p={self.penalty_formatter(penalty)}_{k=}"
        os.makedirs(os.path.dirname(file_root), exist_ok=True)
        rho_basename = self.save_rho(rho, file_root)
        record = IterationData((self.width, self.height), objective, k, rho_basename, penalty)
        with open(f"{file_root}.dat", "wb") as fh:
            pickle.dump(record, fh)

    def save_result(self, objectives: list[float], times: list[float], penalty: float,
                    exit_condition: str):
        file_root = f"{self.output_folder}/N={self.full_N}_p={self.penalty_formatter(penalty)}_result"
        min_index = int(np.argmin(objectives))
        if min_index != len(objectives) - 1 and self.verbose:
            print(
                "\nWARNING: Final objective is not optimal. "
                f"Lowest objective was achived at iteration {min_index}, "
                f"with a value of {objectives[min_index]:.6g}."
            )
        record = SolverResult(exit_condition, objectives[min_index], objectives, len(objectives),
                              min_index, times)
        with open(f"{file_root}.dat", "wb") as fh:
            pickle.dump(record, fh)
