"""Result records, timer and root-finding helper of the optimiser
(reference: src/utils.py:12-60,138-157).  The two records are pickled into the output tree;
their ``__module__`` is set to the reference's ``src.utils`` so that files written here are
readable by the reference's tools (``plot.py``) and vice versa (see ``src/utils.py`` shim)."""
from __future__ import annotations

import os
import pickle
import time
from dataclasses import dataclass

from scipy import optimize


@dataclass
class IterationData:
    domain_size: tuple[float, float]
    objective: float
    iteration: int
    rho_file: str
    penalty: float


@dataclass
class SolverResult:
    exit_condition: str
    min_objective: float
    objectives: list[float]
    iterations: int
    min_index: int
    times: list[float]


IterationData.__module__ = "src.utils"
SolverResult.__module__ = "src.utils"


def _register_reference_module_alias():
    """Pickle stores the two records by reference as ``src.utils.<name>``.  When the repo-root ``src/`` shim
    is not importable (package used from another working directory, or installed), register this module
    under that name so that dumping and loading still work; an importable ``src.utils`` that already
    carries the records (the shim, or the reference's own module) is left alone."""
    import importlib
    import sys
    import types

    try:
        mod = importlib.import_module("src.utils")
        if getattr(mod, "IterationData", None) is IterationData:
            return
        if hasattr(mod, "IterationData"):
            return  # the reference's own src/utils.py: its classes pickle under the same path
    except Exception:
        pass
    this = sys.modules[__name__]
    pkg = sys.modules.get("src")
    if pkg is None:
        pkg = types.ModuleType("src")
        pkg.__path__ = []  # a namespace stand-in
        sys.modules["src"] = pkg
    sys.modules["src.utils"] = this
    setattr(pkg, "utils", this)



class Timer:
    def __init__(self):
        self.restart()

    def restart(self):
        self.start_time = time.time()

    def get_time_seconds(self) -> float:
        return time.time() - self.start_time

    def get_time_string(self) -> str:
        return prettify_seconds(self.get_time_seconds())


def prettify_seconds(seconds: float) -> str:
    whole = int(seconds)
    ms = (seconds - whole) * 1000
    if whole == 0:
        return f"{ms:.3g}ms"
    h, rem = divmod(whole, 3600)
    m, s = divmod(rem, 60)
    parts = []
    if h:
        parts.append(f"{h:d}h")
    if h or m:
        parts.append(f"{m:d}m")
    parts.append(f"{s:d}s")
    parts.append(f"{ms:.0f}ms")
    return " ".join(parts)


def smart_brentq(f, initial_radius: float, max_radius: float):
    """Brent's method on [-r, r], doubling r from ``initial_radius`` until f changes sign;
    ValueError once r exceeds ``max_radius`` (reference: src/utils.py:138-157)."""
    r = initial_radius
    while r <= max_radius:
        try:
            return optimize.brentq(f, -r, r, full_output=True)
        except ValueError:
            r *= 2
    raise ValueError("f(-max_radius) and f(max_radius) must have different signs!")


def get_solver_data(solver: str, design: str, root_folder="output"):
    """Read an output tree back: ``(results, data_list)`` with results = [(N, p, SolverResult)]
    and data_list = [(N, p, k, IterationData)], parsed from the file names
    ``N=.._p=.._k=...dat`` / ``N=.._p=.._result.dat`` (reference: src/utils.py:99-121)."""
    data_folder = os.path.join(root_folder, solver, design, "data")
    results, data_list = [], []
    for data_file in os.listdir(data_folder):
        data_path = os.path.join(data_folder, data_file)
        if "result" in data_file:
            n_str, p_str = [v.split("=")[1] for v in data_file.split("_")[:2]]
            with open(data_path, "rb") as fh:
                results.append((int(n_str), p_str, pickle.load(fh)))
        elif "rho" not in data_file:
            n_str, p_str, k_str = [v.split("=")[1] for v in data_file.split("_")[:3]]
            with open(data_path, "rb") as fh:
                data_list.append((int(n_str), p_str, int(k_str[:-4]), pickle.load(fh)))
    return results, data_list


_register_reference_module_alias()
