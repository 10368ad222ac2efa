"""Console table of the optimisation loop (same columns as the reference's printer,
src/printer.py:19-115; formatting is cosmetic and not part of the parity contract)."""
from __future__ import annotations

COLUMNS = ("Iteration", "Objective", "ΔObjective", "Δρ", "Tolerance", "Time", "Total time")
WIDTH = 12


def _cell(value) -> str:
    if value is None:
        return " " * WIDTH
    if isinstance(value, str):
        return f"{value:^{WIDTH}}"
    if isinstance(value, int):
        return f"{value:^{WIDTH}d}"
    return f"{value:^{WIDTH}.5g}"


class Printer:
    def __init__(self, enabled: bool = True):
        self.enabled = enabled
        self.total_time = 0.0
        self.previous_objective = None
        self.row = {}

    def title_length(self):
        return (WIDTH + 1) * len(COLUMNS) + 1

    def print_title(self):
        if self.enabled:
            print("|" + "|".join(f"{c:^{WIDTH}}" for c in COLUMNS) + "|")

    def set(self, *, iteration, objective, seconds, tolerance=None, delta_rho=None):
        delta = None if self.previous_objective is None else objective - self.previous_objective
        self.previous_objective = objective
        self.total_time += seconds
        self.row = dict(iteration=iteration, objective=objective, delta=delta, delta_rho=delta_rho,
                        tolerance=tolerance, seconds=seconds)

    def set_delta_rho(self, delta_rho):
        self.row["delta_rho"] = delta_rho

    def print_values(self):
        if not self.enabled or not self.row:
            return
        r = self.row
        from .utils import prettify_seconds
        cells = [r["iteration"], r["objective"], r["delta"], r["delta_rho"], r["tolerance"],
                 prettify_seconds(r["seconds"]), prettify_seconds(self.total_time)]
        print("|" + "|".join(_cell(c) for c in cells) + "|")

    def exit(self, condition: str):
        if self.enabled:
            self.print_values()
            print(f"EXIT: {condition}")
