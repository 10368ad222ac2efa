"""Problem interface of the optimiser (reference: src/problem.py:7-23)."""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any


class Problem(ABC):
    @abstractmethod
    def set_penalization(self, penalization: float): ...

    @abstractmethod
    def calculate_objective_gradient(self) -> Any: ...

    @abstractmethod
    def calculate_objective(self, rho: Any) -> float: ...

    @abstractmethod
    def forward(self, rho: Any) -> Any: ...
