"""Material interpolation (reference: src/penalizers.py:8-46).  Only the elastic SIMP law is
on the accelerated path; it is evaluated inside the CUDA kernels, this class carries its
parameters and serves host-side callers."""
from __future__ import annotations

from abc import ABC, abstractmethod


class Penalizer(ABC):
    def __init__(self):
        self.penalization: float | None = None

    def set_penalization(self, penalization: float):
        self.penalization = penalization

    def assert_has_penalization(self) -> float:
        if self.penalization is None:
            raise ValueError("You must set penalization before calling penalizer")
        return self.penalization

    @abstractmethod
    def __call__(self, rho): ...

    @abstractmethod
    def derivative(self, rho): ...


class ElasticPenalizer(Penalizer):
    """SIMP: r(rho) = m + (1 - m) rho^p with m = 1e-6."""

    def __init__(self):
        super().__init__()
        self.minimum = 1e-6

    def __call__(self, rho):
        p = self.assert_has_penalization()
        return self.minimum + (1 - self.minimum) * rho**p

    def derivative(self, rho):
        p = self.assert_has_penalization()
        return (1 - self.minimum) * p * rho ** (p - 1)
