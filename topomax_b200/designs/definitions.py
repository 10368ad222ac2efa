"""Design-file schema: the typed view of a ``designs/*.json`` file.

Mirrors the public names of the reference's ``designs/definitions.py``
(reference: designs/definitions.py:22-223) so that code written against the
reference (``Side.LEFT``, ``ElasticityParameters.fixed_sides``,
``Traction.to_tuple()`` ...) keeps working.  Only the elasticity branch is on
the accelerated path; the fluid records are parsed so that the same parser
accepts every design file the reference ships, but no fluid solver exists here.
"""
from __future__ import annotations

from dataclasses import dataclass
from enum import Enum
from typing import Sequence


def to_2_tuple(ray: Sequence[float]) -> tuple[float, float]:
    """Pair from a 2-element sequence; anything else is a malformed design
    (reference: designs/definitions.py:8-19, exercised by
    tests/test_design_parser.py:48-49 through ``broken_design.json``)."""
    if len(ray) != 2:
        raise ValueError(
            f"Got array that should have had 2 elements, but had {len(ray)} instead: '{ray}'"
        )
    return (ray[0], ray[1])


class _NamedEnum(Enum):
    """Enum whose members are looked up from the JSON spelling."""

    @classmethod
    def _spellings(cls) -> dict[str, "_NamedEnum"]:
        raise NotImplementedError

    @classmethod
    def _what(cls) -> str:
        return cls.__name__.lower()

    @classmethod
    def from_string(cls, string: str):
        table = cls._spellings()
        try:
            return table[string]
        except KeyError:
            legal = ", ".join(f"'{k}'" for k in table)
            raise ValueError(
                f"Malformed {cls._what()}: '{string}'\nLegal values are: {legal}"
            ) from None


class Side(_NamedEnum):
    LEFT = 1
    RIGHT = 2
    TOP = 3
    BOTTOM = 4

    @classmethod
    def _spellings(cls):
        return {"Left": cls.LEFT, "Right": cls.RIGHT, "Top": cls.TOP, "Bottom": cls.BOTTOM}

    @classmethod
    def get_all(cls):
        return [cls.LEFT, cls.RIGHT, cls.TOP, cls.BOTTOM]


class ElasticityObjective(_NamedEnum):
    MINIMIZE_COMPLIANCE = 1

    @classmethod
    def _what(cls):
        return "objective"

    @classmethod
    def _spellings(cls):
        return {"MinimizeCompliance": cls.MINIMIZE_COMPLIANCE}


class FluidObjective(_NamedEnum):
    MINIMIZE_POWER = 1

    @classmethod
    def _what(cls):
        return "objective"

    @classmethod
    def _spellings(cls):
        return {"MinimizePower": cls.MINIMIZE_POWER}


class ProblemType(_NamedEnum):
    FLUID = 1
    ELASTICITY = 2

    @classmethod
    def _what(cls):
        return "problem"

    @classmethod
    def _spellings(cls):
        return {"Fluid": cls.FLUID, "Elasticity": cls.ELASTICITY}


@dataclass
class CircularRegion:
    center: tuple[float, float]
    radius: float

    @classmethod
    def from_dict(cls, d: dict):
        return cls(center=to_2_tuple(d["center"]), radius=d["radius"])


@dataclass
class Force:
    """Body force ``value`` applied inside ``region`` (a disc)."""

    region: CircularRegion
    value: tuple[float, float]

    @classmethod
    def from_dict(cls, d: dict):
        return cls(region=CircularRegion.from_dict(d["region"]), value=to_2_tuple(d["value"]))


@dataclass
class Traction:
    """Boundary traction ``value`` on a window of ``side`` centred at ``center``."""

    side: Side
    center: float
    length: float
    value: tuple[float, float]

    @classmethod
    def from_dict(cls, d: dict):
        return cls(
            side=Side.from_string(d["side"]),
            center=d["center"],
            length=d["length"],
            value=to_2_tuple(d["value"]),
        )

    def to_tuple(self):
        return (self.side, self.center, self.length, self.value)


@dataclass
class ElasticityParameters:
    fixed_sides: list[Side]
    body_force: Force | None
    tractions: list[Traction] | None
    filter_radius: float
    young_modulus: float
    poisson_ratio: float

    @classmethod
    def from_dict(cls, d: dict):
        force = d.get("body_force")
        tractions = d.get("tractions")
        return cls(
            fixed_sides=[Side.from_string(s) for s in d["fixed_sides"]],
            body_force=None if force is None else Force.from_dict(force),
            tractions=None if tractions is None else [Traction.from_dict(t) for t in tractions],
            filter_radius=d["filter_radius"],
            young_modulus=d["young_modulus"],
            poisson_ratio=d["poisson_ratio"],
        )


@dataclass
class Flow:
    side: Side
    center: float
    length: float
    rate: float

    @classmethod
    def from_dict(cls, d: dict):
        return cls(Side.from_string(d["side"]), d["center"], d["length"], d["rate"])

    def to_tuple(self):
        return (self.side, self.center, self.length, self.rate)


@dataclass
class FluidParameters:
    flows: list[Flow]
    viscosity: float

    @classmethod
    def from_dict(cls, d: dict):
        return cls([Flow.from_dict(f) for f in d["flows"]], d["viscosity"])


@dataclass
class DomainParameters:
    width: float
    height: float
    problem: ProblemType
    fem_step_size: float
    dem_step_size: float
    penalties: list[float]
    volume_fraction: float

    @classmethod
    def from_dict(cls, problem: str, d: dict):
        return cls(
            width=d["width"],
            height=d["height"],
            problem=ProblemType.from_string(problem),
            fem_step_size=d["fem_step_size"],
            dem_step_size=d["dem_step_size"],
            penalties=d["penalties"],
            volume_fraction=d["volume_fraction"],
        )
