"""``parse_design``: design JSON -> (DomainParameters, problem parameters).

Same contract as the reference parser (reference: designs/design_parser.py:12-34):
the JSON root has exactly one key, ``"Elasticity"`` or ``"Fluid"``; unknown extra
keys next to ``domain_parameters`` / ``problem_parameters`` (e.g. ``"objective"``
in the reference's test designs) are ignored.
"""
from __future__ import annotations

import json

from .definitions import (
    DomainParameters,
    ElasticityParameters,
    FluidParameters,
    ProblemType,
)

_PARAMETER_RECORDS = {
    ProblemType.FLUID: FluidParameters,
    ProblemType.ELASTICITY: ElasticityParameters,
}


def parse_design(filename: str):
    with open(filename, "rb") as fh:
        root = json.load(fh)

    keys = list(root.keys())
    if len(keys) != 1:
        raise ValueError(f"Malformed design: expected exactly one root key, got: {keys}")
    (problem_name,) = keys
    body = root[problem_name]

    domain_parameters = DomainParameters.from_dict(problem_name, body["domain_parameters"])
    record = _PARAMETER_RECORDS.get(domain_parameters.problem)
    if record is None:
        raise ValueError(f"Unknown problem: {domain_parameters.problem}")
    return domain_parameters, record.from_dict(body["problem_parameters"])
