"""Writes the design files into ``designs/``: the four elasticity designs of the hot path and the
three fluid designs of the reference (the latter are consumed by ``oracle/fluid_oracle.py`` only).

The reference generates its JSONs with a small Rust program
(reference: designs/design_creator/src/main.rs:11-28); there is no cargo in this
image and the files are plain data, so this script is the equivalent emitter for
the four elasticity designs BASELINE.json names.  Field values follow
designs/{cantilever,short_cantilever,bridge,triangle}.json of the reference.

    python designs/make_designs.py
"""
from __future__ import annotations

import json
import math
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def elasticity(width, height, fem_step, dem_step, vf, fixed, force, tractions, radius, E, nu):
    return {
        "Elasticity": {
            "domain_parameters": {
                "width": width,
                "height": height,
                "fem_step_size": fem_step,
                "dem_step_size": dem_step,
                "penalties": [3.0],
                "volume_fraction": vf,
            },
            "problem_parameters": {
                "fixed_sides": fixed,
                "body_force": force,
                "tractions": tractions,
                "filter_radius": radius,
                "young_modulus": E,
                "poisson_ratio": nu,
            },
        }
    }


def fluid(width, height, fem_step, dem_step, penalties, vf, flows, viscosity=1.0):
    """Fluid branch of the schema (reference: designs/definitions.py, Flow / FluidParameters)."""
    return {
        "Fluid": {
            "domain_parameters": {
                "width": width,
                "height": height,
                "fem_step_size": fem_step,
                "dem_step_size": dem_step,
                "penalties": penalties,
                "volume_fraction": vf,
            },
            "problem_parameters": {
                "flows": [{"side": s, "center": c, "length": l, "rate": r} for s, c, l, r in flows],
                "viscosity": viscosity,
            },
        }
    }


def disc_force(cx, cy, r, fx, fy):
    return {"region": {"center": [cx, cy], "radius": r}, "value": [fx, fy]}


def traction(side, center, length, tx, ty):
    return {"side": side, "center": center, "length": length, "value": [tx, ty]}


# Helmholtz radius equivalent to a cone filter of radius 0.25: r / (2 sqrt 3)
STEEL_RADIUS = 0.25 / (2.0 * math.sqrt(3.0))

DESIGNS = {
    "cantilever": elasticity(3.0, 1.0, 25.0, 37500.0, 0.5, ["Left"],
                             disc_force(2.9, 0.5, 0.05, 0.0, -1.0), None, 0.02, 2.5, 0.25),
    "triangle": elasticity(1.0, 1.0, 25.0, 100000.0, 0.5, ["Bottom"],
                           disc_force(0.5, 0.5, 0.05, 0.0, -10.0), None, 0.02, 2.5, 0.25),
    "short_cantilever": elasticity(10.0, 5.0, 0.03, 1.5, 0.4, ["Left"], None,
                                   [traction("Right", 2.5, 1.0 / 9.0, 0.0, -2000.0)],
                                   STEEL_RADIUS, 200000.0, 0.3),
    "bridge": elasticity(12.0, 2.0, 0.001, 0.2, 0.4, ["Left", "Right"], None,
                         [traction("Top", 6.0, 0.5, 0.0, -2000.0)],
                         STEEL_RADIUS, 200000.0, 0.3),
    # values of the reference's designs/{diffuser,pipe_bend,twin_pipe}.json
    "diffuser": fluid(1.0, 1.0, 0.0011, 0.0002, [0.1], 0.5,
                      [("Left", 0.5, 1.0, 1.0), ("Right", 0.5, 1.0 / 3.0, -3.0)]),
    "pipe_bend": fluid(1.0, 1.0, 0.0015, 0.0001, [0.1], 0.251,
                       [("Left", 0.8, 0.2, 1.0), ("Bottom", 0.8, 0.2, -1.0)]),
    "twin_pipe": fluid(1.5, 1.0, 0.0015, 0.0004, [0.01, 0.1], 1.0 / 3.0,
                       [("Left", 0.25, 1.0 / 6.0, 1.0), ("Left", 0.75, 1.0 / 6.0, 1.0),
                        ("Right", 0.25, 1.0 / 6.0, -1.0), ("Right", 0.75, 1.0 / 6.0, -1.0)]),
}


def main():
    for name, design in DESIGNS.items():
        with open(os.path.join(HERE, f"{name}.json"), "w") as fh:
            json.dump(design, fh, indent=2)
            fh.write("\n")
        print("wrote", name)


if __name__ == "__main__":
    main()
