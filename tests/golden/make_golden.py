"""Regenerates the committed golden fixtures under ``tests/golden/``.

Run in the build container only (it reads ``/root/reference``, which does not exist on the
GPU box):   python tests/golden/make_golden.py

1. ``triangle_N10_reference.json`` -- the reference's own known-answer fixture for the path
   (reference: tests/test_data/FEM/triangle/data/correct_{data,rho}.dat, asserted by
   tests/test_elasticity_solver.py:30-55), converted from pickle to JSON.  The pickles are
   untrusted content: they are read through a whitelist unpickler that only admits numpy
   array reconstruction.  ``rho_dolfin_order`` is the vector as stored; ``rho_lex`` is the
   same data re-ordered to the row-major vertex grid through the dolfin P1 dof permutation
   (dofs sorted by (ix-iy, ix)).
   ``diffuser_N20_reference.json`` -- the same for the fluid fixture
   (tests/test_data/FEM/diffuser/data/correct_{data,rho}.dat, tests/test_fluid_solver.py:33-60);
   its data file pickles a ``src.utils.IterationData`` dataclass, admitted as an inert stub.
2. ``oracle_anchors.json`` -- outputs of ``oracle/`` (NOT of the reference) at a few
   configurations, used as regression anchors for oracle and CUDA path alike.  The traction
   designs are parity-unpinned by the reference; that is recorded in the file.
"""
from __future__ import annotations

import importlib
import json
import os
import pickle
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


class IterationDataStub:
    """Inert stand-in for src/utils.py:12-19 ``IterationData`` (plain attribute bag)."""


class NumpyOnlyUnpickler(pickle.Unpickler):
    ALLOWED = {
        ("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"),
        ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"),
        ("numpy", "ndarray"), ("numpy", "dtype"),
    }

    def find_class(self, module, name):
        if (module, name) == ("src.utils", "IterationData"):
            return IterationDataStub
        if (module, name) not in self.ALLOWED:
            raise pickle.UnpicklingError(f"refusing to load {module}.{name}")
        return getattr(importlib.import_module(module), name)


def load(path):
    with open(path, "rb") as fh:
        return NumpyOnlyUnpickler(fh).load()


def dolfin_p1_permutation(nx, ny):
    verts = [(ix, iy) for iy in range(ny + 1) for ix in range(nx + 1)]
    return sorted(range(len(verts)), key=lambda v: (verts[v][0] - verts[v][1], verts[v][0]))


def reference_fixture():
    base = os.path.join(REF, "tests", "test_data", "FEM", "triangle", "data")
    data = load(os.path.join(base, "correct_data.dat"))
    rho = load(os.path.join(base, "correct_rho.dat"))
    n = int(rho["N"])
    vec = np.asarray(rho["vector"], dtype=np.float64)
    perm = dolfin_p1_permutation(n, n)
    lex = np.empty_like(vec)
    lex[perm] = vec
    out = {
        "source": "reference tests/test_data/FEM/triangle/data/correct_{data,rho}.dat",
        "design": "designs/triangle.json", "N": n,
        "objective": float(data["objective"]), "iteration": int(data["iteration"]),
        "penalty": float(data["penalty"]),
        "domain_size": [float(x) for x in data["domain_size"]],
        "rho_dolfin_order": [float(x) for x in vec],
        "rho_lex": [float(x) for x in lex],
    }
    with open(os.path.join(HERE, "triangle_N10_reference.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    return out


def reference_fluid_fixture():
    base = os.path.join(REF, "tests", "test_data", "FEM", "diffuser", "data")
    data = load(os.path.join(base, "correct_data.dat")).__dict__
    rho = load(os.path.join(base, "correct_rho.dat"))
    n = int(rho["N"])
    vec = np.asarray(rho["vector"], dtype=np.float64)
    perm = dolfin_p1_permutation(n, n)
    lex = np.empty_like(vec)
    lex[perm] = vec
    out = {
        "source": "reference tests/test_data/FEM/diffuser/data/correct_{data,rho}.dat",
        "design": "designs/diffuser.json", "N": n,
        "objective": float(data["objective"]), "iteration": int(data["iteration"]),
        "penalty": float(data["penalty"]),
        "domain_size": [float(x) for x in data["domain_size"]],
        "rho_dolfin_order": [float(x) for x in vec],
        "rho_lex": [float(x) for x in lex],
    }
    with open(os.path.join(HERE, "diffuser_N20_reference.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    return out


def oracle_anchors(ref):
    from oracle.md_oracle import OracleSolver

    anchors = {"source": "oracle/ (scipy restatement), NOT the reference", "cases": []}
    # the pinned case: must agree with the reference fixture
    s = OracleSolver(10, os.path.join(ROOT, "designs", "triangle.json"))
    r = s.solve(history=True)
    assert r["k_final"] == ref["iteration"], r["k_final"]
    assert abs(r["objectives"][-1] - ref["objective"]) < 1e-14
    assert np.abs(r["rho"] - np.array(ref["rho_lex"])).max() < 1e-12
    anchors["cases"].append({
        "design": "triangle", "N": 10, "pinned_by_reference": True,
        "objectives": r["objectives"], "k_final": r["k_final"],
        "exit_condition": r["exit_condition"], "deltas": r["deltas"],
    })
    for design, N, steps in (("cantilever", 40, 3), ("short_cantilever", 50, 3), ("bridge", 20, 3)):
        s = OracleSolver(N, os.path.join(ROOT, "designs", f"{design}.json"))
        r = s.solve(fixed_iterations=steps)
        anchors["cases"].append({
            "design": design, "N": N, "pinned_by_reference": False,
            "objectives": r["objectives"], "fixed_iterations": steps,
            "deltas": r["deltas"],
            "load_sum": [float(s.problem.b[0::2].sum()), float(s.problem.b[1::2].sum())],
            "rho_checksum": float(np.dot(s.w, r["rho"])),
            "rho_l2": float(np.sqrt(np.dot(s.w, r["rho"] ** 2))),
        })
    with open(os.path.join(HERE, "oracle_anchors.json"), "w") as fh:
        json.dump(anchors, fh, indent=1)


if __name__ == "__main__":
    ref = reference_fixture()
    reference_fluid_fixture()
    oracle_anchors(ref)
    print("golden fixtures written to", HERE)
