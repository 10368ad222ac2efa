"""The oracle against the reference's own fixtures (CPU only).

Mirrors reference tests/test_elasticity_solver.py:30-55 and tests/test_filter.py:25-60.
"""
import json
import os

import numpy as np
import pytest

from oracle.fem_oracle import StructuredMesh, l2_error_p1, solve_spd, triangle_rule
from oracle.md_oracle import OracleSolver


def test_triangle_rule_exact_to_degree_6():
    from math import factorial

    pts, wts = triangle_rule(4)
    assert abs(wts.sum() - 1) < 1e-15
    for i in range(7):
        for j in range(7 - i):
            for k in range(7 - i - j):
                exact = 2 * factorial(i) * factorial(j) * factorial(k) / factorial(i + j + k + 2)
                num = float(np.sum(wts * pts[:, 0] ** i * pts[:, 1] ** j * pts[:, 2] ** k))
                assert abs(num - exact) < 1e-15


def test_golden_triangle_n10(repo_root, golden_dir):
    """tests/test_elasticity_solver.py:30-55: k=24, |obj - golden| < 1e-14, int (rho-rho*)^2 < 1e-14."""
    ref = json.load(open(os.path.join(golden_dir, "triangle_N10_reference.json")))
    s = OracleSolver(10, os.path.join(repo_root, "designs", "triangle.json"))
    r = s.solve()
    assert r["k_final"] == ref["iteration"] == 24
    assert r["exit_condition"] == "Convergence treshold reached"
    assert abs(r["objectives"][-1] - ref["objective"]) < 1e-14
    diff = r["rho"] - np.array(ref["rho_lex"])
    _, M1 = s.mesh.p1_matrices()
    assert float(diff @ (M1 @ diff)) < 1e-14
    assert np.abs(diff).max() < 1e-12


def test_negative_control_other_diagonal_differs(repo_root, golden_dir):
    """The fixture discriminates: the sorted golden vector is not symmetric under x-mirror."""
    ref = json.load(open(os.path.join(golden_dir, "triangle_N10_reference.json")))
    rho = np.array(ref["rho_lex"]).reshape(11, 11)
    assert np.abs(rho - rho[:, ::-1]).max() > 1e-3


def test_filter_identity_and_convergence():
    """tests/test_filter.py:25-60, including the un-converted Polynomial.fit coefficient quirk
    of tests/utils.py:11-33."""
    mesh = StructuredMesh(1.0, 1.0, 10, 10)
    K1, M1 = mesh.p1_matrices()
    np.random.seed(198)
    rho = np.random.random(mesh.n1)
    xi = solve_spd(0.0 * K1 + M1, M1 @ rho)
    d = xi - rho
    assert np.sqrt(d @ (M1 @ d)) < 1e-14

    eps = np.e / np.pi

    def err(N):
        m = StructuredMesh(1.0, 1.0, N, N)
        K, M = m.p1_matrices()
        X, Y = np.meshgrid(m.xv, m.yv, indexing="xy")
        rho = ((8 * eps * eps * np.pi**2 + 1) * np.cos(2 * np.pi * X) * np.cos(2 * np.pi * Y)).ravel()
        xi = solve_spd(eps * eps * K + M, M @ rho)
        return l2_error_p1(m, xi, lambda x, y: np.cos(2 * np.pi * x) * np.cos(2 * np.pi * y))

    Ns = list(range(10, 91, 10))
    errors = [err(N) for N in Ns]
    poly = np.polynomial.Polynomial.fit(np.log(Ns), np.log(errors), 1)
    assert poly.coef[1] <= -2  # reference-style (scaled) coefficient
    true_slope = poly.convert().coef[1]
    assert -2.1 < true_slope < -1.9


def test_nodal_weights_pattern():
    """interior h^2, edge h^2/2, corners BL & TR h^2/3, BR & TL h^2/6."""
    m = StructuredMesh(3.0, 1.0, 6, 2)
    w = m.nodal_weights().reshape(3, 7)
    h2 = 0.25
    assert np.allclose(w[1, 1:-1], h2)
    assert np.allclose(w[0, 1:-1], h2 / 2) and np.allclose(w[1, 0], h2 / 2)
    assert np.isclose(w[0, 0], h2 / 3) and np.isclose(w[-1, -1], h2 / 3)
    assert np.isclose(w[0, -1], h2 / 6) and np.isclose(w[-1, 0], h2 / 6)
    assert np.isclose(w.sum(), 3.0)


def test_oracle_anchors_reproduce(repo_root, golden_dir):
    anchors = json.load(open(os.path.join(golden_dir, "oracle_anchors.json")))
    case = next(c for c in anchors["cases"] if c["design"] == "short_cantilever")
    s = OracleSolver(case["N"], os.path.join(repo_root, "designs", "short_cantilever.json"))
    s.problem.set_penalization(3.0)
    obj = s.problem.calculate_objective(s.rho)
    assert abs(obj - case["objectives"][0]) / obj < 1e-9
    assert np.isclose(s.problem.b[1::2].sum(), case["load_sum"][1], rtol=1e-12)


def test_elasticity_matrix_symmetric_and_rigid_body_nullspace():
    m = StructuredMesh(2.0, 1.0, 4, 2)
    rng = np.random.default_rng(0)
    xi = rng.random(m.n1)
    K = m.elasticity_matrix(xi, 1.0, 1.0)
    assert abs(K - K.T).max() < 1e-13
    X, Y = m.node_coordinates()
    for mode in (np.stack([np.ones_like(X), 0 * X], 1), np.stack([0 * X, np.ones_like(X)], 1),
                 np.stack([-Y, X], 1)):
        assert np.abs(K @ mode.ravel()).max() < 1e-12
