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


def test_oracle_point_evaluation_reproduces_polynomials():
    """evaluate_field / sample_function (FEM_src/utils.py:112-162): P1 reproduces linear, vector P2
    quadratic fields exactly at arbitrary points, including on cell edges, diagonals and corners."""
    from oracle.fem_oracle import evaluate_field, sample_function
    mesh = StructuredMesh(3.0, 1.0, 6, 2)
    X, Y = np.meshgrid(mesh.xv, mesh.yv, indexing="xy")
    p1 = (2 * X - 3 * Y + 1).ravel()
    XL, YL = np.meshgrid(mesh.xl, mesh.yl, indexing="xy")
    p2 = np.stack([(XL ** 2 - XL * YL + 2).ravel(), (YL ** 2 + 3 * XL).ravel()], 1).ravel()
    rng = np.random.default_rng(5)
    xs = np.concatenate([rng.random(50) * 3.0, [0.0, 3.0, 0.5, 1.25, 3.0]])
    ys = np.concatenate([rng.random(50), [0.0, 1.0, 0.5, 0.25, 0.0]])
    assert np.abs(evaluate_field(mesh, p1, 1, xs, ys)[:, 0] - (2 * xs - 3 * ys + 1)).max() < 1e-14
    got = evaluate_field(mesh, p2, 2, xs, ys)
    assert np.abs(got[:, 0] - (xs ** 2 - xs * ys + 2)).max() < 1e-13
    assert np.abs(got[:, 1] - (ys ** 2 + 3 * xs)).max() < 1e-13
    with pytest.raises(ValueError):
        evaluate_field(mesh, p1, 1, [3.5], [0.5])
    # sample counts of the reference: multiple of N with >= points per unit length, +1 for "edges"
    rays, grid = sample_function(mesh, p1, 1, 5, "edges", 2)
    assert grid.shape == (7, 19, 1) and len(rays[0]) == 19 and len(rays[1]) == 7
    rays, grid = sample_function(mesh, p2, 2, 7, "center", 2)
    assert grid.shape == (8, 24, 2)
    # values at the vertices are the nodal values
    _, grid = sample_function(mesh, p1, 1, 2, "edges", 2)
    assert np.abs(grid[:, :, 0].ravel() - p1).max() < 1e-14
    with pytest.raises(ValueError):
        sample_function(mesh, p1, 1, 2, "corner", 2)


def test_traction_corner_node_loads_both_sides():
    """TractionExpression.eval is a function of the NODE (FEM_src/elasisity_problem.py:47-70): a corner node
    inside the window of a traction on one side also feeds the P2 interpolant on the corner edge of the
    adjacent side.  Totals: the own side integrates the indicator's interpolant, the adjacent corner edge adds
    int phi_corner = h/6 times the value."""
    from oracle.fem_oracle import StructuredMesh
    W, H, nx, ny = 3.0, 1.0, 12, 4
    mesh = StructuredMesh(W, H, nx, ny)
    hy, hx = H / ny, W / nx
    b = mesh.load_vector(None, [("Right", 0.875, 0.25, 0.0, -2.0)]).reshape(2 * ny + 1, 2 * nx + 1, 2)
    # own side: nodes at y = 0.75 (vertex), 0.875 (mid), 1.0 (corner) carry -2: one full edge + the -2 at y=0.75
    # seen from the edge below (v1 of that edge)
    corner_edge_total = -2.0 * hx / 6   # int of the corner's P2 basis function over the top corner edge
    corner_self = -2.0 * hx * 4 / 30    # of which this much lands on the corner dof itself
    own = b[:, 2 * nx, 1].sum()         # right side, the shared corner dof included
    assert abs(own - ((-2.0) * (hy + hy / 6) + corner_self)) < 1e-14
    top = b[2 * ny, : 2 * nx, 1].sum()  # adjacent side without the shared corner dof
    assert abs(top - (corner_edge_total - corner_self)) < 1e-14
    assert not b[:, :, 0].any()
