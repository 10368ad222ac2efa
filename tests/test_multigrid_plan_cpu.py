"""CPU checks of the multigrid cycle's shape: the Python mirror of the library's planning rule against what
the library itself reported on the B200 (committed bench lines), and the numerical claim behind the W window on
the oracle's matrices (scipy; the study tool's hierarchy: nested P2 prolongation, Galerkin products)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools", "studies"))

from oracle.fem_oracle import StructuredMesh, lame  # noqa: E402
from topomax_b200 import multigrid_plan as mp  # noqa: E402


def _last_line(name):
    return json.loads(open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines()[-1])


def test_plan_mirror_matches_what_the_library_reported_on_the_gpu():
    one = _last_line("r2w_bench.json")
    two = _last_line("r2v_bench_2gpu.json")
    bridge = mp.plan(12288, 2048)
    assert mp.cycle_window(bridge) == one["pcg"]["multigrid_cycle_window"] == [6, 9, 2]
    n512 = mp.plan(1020, 510)
    assert mp.cycle_window(n512) == one["secondary"]["multigrid_cycle_window"] == [5, 6, 2]
    triangle = mp.plan(4096, 4096)
    assert mp.cycle_window(triangle) == two["pcg"]["multigrid_cycle_window"] == [7, 9, 2]
    # tools/cycle_study.py printed levels / first tail level of the same meshes (profiles/r2t_cycle_study_*.jsonl)
    for name, levels in (("r2t_cycle_study_bridge2048.jsonl", bridge), ("r2t_cycle_study_n512.jsonl", n512)):
        head = json.loads(open(os.path.join(ROOT, "profiles", name)).readline())
        assert head["levels"] == len(levels)
        assert head["tail_first_level"] == min(lv.index for lv in levels if lv.in_tail)
        assert head["cycle_window"] == mp.cycle_window(levels)
    # shape of the cycle: one step on the finest level, two on levels 1-2, three below, exact coarsest solve
    assert [lv.smoothing_steps for lv in bridge[:4]] == [1, 2, 2, 3] and bridge[-1].smoothing_steps == 0
    assert max(bridge[-1].nx, bridge[-1].ny) <= 4 and bridge[-1].lattice_nodes * 2 <= 162
    # per application of the preconditioner the window's levels are visited 2, 4, 8, 16 times, everything below 16
    assert mp.visits(bridge)[5:11] == [1, 2, 4, 8, 16, 16]
    # the 8-GPU north-star mesh: window on replicated levels only (levels of <= 262144 cells are not sharded)
    ns = mp.plan(49152, 16384)
    first, last, gamma = mp.cycle_window(ns)
    assert gamma == 2 and ns[first].nx * ns[first].ny <= 262144 and (ns[first].ny, ns[last].ny) == (32, 4)
    # tiny meshes (tests): no level qualifies -> plain V-cycle
    assert mp.cycle_window(mp.plan(10, 10)) == [-1, -1, 1]


def test_nested_p2_prolongation_and_galerkin_identity():
    """The study tool's prolongation reproduces quadratics exactly, and for a uniform density the Galerkin
    product P^T K_fine P IS the coarse mesh's stiffness matrix (nested spaces) -- the identity the library's
    six-moments-per-triangle coarsening rests on."""
    import elast_smoother_study as st

    nxc, nyc = 6, 4
    P = st.prolongation(nxc, nyc)
    coarse, fine = StructuredMesh(3.0, 2.0, nxc, nyc), StructuredMesh(3.0, 2.0, 2 * nxc, 2 * nyc)

    def quad(mesh):
        X, Y = np.meshgrid(mesh.xl, mesh.yl, indexing="xy")
        f = 0.3 + X - 2 * Y + 0.5 * X * X - X * Y + 0.25 * Y * Y
        return np.stack([f.ravel(), -2 * f.ravel()], 1).ravel()

    assert np.abs(P @ quad(coarse) - quad(fine)).max() < 1e-12
    lam, mu = lame(2.0e5, 0.3)
    Kf = fine.elasticity_matrix(np.full(fine.n1, 0.4), lam, mu)
    Kc = coarse.elasticity_matrix(np.full(coarse.n1, 0.4), lam, mu)
    G = (P.T @ Kf @ P - Kc)
    assert abs(G).max() < 1e-9 * abs(Kc).max()


def test_cycling_the_small_levels_twice_cuts_pcg_iterations_on_a_high_contrast_design():
    import elast_smoother_study as st

    mesh = StructuredMesh(8.0, 2.0, 128, 32)
    X, Y = np.meshgrid(mesh.xv, mesh.yv, indexing="xy")
    xi = np.where((np.mod(X + 0.3 * Y, 1.0) < 0.3) | (np.mod(Y, 0.5) < 0.15), 1.0, 1e-3).ravel()
    lam, mu = lame(2.0e5, 0.3)
    fixed = mesh.dirichlet_mask(["Left", "Right"])
    st.Hierarchy.fixed_sides = ["Left", "Right"]
    h = st.Hierarchy(mesh, mesh.elasticity_matrix(xi, lam, mu), fixed)
    b = np.where(fixed, 0.0, mesh.load_vector(None, [("Top", 4.0, 0.5, 0.0, -1.0)]))
    its = {}
    for name, kw in (("V", {}), ("W window", dict(window=(4, 16))), ("W window, light levels", dict(window=(4, 16), light=2))):
        x, its[name] = st.pcg(h.A[0], b, st.make_vcycle(h, **kw), rtol=1e-8)
        assert np.linalg.norm(b - h.A[0] @ x) <= 2e-8 * np.linalg.norm(b)
    print(its)
    assert its["W window"] <= 0.8 * its["V"] and its["W window, light levels"] <= 0.85 * its["V"]  # seen: 72 / 54 / 58


def test_windowed_cycle_is_a_symmetric_positive_definite_operator():
    """PCG needs a fixed SPD preconditioner.  The cycle with a W window (repeated visits start from the last
    iterate, pre- and post-smoother are the same polynomial) applied to every unit vector of a small mesh gives
    a symmetric matrix with positive spectrum, and so does the variant with two smoothing steps on levels 1-2."""
    import elast_smoother_study as st

    mesh = StructuredMesh(4.0, 1.0, 32, 8)
    X, Y = np.meshgrid(mesh.xv, mesh.yv, indexing="xy")
    xi = np.where((np.mod(X + 0.3 * Y, 1.0) < 0.3) | (np.mod(Y, 0.5) < 0.15), 1.0, 1e-3).ravel()
    lam, mu = lame(2.0e5, 0.3)
    fixed = mesh.dirichlet_mask(["Left"])
    st.Hierarchy.fixed_sides = ["Left"]
    h = st.Hierarchy(mesh, mesh.elasticity_matrix(xi, lam, mu), fixed)
    free = np.flatnonzero(~fixed)
    for kw in (dict(window=(2, 8)), dict(window=(2, 8), light=2), dict(gamma=3, gamma_levels={1, 2})):
        M = st.make_vcycle(h, **kw)
        B = np.empty((mesh.nu, free.size))
        e = np.zeros(mesh.nu)
        for c, j in enumerate(free):
            e[j] = 1.0
            B[:, c] = M(e)
            e[j] = 0.0
        B = B[free]
        assert np.abs(B - B.T).max() <= 1e-10 * np.abs(B).max(), kw
        assert np.linalg.eigvalsh(0.5 * (B + B.T)).min() > 0.0, kw


def test_cycle_and_stop_options_are_named_alike_in_header_and_python_seam():
    """The option numbers of include/topomax_b200.h, of the ctypes layer and the keyword arguments of
    ElasticityProblem that select them (defaults leave the library's automatic choices untouched)."""
    import inspect
    import re

    from topomax_b200 import _lib
    from topomax_b200.elasticity_problem import ElasticityProblem

    header = open(os.path.join(ROOT, "include", "topomax_b200.h")).read()
    for name, value in (("TM_OPT_CYCLE_FIRST", _lib.OPT_CYCLE_FIRST), ("TM_OPT_CYCLE_LAST", _lib.OPT_CYCLE_LAST),
                        ("TM_OPT_CYCLE_GAMMA", _lib.OPT_CYCLE_GAMMA), ("TM_OPT_FP_FLOOR_FACTOR", _lib.OPT_FP_FLOOR_FACTOR)):
        m = re.search(name + r"\s*=\s*(\d+)", header)
        assert m and int(m.group(1)) == value, name
    params = inspect.signature(ElasticityProblem.__init__).parameters
    assert params["multigrid_cycle"].default == "auto" and params["attainable_accuracy_stop"].default is None
    src = open(os.path.join(ROOT, "topomax_b200", "csrc", "tm_engine.cu")).read()
    for opt in ("case 133:", "case 134:", "case 135:", "case 137:"):
        assert opt in src
