"""SURVEY.md 8f-3, CPU side: the element arithmetic, the per-triangle work items and the MINRES loop
of the CUDA fluid path (csrc/tm_fluid.cuh), compiled for the host and run serially
(tests/hostcheck/fluid_host.cpp), against the scipy oracle (oracle/fluid_oracle.py, "mean"
regularisation).  What the GPU tests then add is the launch plumbing and the atomics."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp

from oracle.fluid_oracle import OracleFluidSolver

HERE = os.path.dirname(os.path.abspath(__file__))
D, I = ctypes.c_double, ctypes.c_int
P = ctypes.POINTER(ctypes.c_double)


def ptr(a):
    return a.ctypes.data_as(P) if a is not None else None


@pytest.fixture(scope="module")
def hc():
    build = os.path.join(HERE, "_build")
    os.makedirs(build, exist_ok=True)
    so = os.path.join(build, "libfluid_host.so")
    src = os.path.join(HERE, "hostcheck", "fluid_host.cpp")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", src, "-o", so])
    lib = ctypes.CDLL(so)
    common = [I, I, D, D, D, D, D, D]
    lib.hc_fluid_apply.argtypes = common + [P, P, P, I, P]
    lib.hc_fluid_solve.argtypes = common + [P, P, D, I, P, P]
    lib.hc_fluid_solve.restype = I
    lib.hc_fluid_objective.argtypes = common + [P, P]
    lib.hc_fluid_objective.restype = D
    lib.hc_fluid_sens.argtypes = common + [P, P, P]
    return lib


def oracle_case(repo_root, N, design="diffuser", seed=3):
    s = OracleFluidSolver(N, os.path.join(repo_root, "designs", f"{design}.json"))
    pr, m = s.problem, s.mesh
    pr.set_penalization(0.1)
    rng = np.random.default_rng(seed)
    rho = 0.05 + 0.9 * rng.random(m.n1)
    args = (m.nx, m.ny, m.W, m.H, 0.1, pr.MIN, pr.MAX, pr.viscosity)
    g = np.zeros(m.nu)
    g[pr.bc_dofs] = pr.bc_vals
    interior = np.ones(m.nu, bool)
    interior[pr.bc_dofs] = False
    return s, pr, m, rho, args, g, interior


@pytest.mark.parametrize("N,design", [(6, "diffuser"), (5, "twin_pipe")])
def test_operator_and_lifting_match_the_oracle_matrices(hc, repo_root, N, design):
    s, pr, m, rho, args, g, interior = oracle_case(repo_root, N, design)
    nu, n1 = m.nu, m.n1
    full = (pr.A0 + pr._brinkman(rho)).tocsr()
    A, G, Dm = full[:nu, :nu], full[:nu, nu:], full[nu:, :nu]
    Pi = sp.diags(interior.astype(float))
    # the reference's (grad p, v) block is -D^T on interior rows
    assert abs(Pi @ (G + Dm.T)).max() < 1e-14
    K = sp.bmat([[Pi @ A @ Pi, -(Dm @ Pi).T], [-(Dm @ Pi), None]], format="csr")
    rng = np.random.default_rng(1)
    x = rng.standard_normal(nu + n1)
    x[:nu][~interior] = 0.0
    y = np.zeros(nu + n1)
    diag = np.zeros(nu + n1)
    hc.hc_fluid_apply(*args, ptr(rho), ptr(x), ptr(y), 0, ptr(diag))
    ref = K @ x
    assert np.abs(y - ref).max() < 1e-12 * np.abs(ref).max()
    dA = (Pi @ A @ Pi).diagonal() + (~interior)
    assert np.abs(diag[:nu] - dA).max() < 1e-12 * dA.max()
    assert diag[nu:].min() > 0
    # lifting of the boundary values
    xg = np.concatenate([g, np.zeros(n1)])
    hc.hc_fluid_apply(*args, ptr(rho), ptr(xg), ptr(y), 1, None)
    ref_u, ref_p = Pi @ (A @ g), -(Dm @ g)
    assert np.abs(y[:nu] - ref_u).max() < 1e-12 * np.abs(ref_u).max()
    assert np.abs(y[nu:] - ref_p).max() < 1e-12 * max(np.abs(ref_p).max(), 1e-300)


@pytest.mark.parametrize("N,design", [(8, "diffuser"), (6, "pipe_bend")])
def test_state_solve_objective_and_sensitivity_match_the_oracle(hc, repo_root, N, design):
    s, pr, m, rho, args, g, interior = oracle_case(repo_root, N, design, seed=9)
    up = np.zeros(m.nu + m.n1)
    relres = D(0.0)
    its = hc.hc_fluid_solve(*args, ptr(rho), ptr(g), 1e-12, 20000, ptr(up), ctypes.byref(relres))
    assert its > 0, (its, relres.value)
    obj_o = pr.calculate_objective(rho)          # also sets pr.u
    u = up[:m.nu]
    assert np.abs(u - pr.u).max() < 1e-8 * np.abs(pr.u).max()
    # pressures agree up to the constant
    dp = up[m.nu:] - pr.p
    assert np.abs(dp - dp.mean()).max() < 1e-6 * max(np.abs(pr.p - pr.p.mean()).max(), 1e-300)
    obj = hc.hc_fluid_objective(*args, ptr(rho), ptr(np.ascontiguousarray(pr.u)))
    assert abs(obj - obj_o) < 1e-12 * obj_o
    rhs = np.zeros(m.n1)
    hc.hc_fluid_sens(*args, ptr(rho), ptr(np.ascontiguousarray(pr.u)), ptr(rhs))
    grad_o = pr.calculate_objective_gradient()   # = M1^-1 rhs
    assert np.abs(rhs - pr.M1 @ grad_o).max() < 1e-12 * np.abs(rhs).max()


@pytest.mark.parametrize("design,N", [("diffuser", 20), ("pipe_bend", 10), ("twin_pipe", 12)])
def test_product_boundary_values_and_penalizer_match_the_oracle(repo_root, design, N):
    """Host logic of topomax_b200/fluid_problem.py (no GPU needed): the Dirichlet values of
    BoundaryFlows + the two DirichletBC objects, and FluidPenalizer."""
    from topomax_b200.designs.design_parser import parse_design
    from topomax_b200.fluid_problem import BoundaryFlows, FluidPenalizer
    from topomax_b200.mesh import RectangleMesh

    path = os.path.join(repo_root, "designs", f"{design}.json")
    s = OracleFluidSolver(N, path)
    dom, prm = parse_design(path)
    mesh = RectangleMesh(dom.width, dom.height, s.mesh.nx, s.mesh.ny)
    g = BoundaryFlows((dom.width, dom.height), prm.flows, mesh).nodal_values().reshape(-1)
    ref = np.zeros(s.mesh.nu)
    ref[s.problem.bc_dofs] = s.problem.bc_vals
    assert np.array_equal(g, ref)
    assert np.abs(ref).max() > 0
    pen = FluidPenalizer()
    with pytest.raises(ValueError):
        pen(0.5)
    pen.set_penalization(0.1)
    s.problem.set_penalization(0.1)
    rho = np.linspace(0, 1, 11)
    assert np.array_equal(pen(rho), s.problem.r(rho))
    assert np.array_equal(pen.derivative(rho), s.problem.r_prime(rho))


def test_marker_boundary_values_equal_the_plain_on_boundary_condition():
    """reference tests/test_marker.py:56-93 in our terms: with flows on Left/Right that vanish at
    the corners, the two marker-based DirichletBC objects (profile on the flow sides, then no-slip on
    the others) prescribe the same nodal values as one "on_boundary" condition carrying the
    BoundaryFlows expression -- so the two solves of that test coincide."""
    from topomax_b200.designs.definitions import Flow, Side
    from topomax_b200.fluid_problem import BoundaryFlows
    from topomax_b200.mesh import RectangleMesh

    domain_size = (1.5, 1.0)
    mesh = RectangleMesh(domain_size[0], domain_size[1], 30, 20)
    flows = [Flow(Side.LEFT, 0.25, 1 / 6, 1), Flow(Side.LEFT, 0.75, 1 / 6, 1),
             Flow(Side.RIGHT, 0.25, 1 / 6, -1), Flow(Side.RIGHT, 0.75, 1 / 6, -1)]
    bf = BoundaryFlows(domain_size, flows, mesh)
    marker = bf.nodal_values()
    # "on_boundary": evaluate the expression (FEM_src/fluid_problem.py:26-42) at every boundary node
    X, Y = bf.lattice_coordinates()
    W, H = domain_size
    ux = np.zeros_like(X)
    for flow in flows:
        side, center, length, rate = flow.to_tuple()
        sign, where = (1.0, X == 0.0) if side == Side.LEFT else (-1.0, X == W)
        ux += sign * np.where(where, bf.get_flow(Y, center, length, rate), 0.0)
    on_boundary = (X == 0.0) | (X == W) | (Y == 0) | (Y == H)
    simple = np.stack([np.where(on_boundary, ux, 0.0), np.zeros_like(X)], axis=-1)
    assert np.array_equal(marker, simple)
    assert np.abs(marker).max() == 1.0 or np.abs(marker).max() > 0.9
    # discrete flux balance of this design: inflow and outflow profiles are mirror images
    assert abs(marker[:, 0, 0].sum() - marker[:, -1, 0].sum()) < 1e-14
