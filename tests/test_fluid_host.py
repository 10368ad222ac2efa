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


# ---------------------------------------------------------------------------------------------
# multigrid on per-triangle local matrices (csrc/tm_trimg.cuh): Galerkin checks against scipy
# ---------------------------------------------------------------------------------------------
def _p2_prolongation(mc, mf):
    """scalar P2 prolongation by evaluating every coarse basis function at the fine nodes"""
    from oracle.fem_oracle import evaluate_field

    X, Y = np.meshgrid(mf.xl, mf.yl, indexing="xy")
    P = sp.lil_matrix((mf.n2, mc.n2))
    for j in range(mc.n2):
        vals = np.zeros(mc.nu)
        vals[2 * j] = 1.0
        col = np.asarray(evaluate_field(mc, vals, 2, X.ravel(), Y.ravel()))
        col = col[:, 0] if col.ndim == 2 else col[0::2]
        nz = np.flatnonzero(np.abs(col) > 1e-14)
        P[nz, j] = col[nz]
    return P.tocsr()


def _p1_prolongation(mc, mf):
    from oracle.fem_oracle import evaluate_field

    X, Y = np.meshgrid(mf.xv, mf.yv, indexing="xy")
    P = sp.lil_matrix((mf.n1, mc.n1))
    for j in range(mc.n1):
        vals = np.zeros(mc.n1)
        vals[j] = 1.0
        col = np.asarray(evaluate_field(mc, vals, 1, X.ravel(), Y.ravel())).ravel()
        nz = np.flatnonzero(np.abs(col) > 1e-14)
        P[nz, j] = col[nz]
    return P.tocsr()


def _boundary_mask(m):
    I, J = np.meshgrid(np.arange(m.Lx), np.arange(m.Ly), indexing="xy")
    return ((I == 0) | (I == m.Lx - 1) | (J == 0) | (J == m.Ly - 1)).ravel()


@pytest.fixture(scope="module")
def hc_mg(hc):
    common = [I, I, D, D, D, D, D, D]
    hc.hc_trimg_level_apply.argtypes = common + [P, I, I, P, P]
    hc.hc_trimg_level_apply.restype = I
    hc.hc_trimg_prolong.argtypes = [I, I, I, P, P]
    hc.hc_trimg_restrict.argtypes = [I, I, I, P, P]
    hc.hc_fluid_solve_mg.argtypes = common + [P, P, D, I, P, P]
    hc.hc_fluid_solve_mg.restype = I
    return hc


def test_multigrid_transfers_and_galerkin_operators(hc_mg, repo_root):
    from oracle.fem_oracle import StructuredMesh

    s, pr, m, rho, args, g, interior = oracle_case(repo_root, 8, "diffuser", seed=5)
    mc = StructuredMesh(m.W, m.H, m.nx // 2, m.ny // 2)
    mcc = StructuredMesh(m.W, m.H, m.nx // 4, m.ny // 4)
    rng = np.random.default_rng(2)
    # ---- velocity hierarchy: scalar operator M_r + K on both components, Dirichlet boundary
    P2 = _p2_prolongation(mc, m)
    bf, bc = _boundary_mask(m), _boundary_mask(mc)
    P2m = sp.diags((~bf).astype(float)) @ P2 @ sp.diags((~bc).astype(float))
    xc = rng.standard_normal(2 * mc.n2)
    xf = np.zeros(2 * m.n2)
    hc_mg.hc_trimg_prolong(0, m.nx, m.ny, ptr(xc), ptr(xf))
    ref = np.stack([P2m @ xc[0::2], P2m @ xc[1::2]], 1).ravel()
    assert np.abs(xf - ref).max() < 1e-13
    rf = rng.standard_normal(2 * m.n2)
    rc = np.zeros(2 * mc.n2)
    hc_mg.hc_trimg_restrict(0, m.nx, m.ny, ptr(rf), ptr(rc))
    ref = np.stack([P2m.T @ rf[0::2], P2m.T @ rf[1::2]], 1).ravel()
    assert np.abs(rc - ref).max() < 1e-13
    full = (pr.A0 + pr._brinkman(rho)).tocsr()
    As = full[: m.nu, : m.nu][0::2][:, 0::2]                       # the scalar operator on the lattice
    A0 = sp.diags((~bf).astype(float)) @ As @ sp.diags((~bf).astype(float))
    A1 = (P2m.T @ A0 @ P2m).tocsr()
    P2cc = _p2_prolongation(mcc, mc)
    bcc = _boundary_mask(mcc)
    P2ccm = sp.diags((~bc).astype(float)) @ P2cc @ sp.diags((~bcc).astype(float))
    A2 = (P2ccm.T @ A1 @ P2ccm).tocsr()
    for level, (A, mesh) in enumerate(((A0, m), (A1, mc), (A2, mcc))):
        x = rng.standard_normal(2 * mesh.n2)
        y = np.zeros_like(x)
        nl = hc_mg.hc_trimg_level_apply(*args, ptr(rho), 0, level, ptr(x), ptr(y))
        assert nl == 4                                              # 8 -> 4 -> 2 -> 1 cells per side
        ref = np.stack([A @ x[0::2], A @ x[1::2]], 1).ravel()
        assert np.abs(y - ref).max() < 1e-11 * np.abs(ref).max(), level
    # ---- pressure hierarchy: P1 Darcy Laplacian int (1/r) grad.grad (+ 1e-8 relative on the diagonal)
    P1 = _p1_prolongation(mc, m)
    xc = rng.standard_normal(mc.n1)
    xf = np.zeros(m.n1)
    hc_mg.hc_trimg_prolong(1, m.nx, m.ny, ptr(xc), ptr(xf))
    assert np.abs(xf - P1 @ xc).max() < 1e-13
    rows, cols, vals = [], [], []
    for t in ("A", "B"):
        area, gl = m.geom[t]
        conn = m.tri_v[t]
        w = (1.0 / pr.r(rho[conn])).mean(axis=1) * area
        Ke = gl @ gl.T
        loc = w[:, None, None] * Ke[None]
        loc = loc + 1e-8 * np.trace(loc, axis1=1, axis2=2)[:, None, None] / 3.0 * np.eye(3)[None]
        rows.append(np.repeat(conn, 3, axis=1).ravel())
        cols.append(np.tile(conn, (1, 3)).ravel())
        vals.append(loc.reshape(len(conn), 9).ravel())
    L0 = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(m.n1, m.n1))
    L1 = (P1.T @ L0 @ P1).tocsr()
    for level, (A, n) in enumerate(((L0, m.n1), (L1, mc.n1))):
        x = rng.standard_normal(n)
        y = np.zeros(n)
        hc_mg.hc_trimg_level_apply(*args, ptr(rho), 1, level, ptr(x), ptr(y))
        assert np.abs(y - A @ x).max() < 1e-11 * np.abs(A @ x).max(), level


def test_multigrid_preconditioned_solve_matches_oracle_with_few_iterations(hc_mg, repo_root):
    s, pr, m, rho, args, g, interior = oracle_case(repo_root, 16, "diffuser", seed=9)
    up = np.zeros(m.nu + m.n1)
    relres = D(0.0)
    rho_u = np.full(m.n1, 0.5)
    its_mg = hc_mg.hc_fluid_solve_mg(*args, ptr(rho_u), ptr(g), 1e-10, 2000, ptr(up), ctypes.byref(relres))
    its_diag = hc_mg.hc_fluid_solve(*args, ptr(rho_u), ptr(g), 1e-10, 20000, ptr(up.copy()), ctypes.byref(relres))
    assert 0 < its_mg < 100 and its_diag > 2 * its_mg, (its_mg, its_diag)   # numpy prototype: 64 vs ~200
    pr.calculate_objective(rho_u)
    assert np.abs(up[:m.nu] - pr.u).max() < 1e-7 * np.abs(pr.u).max()
    # a rough random density
    its = hc_mg.hc_fluid_solve_mg(*args, ptr(rho), ptr(g), 1e-11, 5000, ptr(up), ctypes.byref(relres))
    assert its > 0, (its, relres.value)
    pr.calculate_objective(rho)
    assert np.abs(up[:m.nu] - pr.u).max() < 1e-7 * np.abs(pr.u).max()


def test_vcycles_are_symmetric_positive_definite(hc_mg, repo_root):
    """MINRES needs a symmetric positive definite preconditioner: the V-cycle (Chebyshev pre-smoothing
    from zero, coarse correction, the adjoint post-smoothing) must satisfy <V x, y> = <x, V y> and
    <V x, x> > 0, and contract the error: rho(I - V A) < 1."""
    common = [I, I, D, D, D, D, D, D]
    hc_mg.hc_trimg_vcycle.argtypes = common + [P, I, P, P]
    s, pr, m, rho, args, g, interior = oracle_case(repo_root, 8, "diffuser", seed=4)
    rng = np.random.default_rng(8)
    bmask = np.repeat(_boundary_mask(m), 2)
    for kind, n in ((0, m.nu), (1, m.n1)):
        x, y = rng.standard_normal(n), rng.standard_normal(n)
        if kind == 0:
            x[bmask] = 0.0
            y[bmask] = 0.0
        vx, vy = np.zeros(n), np.zeros(n)
        hc_mg.hc_trimg_vcycle(*args, ptr(rho), kind, ptr(x), ptr(vx))
        hc_mg.hc_trimg_vcycle(*args, ptr(rho), kind, ptr(y), ptr(vy))
        assert abs(vx @ y - x @ vy) < 1e-10 * max(abs(vx @ y), 1e-300), kind
        assert vx @ x > 0 and vy @ y > 0
        # error contraction of the stationary iteration e <- (I - V A) e, A applied through level 0
        e = x.copy()
        norms = []
        for _ in range(8):
            ae, ve = np.zeros(n), np.zeros(n)
            hc_mg.hc_trimg_level_apply(*args, ptr(rho), kind, 0, ptr(e), ptr(ae))
            hc_mg.hc_trimg_vcycle(*args, ptr(rho), kind, ptr(ae), ptr(ve))
            e = e - ve
            norms.append(np.linalg.norm(e))
        # measured: ~0.58 per cycle for the velocity block (P2, 2 smoothing steps), better for the P1 Laplacian
        assert norms[-1] < 0.7 * norms[-2] or norms[-1] < 1e-8 * norms[0], (kind, norms)


# ---------------------------------------------------------------------------------------------
# the CUDA driver itself (FluidSolver / CudaTriMG) run serially through tests/hostcheck/cuda_host_shim.h
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def driver():
    build = os.path.join(HERE, "_build")
    os.makedirs(build, exist_ok=True)
    so = os.path.join(build, "libfluid_cuda_host.so")
    src = os.path.join(HERE, "hostcheck", "fluid_cuda_host.cpp")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I" + os.path.join(HERE, "hostcheck"), "-shared", "-fPIC",
                           src, "-o", so])
    lib = ctypes.CDLL(so)
    lib.hc_driver_solve.argtypes = [I, I, D, D, D, D, D, D, P, P, D, I, I, P, P, P]
    lib.hc_driver_solve.restype = I
    return lib


@pytest.mark.parametrize("design,N", [("diffuser", 16), ("twin_pipe", 8), ("pipe_bend", 10)])
def test_cuda_driver_on_the_host_matches_oracle_with_both_preconditioners(driver, hc_mg, repo_root, design, N):
    s, pr, m, rho, args, g, interior = oracle_case(repo_root, N, design, seed=13)
    obj_o = pr.calculate_objective(rho)
    grad_o = pr.calculate_objective_gradient()
    counts = {}
    for precond in (0, 1):
        up, rhs, out3 = np.zeros(m.nu + m.n1), np.zeros(m.n1), np.zeros(3)
        its = driver.hc_driver_solve(*args, ptr(rho), ptr(g), 1e-11, 20000, precond, ptr(up), ptr(rhs), ptr(out3))
        assert its > 0 and out3[2] == 0.0, (precond, its, list(out3))
        assert np.abs(up[:m.nu] - pr.u).max() < 1e-7 * np.abs(pr.u).max()
        assert abs(out3[1] - obj_o) < 1e-8 * obj_o
        assert np.abs(rhs - pr.M1 @ grad_o).max() < 1e-7 * np.abs(rhs).max()
        counts[precond] = its
    # the same iteration counts as the element-level host check (same algorithm, same order)
    relres = D(0.0)
    up = np.zeros(m.nu + m.n1)
    assert counts[0] == hc_mg.hc_fluid_solve(*args, ptr(rho), ptr(g), 1e-11, 20000, ptr(up), ctypes.byref(relres))
    if m.nx % 2 == 0 and m.ny % 2 == 0:
        assert abs(counts[1] - hc_mg.hc_fluid_solve_mg(*args, ptr(rho), ptr(g), 1e-11, 20000, ptr(up),
                                                       ctypes.byref(relres))) <= 2
        assert counts[1] < counts[0]


def test_uncoarsenable_mesh_fails_cleanly_and_the_solver_stays_usable(driver, repo_root):
    """50 x 50 cells halve once to 25 x 25: the coarsest velocity level (51^2 nodes) exceeds the dense
    inverse's limit.  The multigrid request must throw (twice, no half-built state in between) and the
    diagonal solver of the same object must still return the oracle's solution."""
    driver.hc_driver_uncoarsenable.argtypes = [I, I, D, D, D, D, D, D, P, P, D, I, P, P]
    driver.hc_driver_uncoarsenable.restype = I
    s, pr, m, rho, args, g, interior = oracle_case(repo_root, 50, "diffuser", seed=5)
    assert (m.nx, m.ny) == (50, 50)
    obj_o = pr.calculate_objective(rho)
    up, out3 = np.zeros(m.nu + m.n1), np.zeros(3)
    its = driver.hc_driver_uncoarsenable(*args, ptr(rho), ptr(g), 1e-10, 40000, ptr(up), ptr(out3))
    assert its > 0 and out3[2] == 2.0, (its, list(out3))
    assert abs(out3[1] - obj_o) < 1e-7 * obj_o
    assert np.abs(up[:m.nu] - pr.u).max() < 1e-6 * np.abs(pr.u).max()


@pytest.mark.parametrize("precond", [0, 1])
def test_warm_start_keeps_the_solution_and_saves_iterations(driver, repo_root, precond):
    """Opt-in warm start of the fluid solver (TM_FLUID_OPT_WARM_START): the second of two solves on
    nearby densities starts from the first solution, stops at the same tolerance relative to the
    ORIGINAL right-hand side, and returns the oracle's solution."""
    driver.hc_driver_sequence.argtypes = [I, I, D, D, D, D, D, D, P, P, P, D, I, I, I, P, P]
    driver.hc_driver_sequence.restype = I
    s, pr, m, rho1, args, g, interior = oracle_case(repo_root, 16, "diffuser", seed=21)
    rng = np.random.default_rng(22)
    rho2 = np.clip(rho1 + 0.02 * rng.standard_normal(m.n1), 0.01, 0.99)
    pr.calculate_objective(rho2)
    its = {}
    for warm in (0, 1):
        up, out3 = np.zeros(m.nu + m.n1), np.zeros(3)
        its[warm] = driver.hc_driver_sequence(*args, ptr(rho1), ptr(rho2), ptr(g), 1e-10, 20000, precond, warm,
                                              ptr(up), ptr(out3))
        assert its[warm] > 0 and out3[2] == float(warm), (warm, its[warm], list(out3))
        assert np.abs(up[:m.nu] - pr.u).max() < 2e-7 * np.abs(pr.u).max()
    # a Krylov method only gains the digits the guess already has: 243 -> 226 (diagonal) for a 2 %
    # change of a random density at rtol 1e-10; the gain grows as the design settles
    assert its[1] < its[0], its


@pytest.mark.parametrize("precond", [0, 1])
def test_minres_with_device_resident_scalars(driver, repo_root, precond):
    """TM_FLUID_OPT_DEVICE_SCALARS: the Lanczos / Givens recurrences advance in one-thread scalar steps
    on the device and the host looks at the residual every 7 iterations: same solution, and the
    iteration count is the host-scalar count rounded up to the next check."""
    s, pr, m, rho, args, g, interior = oracle_case(repo_root, 16, "diffuser", seed=31)
    pr.calculate_objective(rho)
    counts = {}
    for mode in (precond, precond | 2):
        up, out3 = np.zeros(m.nu + m.n1), np.zeros(3)
        its = driver.hc_driver_solve(*args, ptr(rho), ptr(g), 1e-10, 20000, mode, ptr(up), None, ptr(out3))
        assert its > 0 and out3[2] == 0.0, (mode, its, list(out3))
        assert np.abs(up[:m.nu] - pr.u).max() < 1e-7 * np.abs(pr.u).max()
        counts[mode] = its
    host_its, dev_its = counts[precond], counts[precond | 2]
    assert dev_its % 7 == 0 and host_its <= dev_its < host_its + 7, counts
    # with a warm start as well
    driver.hc_driver_sequence.argtypes = [I, I, D, D, D, D, D, D, P, P, P, D, I, I, I, P, P]
    driver.hc_driver_sequence.restype = I
    rho2 = np.clip(rho + 0.01, 0.01, 0.99)
    pr.calculate_objective(rho2)
    up, out3 = np.zeros(m.nu + m.n1), np.zeros(3)
    its = driver.hc_driver_sequence(*args, ptr(rho), ptr(rho2), ptr(g), 1e-10, 20000, precond | 2, 1, ptr(up), ptr(out3))
    assert its > 0 and out3[2] == 1.0
    assert np.abs(up[:m.nu] - pr.u).max() < 2e-7 * np.abs(pr.u).max()


@pytest.mark.parametrize("design,N", [("diffuser", 16), ("twin_pipe", 8)])
def test_deterministic_gather_mode_matches_the_scatter_mode(driver, repo_root, design, N):
    """TM_FLUID_OPT_DETERMINISTIC: every scatter-with-atomics work item replaced by its gather form
    (one work item per output entry, fixed summation order): same operator, diagonals, sensitivity,
    transfers and therefore the same iteration counts (+-1) and solution, with both preconditioners."""
    s, pr, m, rho, args, g, interior = oracle_case(repo_root, N, design, seed=17)
    pr.calculate_objective(rho)
    grad_o = pr.calculate_objective_gradient()
    for precond in (0, 1):
        res = {}
        for mode in (precond, precond | 4):
            up, rhs, out3 = np.zeros(m.nu + m.n1), np.zeros(m.n1), np.zeros(3)
            its = driver.hc_driver_solve(*args, ptr(rho), ptr(g), 1e-11, 20000, mode, ptr(up), ptr(rhs), ptr(out3))
            assert its > 0 and out3[2] == 0.0, (mode, its, list(out3))
            res[mode] = (its, up.copy(), rhs.copy(), out3[1])
        (i0, u0, r0, o0), (i1, u1, r1, o1) = res[precond], res[precond | 4]
        assert abs(i0 - i1) <= 1, (precond, i0, i1)
        assert np.abs(u0[:m.nu] - u1[:m.nu]).max() < 1e-8 * np.abs(u0[:m.nu]).max()
        assert np.abs(r0 - r1).max() < 1e-9 * np.abs(r0).max() and abs(o0 - o1) < 1e-9 * abs(o0)
        assert np.abs(u1[:m.nu] - pr.u).max() < 1e-7 * np.abs(pr.u).max()
        assert np.abs(r1 - pr.M1 @ grad_o).max() < 1e-7 * np.abs(r1).max()


@pytest.mark.parametrize("precond", [0, 1, 5])
def test_minres_replayed_from_a_captured_graph(driver, repo_root, precond):
    """TM_FLUID_OPT_GRAPH: six iterations (the period of the vector roles) are captured once per solve and
    replayed.  The CUDA shim models stream capture -- launches and async memory operations are recorded,
    not executed, and anything illegal inside a capture throws -- so the contents and the legality of
    the captured block (including the multigrid V-cycles, precond = 1, and the gather kernels, 5) are
    checked here; the count is the host-scalar count rounded up to the 12-iteration check."""
    s, pr, m, rho, args, g, interior = oracle_case(repo_root, 16, "diffuser", seed=41)
    pr.calculate_objective(rho)
    counts = {}
    for mode in (precond, precond | 8):
        up, out3 = np.zeros(m.nu + m.n1), np.zeros(3)
        its = driver.hc_driver_solve(*args, ptr(rho), ptr(g), 1e-10, 20000, mode, ptr(up), None, ptr(out3))
        assert its > 0 and out3[2] == 0.0, (mode, its, list(out3))
        assert np.abs(up[:m.nu] - pr.u).max() < 1e-7 * np.abs(pr.u).max()
        counts[mode] = its
    plain, graph = counts[precond], counts[precond | 8]
    assert graph % 12 == 0 and plain <= graph < plain + 12, counts
