"""SURVEY.md 8f-3 on the GPU: the fluid problem through the C ABI (``tm_fluid_*``) and the
``FluidProblem`` / ``FEMSolver`` mirrors against the scipy oracle (oracle/fluid_oracle.py, "mean"
regularisation) and the reference's diffuser fixture.  fp64.  Tolerances: operator 1e-12, velocity
1e-7 of its maximum and objective 1e-8 per solve (MINRES to 1e-12), sensitivity 1e-7; whole runs:
objective trace 1e-6, designs 1e-5; the reference fixture itself only to ~1e-3, the width of the
reference's own indeterminacy (see tests/test_oracle_fluid.py).

The element arithmetic and the MINRES loop are also checked on the CPU (tests/test_fluid_host.py).
"""
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle.fluid_oracle import OracleFluidSolver

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _t(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64).cuda()


def make(repo_root, design, N, **options):
    from topomax_b200.designs.design_parser import parse_design
    from topomax_b200.fluid_problem import FluidProblem
    from topomax_b200.mesh import Function, RectangleMesh

    path = os.path.join(repo_root, "designs", f"{design}.json")
    s = OracleFluidSolver(N, path)
    dom, prm = parse_design(path)
    mesh = RectangleMesh(dom.width, dom.height, s.mesh.nx, s.mesh.ny)
    problem = FluidProblem(mesh, prm, dom, **options)
    return s, problem, Function


@pytest.mark.parametrize("design,N", [("diffuser", 6), ("twin_pipe", 5), ("pipe_bend", 33)])
def test_operator_and_lifting(repo_root, design, N):
    s, problem, Function = make(repo_root, design, N)
    pr, m = s.problem, s.mesh
    pr.set_penalization(0.1)
    problem.set_penalization(0.1)
    rng = np.random.default_rng(N)
    rho = 0.05 + 0.9 * rng.random(m.n1)
    problem.set_density(Function(problem.control_space, _t(rho)))
    nu, n1 = m.nu, m.n1
    full = (pr.A0 + pr._brinkman(rho)).tocsr()
    A, Dm = full[:nu, :nu], full[nu:, :nu]
    interior = np.ones(nu, bool)
    interior[pr.bc_dofs] = False
    Pi = sp.diags(interior.astype(float))
    K = sp.bmat([[Pi @ A @ Pi, -(Dm @ Pi).T], [-(Dm @ Pi), None]], format="csr")
    x = rng.standard_normal(nu + n1)
    x[:nu][~interior] = 0.0
    y = problem.apply_operator(_t(x), 0).cpu().numpy()
    ref = K @ x
    assert np.abs(y - ref).max() < 1e-12 * np.abs(ref).max()
    g = np.zeros(nu)
    g[pr.bc_dofs] = pr.bc_vals
    assert np.array_equal(problem.boundary_velocity.cpu().numpy(), g)
    y = problem.apply_operator(_t(np.concatenate([g, np.zeros(n1)])), 1).cpu().numpy()
    ref_u, ref_p = Pi @ (A @ g), -(Dm @ g)
    assert np.abs(y[:nu] - ref_u).max() < 1e-12 * np.abs(ref_u).max()
    assert np.abs(y[nu:] - ref_p).max() < 1e-12 * np.abs(ref_p).max()


@pytest.mark.parametrize("design,N,q", [("diffuser", 16, 0.1), ("pipe_bend", 12, 0.1), ("twin_pipe", 10, 0.01)])
def test_state_solve_objective_and_gradient(repo_root, design, N, q):
    s, problem, Function = make(repo_root, design, N, state_rtol=1e-12)
    pr, m = s.problem, s.mesh
    pr.set_penalization(q)
    rng = np.random.default_rng(N + 1)
    rho = 0.05 + 0.9 * rng.random(m.n1)
    with pytest.raises(ValueError):   # src/penalizers.py:14-21
        problem.calculate_objective(Function(problem.control_space, _t(rho)))
    problem.set_penalization(q)
    with pytest.raises(ValueError):   # FEM_src/fluid_problem.py:104-108
        problem.calculate_objective_gradient()
    obj = problem.calculate_objective(Function(problem.control_space, _t(rho)))
    obj_o = pr.calculate_objective(rho)
    u = problem.u.tensor.cpu().numpy()
    assert np.abs(u - pr.u).max() < 1e-7 * np.abs(pr.u).max()
    assert abs(obj - obj_o) < 1e-8 * obj_o
    dp = problem.p.cpu().numpy() - pr.p
    assert np.abs(dp - dp.mean()).max() < 1e-5 * np.abs(pr.p - pr.p.mean()).max()
    grad = problem.calculate_objective_gradient().tensor.cpu().numpy()
    grad_o = pr.calculate_objective_gradient()
    assert np.abs(grad - grad_o).max() < 1e-7 * np.abs(grad_o).max()
    assert problem.solve_log[-1]["iterations"] > 0


def test_diffuser_run_matches_oracle_and_fixture(repo_root, golden_dir, tmp_path):
    """reference tests/test_fluid_solver.py:33-60: FEMSolver(20, diffuser.json).solve()."""
    from FEM_src.solver import FEMSolver

    design = os.path.join(repo_root, "designs", "diffuser.json")
    solver = FEMSolver(20, design, data_path=str(tmp_path), skip_multiple=999, verbose=False)
    result = solver.solve()
    oracle = OracleFluidSolver(20, design)
    ro = oracle.solve()
    assert result["k_final"] == ro["k_final"] == 20
    assert result["exit_condition"] == ro["exit_condition"] == "Convergence treshold reached"
    trace = max(abs(a - b) / abs(b) for a, b in zip(result["objectives"], ro["objectives"]))
    assert trace < 1e-6
    rho = solver.to_array(solver.rho)
    assert np.abs(rho - ro["rho"]).max() < 1e-5
    golden = json.load(open(os.path.join(golden_dir, "diffuser_N20_reference.json")))
    assert result["k_final"] == golden["iteration"]
    assert abs(result["objectives"][-1] - golden["objective"]) < 1e-3 * golden["objective"]
    assert np.abs(rho - np.array(golden["rho_lex"])).max() < 2e-2
    files = sorted(os.listdir(os.path.join(str(tmp_path), "FEM", "diffuser", "data")))
    assert "N=20_p=0.1_k=20.dat" in files and "N=20_p=0.1_k=20_rho.dat" in files


def test_penalty_continuation_and_hook_loop(repo_root, tmp_path):
    """twin_pipe.json has penalties [0.01, 0.1] (step size capped at 10 steps, src/solver.py:201-206);
    the device loop and the reference's hook loop must agree with the oracle."""
    from FEM_src.solver import FEMSolver

    design = os.path.join(repo_root, "designs", "twin_pipe.json")
    solver = FEMSolver(12, design, data_path=str(tmp_path / "a"), skip_multiple=999, verbose=False)
    result = solver.solve(fixed_iterations=3)
    oracle = OracleFluidSolver(12, design)
    ro = oracle.solve(fixed_iterations=3)
    trace = max(abs(a - b) / abs(b) for a, b in zip(result["objectives"], ro["objectives"]))
    assert trace < 1e-6
    assert np.abs(solver.to_array(solver.rho) - ro["rho"]).max() < 1e-5
    # the generic loop over the numpy hooks, run to its own stop
    design = os.path.join(repo_root, "designs", "pipe_bend.json")
    hooks = FEMSolver(10, design, data_path=str(tmp_path / "b"), skip_multiple=999, verbose=False)
    hooks.solve_generic()
    oracle = OracleFluidSolver(10, design)
    ro = oracle.solve()
    assert hooks.last_result["k_final"] == ro["k_final"]
    assert abs(hooks.last_result["objectives"][-1] - ro["objectives"][-1]) < 1e-6 * ro["objectives"][-1]
    assert np.abs(hooks.to_array(hooks.rho) - ro["rho"]).max() < 1e-5


# (green on the B200 since round 2: profiles/r2a_fluid_optins_pytest.txt)
@pytest.mark.parametrize("design,N", [("diffuser", 16), ("diffuser", 32), ("twin_pipe", 16)])
def test_multigrid_preconditioner_same_solution_fewer_iterations(repo_root, design, N):
    s, problem, Function = make(repo_root, design, N, state_rtol=1e-11)
    _, mg, _ = make(repo_root, design, N, state_rtol=1e-11, preconditioner="multigrid")
    pr, m = s.problem, s.mesh
    rng = np.random.default_rng(N)
    rho = 0.05 + 0.9 * rng.random(m.n1)
    for p in (problem, mg, pr):
        p.set_penalization(0.1)
    obj_d = problem.calculate_objective(Function(problem.control_space, _t(rho)))
    obj_m = mg.calculate_objective(Function(mg.control_space, _t(rho)))
    obj_o = pr.calculate_objective(rho)
    assert abs(obj_m - obj_o) < 1e-8 * obj_o and abs(obj_d - obj_o) < 1e-8 * obj_o
    assert np.abs(mg.u.tensor.cpu().numpy() - pr.u).max() < 1e-7 * np.abs(pr.u).max()
    its_d, its_m = problem.solve_log[-1]["iterations"], mg.solve_log[-1]["iterations"]
    assert its_m < 0.6 * its_d, (its_m, its_d)      # host check: 94 vs 243 at N=16, 102 vs 492 at N=32


@pytest.mark.parametrize("options", [dict(fluid_device_scalars=True), dict(fluid_warm_start=True), dict(fluid_deterministic=True),
                                     dict(fluid_deterministic=True, fluid_preconditioner="multigrid"),
                                     dict(fluid_graph=True), dict(fluid_graph=True, fluid_preconditioner="multigrid"),
                                     dict(fluid_device_scalars=True, fluid_warm_start=True,
                                          fluid_preconditioner="multigrid")])
def test_optin_solver_variants_reproduce_the_default_run(repo_root, tmp_path, options):
    from FEM_src.solver import FEMSolver

    design = os.path.join(repo_root, "designs", "diffuser.json")
    base = FEMSolver(16, design, data_path=str(tmp_path / "a"), skip_multiple=999, verbose=False)
    ref = base.solve(fixed_iterations=4)
    variant = FEMSolver(16, design, data_path=str(tmp_path / "b"), skip_multiple=999, verbose=False,
                        problem_options=options)
    got = variant.solve(fixed_iterations=4)
    trace = max(abs(a - b) / abs(b) for a, b in zip(got["objectives"], ref["objectives"]))
    assert trace < 1e-7, (options, trace)
    assert np.abs(variant.to_array(variant.rho) - base.to_array(base.rho)).max() < 1e-6


def test_automatic_preconditioner_choice(repo_root):
    """Default ``preconditioner="auto"``: multigrid + graph-replayed MINRES from 64 cells per short side when
    the mesh coarsens far enough (profiles/r2d_fluid_bench.txt: 4-8x faster at N = 128 / 256), the diagonal
    preconditioner below; same state and objective either way."""
    from topomax_b200.designs.design_parser import parse_design
    from topomax_b200.fluid_problem import FluidProblem
    from topomax_b200.mesh import Function, RectangleMesh

    assert FluidProblem.multigrid_pays_off(64, 64) and FluidProblem.multigrid_pays_off(256, 128)
    assert not FluidProblem.multigrid_pays_off(32, 32)      # small: launch bound
    assert not FluidProblem.multigrid_pays_off(100, 100)    # 25 x 25 coarsest cells: too large for the dense inverse
    assert not FluidProblem.multigrid_pays_off(65, 64)      # odd: cannot be halved
    path = os.path.join(repo_root, "designs", "diffuser.json")
    dom, prm = parse_design(path)
    mesh = RectangleMesh(dom.width, dom.height, 64, 64)
    rng = np.random.default_rng(3)
    auto = FluidProblem(mesh, prm, dom, state_rtol=1e-11)
    diag = FluidProblem(mesh, prm, dom, state_rtol=1e-11, preconditioner="diagonal")
    assert (auto.preconditioner, auto.graph) == ("multigrid", True) and diag.preconditioner == "diagonal"
    small = FluidProblem(RectangleMesh(dom.width, dom.height, 20, 20), prm, dom)
    assert (small.preconditioner, small.graph) == ("diagonal", False)
    rho_values = 0.2 + 0.6 * rng.random((64 + 1) ** 2)
    objs = []
    for pr in (auto, diag):
        pr.set_penalization(0.1)
        rho = Function(pr.control_space)
        rho.vector()[:] = rho_values
        objs.append(pr.calculate_objective(rho))
    assert abs(objs[0] - objs[1]) < 1e-8 * abs(objs[1])
    assert torch.linalg.norm(auto.u.tensor - diag.u.tensor) < 1e-7 * torch.linalg.norm(diag.u.tensor)
    assert auto.solve_log[-1]["iterations"] < 0.5 * diag.solve_log[-1]["iterations"]


def test_fluid_gradient_finite_differences(repo_root, tmp_path):
    """The reference's tests/test_fluid_gradient.py:7-38 on the CUDA path: FEMSolver(10, twin_pipe), the
    derivative of the objective in the constant direction d = volume / volume_fraction against forward
    differences at t = 1e-3 and 1e-7; the error must shrink with t at the reference's rate (its un-converted
    log-log coefficient >= 4, i.e. a true slope >= 0.87).  The directional derivative of the L2-projected
    gradient G is d * int G dx, because projection preserves integrals against the constant."""
    from topomax_b200.fem_solver import FEMSolver
    from topomax_b200.mesh import Function

    solver = FEMSolver(10, os.path.join(repo_root, "designs", "twin_pipe.json"), data_path=str(tmp_path), verbose=False,
                       problem_options={"state_rtol": 1e-13, "projection_rtol": 1e-13})
    problem = solver.problem
    problem.set_penalization(0.1)
    objective = problem.calculate_objective(solver.rho)
    d = solver.volume / solver.parameters.volume_fraction
    gradient = d * problem.engine.integrate(problem.calculate_objective_gradient().tensor)
    ts, errors = [1e-3, 1e-7], []
    for t in ts:
        moved = Function(solver.control_space, solver.rho.tensor + t * d)
        almost = (problem.calculate_objective(moved) - objective) / t
        errors.append(abs(almost - gradient))
    poly = np.polynomial.Polynomial.fit(np.log(ts), np.log(errors), 1)
    print("fluid gradient FD check: gradient", gradient, "errors", errors, "coefficient", poly.coef[1])
    assert errors[0] < 1e-2 * abs(gradient)
    assert poly.coef[1] >= 4
