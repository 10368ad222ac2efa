"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares, the
host mirror of the reference interface behaves like the reference's (parser, schedules, exit
rules, record formats) and the product path refuses to run without a GPU."""
import os
import pickle
import re

import numpy as np
import pytest


def test_library_exports_every_header_symbol(repo_root):
    from topomax_b200 import _lib
    from topomax_b200 import build as tm_build

    tm_build.build()
    lib = _lib.load_library()
    header = open(os.path.join(repo_root, "include", "topomax_b200.h")).read()
    declared = set(re.findall(r"\b(tm_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.tm_version()


def test_sass_is_sm100a(repo_root):
    import subprocess
    so = os.path.join(repo_root, "topomax_b200", "libtopomax_b200.so")
    out = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_parse_design(repo_root, tmp_path):
    """reference tests/test_design_parser.py:6-49"""
    import json
    from designs.definitions import ElasticityParameters, Side
    from designs.design_parser import parse_design

    dom, prm = parse_design(os.path.join(repo_root, "designs", "cantilever.json"))
    assert isinstance(prm, ElasticityParameters)
    assert dom.width == 3 and dom.height == 1 and dom.penalties == [3]
    assert prm.body_force.region.radius == 0.05 and prm.body_force.region.center == (2.9, 0.5)
    assert prm.body_force.value == (0, -1)
    assert prm.fixed_sides == [Side.LEFT] and prm.tractions is None

    raw = json.load(open(os.path.join(repo_root, "designs", "cantilever.json")))
    raw["Elasticity"]["objective"] = "MinimizeCompliance"  # ignored extra key
    ok = tmp_path / "extra.json"
    ok.write_text(json.dumps(raw))
    parse_design(str(ok))
    raw["Elasticity"]["problem_parameters"]["body_force"]["value"] = [0.0, -1.0, 0]
    bad = tmp_path / "broken.json"
    bad.write_text(json.dumps(raw))
    with pytest.raises(ValueError):
        parse_design(str(bad))
    fluid = {"Fluid": {"domain_parameters": dict(width=1.5, height=1, fem_step_size=1, dem_step_size=1,
                                                 penalties=[0.01, 0.1], volume_fraction=1 / 3),
                       "problem_parameters": {"viscosity": 1.0, "flows": [
                           {"side": "Left", "center": 0.25, "length": 1 / 6, "rate": 1.0}]}}}
    fl = tmp_path / "fluid.json"
    fl.write_text(json.dumps(fluid))
    dom, prm = parse_design(str(fl))
    assert dom.penalties == [0.01, 0.1] and prm.flows[0].side == Side.LEFT
    with pytest.raises(ValueError):
        Side.from_string("Front")


def test_no_cpu_fallback(repo_root):
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from topomax_b200.engine import Engine
    from topomax_b200.fem_solver import FEMSolver
    with pytest.raises(RuntimeError):
        Engine(4, 4, 1.0, 1.0)
    with pytest.raises(RuntimeError):
        FEMSolver(10, os.path.join(repo_root, "designs", "triangle.json"))


def test_product_package_never_imports_oracle(repo_root):
    for base, _, files in os.walk(os.path.join(repo_root, "topomax_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(base, f)).read()
                assert "oracle" not in src, f


class _ToySolver:
    """exercises the generic Solver loop with a quadratic toy problem (no GPU)"""


def test_generic_solver_schedules_and_exit_rules(tmp_path, repo_root):
    from src.problem import Problem
    from src.solver import Solver, expit, logit
    from src.utils import IterationData, SolverResult

    class Quadratic(Problem):
        def __init__(self, n):
            self.target = np.linspace(0.2, 0.8, n)
            self.rho = None
            self.penalization = None

        def set_penalization(self, p):
            self.penalization = p

        def calculate_objective(self, rho):
            self.rho = rho.copy()
            return float(np.mean((rho - self.target) ** 2))

        def calculate_objective_gradient(self):
            if self.rho is None:
                raise ValueError("objective first")
            return 2 * (self.rho - self.target) / len(self.rho)

        def forward(self, rho):
            return rho

    class Toy(Solver):
        def get_name(self): return "TOY"
        def get_step_size(self): return 1.0
        def prepare_domain(self): self.n = 21
        def create_rho(self, vf): return np.full(self.n, vf)
        def create_problem(self, _): return Quadratic(self.n)
        def to_array(self, rho): return rho.copy()
        def set_from_array(self, rho, values): rho[:] = values
        def integrate(self, values): return float(np.mean(values) * self.width * self.height)
        def save_rho(self, rho, file_root):
            np.save(file_root + "_rho.npy", rho)
            return os.path.basename(file_root + "_rho.npy")

    s = Toy(512, os.path.join(repo_root, "designs", "short_cantilever.json"), str(tmp_path), skip_multiple=7)
    assert (s.N, s.full_N) == (102, 510)  # reference src/solver.py:66-67 truncation
    assert s.tolerance(0) == 25e-5 and s.tolerance(100) == 1e-2
    assert s.step_size_at_iter(4) == 5.0
    assert s.penalty_formatter(3.0) == "3.0"
    s.verbose = False
    s.solve()
    r = s.last_result
    assert r["exit_condition"] in ("Convergence treshold reached", "Objective is not decreasing")
    assert abs(s.integrate(s.rho) - s.volume) < 1e-9  # volume constraint holds
    out = tmp_path / "TOY" / "short_cantilever" / "data"
    k = r["k_final"]
    rec = pickle.load(open(out / f"N=510_p=3.0_k={k}.dat", "rb"))
    assert isinstance(rec, IterationData) and rec.iteration == k
    assert type(rec).__module__ == "src.utils"
    res = pickle.load(open(out / "N=510_p=3.0_result.dat", "rb"))
    assert isinstance(res, SolverResult) and res.iterations == len(r["objectives"]) == k + 1
    assert Solver.stop_condition([1.0, float("nan")], 0) == "Objective is NaN"
    assert Solver.stop_condition([1.0, 2.5], 0) == "Objective is increasing"
    assert Solver.stop_condition([1.0] + [1.5] * 51, 50) == "Objective is not decreasing"
    assert np.allclose(expit(logit(np.array([0.3]))), 0.3)


def test_projection_falls_back_to_brent():
    from topomax_b200.solver import find_volume_shift
    # derivative that sends Newton away: flat tails
    f = lambda c: np.tanh(50 * (c - 1.5))
    df = lambda c: 50 / np.cosh(50 * (c - 1.5)) ** 2
    assert abs(find_volume_shift(f, df) - 1.5) < 1e-9
    with pytest.raises(ValueError):
        find_volume_shift(lambda c: 1.0, lambda c: 0.0)


def test_sample_grid_matches_reference_sample_positions():
    """sample counts / positions of sample_function (reference FEM_src/utils.py:136-158)."""
    from topomax_b200.sampling import sample_grid
    for points, kind, N, size in [(5, "edges", 2, (3.0, 1.0)), (7, "center", 2, (3.0, 1.0)),
                                  (40, "center", 102, (10.0, 5.0)), (200, "center", 40, (3.0, 1.0))]:
        nsx, nsy, x0, dx, y0, dy, mult = sample_grid(points, kind, N, size)
        assert mult == int(np.ceil(points / N))
        extra = 1 if kind == "edges" else 0
        assert (nsx, nsy) == (int(size[0] * N * mult) + extra, int(size[1] * N * mult) + extra)
        i = np.arange(nsx)
        ref = (0.5 + i) / (mult * N) if kind == "center" else i / (mult * N)
        assert np.abs(x0 + i * dx - ref).max() < 1e-13
        assert dx == dy
    with pytest.raises(ValueError):
        sample_grid(5, "corner", 2, (1.0, 1.0))


def test_output_tree_round_trip(tmp_path):
    """get_solver_data reads back what save_iteration / save_result wrote (reference:
    src/utils.py:99-121 on the tree of src/solver.py:304-349); records unpickle as src.utils.*"""
    from src.utils import IterationData, SolverResult, get_solver_data
    folder = tmp_path / "FEM" / "bridge" / "data"
    folder.mkdir(parents=True)
    it = IterationData((12.0, 2.0), 3.5, 4, "N=20_p=3.0_k=4_rho.dat", 3.0)
    res = SolverResult("Convergence treshold reached", 3.5, [9.0, 3.5], 4, 1, [0.1, 0.2])
    with open(folder / "N=20_p=3.0_k=4.dat", "wb") as fh:
        pickle.dump(it, fh)
    with open(folder / "N=20_p=3.0_k=4_rho.dat", "wb") as fh:
        pickle.dump({"N": 10, "domain_size": (12.0, 2.0), "problem": "design", "vector": np.zeros(3)}, fh)
    with open(folder / "N=20_p=3.0_result.dat", "wb") as fh:
        pickle.dump(res, fh)
    results, data_list = get_solver_data("FEM", "bridge", str(tmp_path))
    assert results == [(20, "3.0", res)]
    assert data_list == [(20, "3.0", 4, it)]
    assert type(data_list[0][3]).__module__ == "src.utils"


def test_bench_workloads_are_the_baseline_configs(repo_root):
    """bench.py --gpus N runs the configuration BASELINE.json names for N GPUs (SURVEY.md App. B sizes)."""
    import bench

    expect = {1: ("bridge", 2048, (12288, 2048), 201383938), 2: ("triangle", 4096, (4096, 4096), 134250498),
              4: ("triangle", 4096, (4096, 4096), 134250498), 8: ("cantilever", 16384, (49152, 16384), 6442713090)}
    for world, (design, n, mesh, dofs) in expect.items():
        assert bench.workload_for(world) == (design, n)
        nx, ny = bench.mesh_of(bench.design_file(design), n)
        assert (nx, ny) == mesh
        assert bench.workload_description(design, n, nx, ny)["n_displacement_dofs"] == dofs
    assert bench.mesh_of(bench.design_file("short_cantilever"), 512) == (1020, 510)  # the reference's N truncation
    # the CPU arm's sample resolution stays within its budget under the calibrated cost model
    for design in ("bridge", "triangle", "cantilever"):
        n = bench.pick_sample_n(bench.design_file(design), 16384, 25, 300.0)
        assert 8 <= n <= 512 and bench.oracle_cost_estimate(bench.design_file(design), n) * 25 <= 300.0


@pytest.mark.parametrize("design,N", [("short_cantilever", 20), ("triangle", 12), ("bridge", 8)])
def test_bench_cpu_operator_check_on_slabs(repo_root, design, N):
    """bench.py's self-check (the C + OpenMP quadrature operator applied to a displacement) on the oracle's
    direct solution: whole mesh, then cut into rank-like strips with halo rows (2 cell rows below, 1 above)
    whose owned-row sums must add up to the whole, then a band sample.  A perturbed displacement must fail."""
    import bench
    from oracle.fem_oracle import StructuredMesh, lame, solve_spd
    from oracle.md_oracle import read_design

    path = bench.design_file(design)
    d = read_design(path)
    nx, ny = bench.mesh_of(path, N)
    mesh = StructuredMesh(d["width"], d["height"], nx, ny)
    lda, mu = lame(d["E"], d["nu"])
    rng = np.random.default_rng(7)
    xi = 0.1 + 0.8 * rng.random(mesh.n1)
    fixed = mesh.dirichlet_mask(d["fixed_sides"])
    b = mesh.load_vector(d["body_force"], d["tractions"])
    K = mesh.elasticity_matrix(xi, lda, mu)
    free = ~fixed
    u = np.zeros(mesh.nu)
    u[free] = solve_spd(K[free][:, free].tocsc(), b[free])
    shape = (2 * ny + 1, 2 * nx + 1, 2)
    U, B, XI = u.reshape(shape), np.where(fixed, 0.0, b).reshape(shape), xi.reshape(ny + 1, nx + 1)
    bb = float(np.vdot(B, B))
    X = rng.standard_normal(shape)
    Y = np.where(fixed, X.ravel(), K @ np.where(fixed, 0.0, X.ravel())).reshape(shape)  # the oracle's CSR operator
    whole, _, _ = bench.slab_operator_sums(d, nx, ny, 0, ny, U, B, XI, x=X, y_gpu=Y)
    assert np.sqrt(whole[0] / bb) < 1e-11 and whole[2] == mesh.nu
    assert np.sqrt(whole[4] / whole[5]) < 1e-13  # operator parity on a random vector
    # the direct solution's residual sits within a small factor of the half-ulp perturbation floor
    assert 0.2 < np.sqrt(whole[0] / whole[3]) < 4.0
    assert abs(whole[1] - u @ b) <= 1e-10 * abs(u @ b)
    # three strips, the library's storage rule
    cuts = [0, ny // 3, 2 * ny // 3, ny]
    total = np.zeros(6)
    for r in range(3):
        c0, c1 = cuts[r], cuts[r + 1]
        cl0, cl1 = max(0, c0 - 2), min(ny, c1 + 1)
        own = (2 * (c0 - cl0), 2 * (c1 - cl0) + (1 if r == 2 else 0))
        sl = slice(2 * cl0, 2 * cl1 + 1)
        sums, _, _ = bench.slab_operator_sums(d, nx, ny, cl0, cl1 - cl0, U[sl], B[sl], XI[cl0:cl1 + 1], own)
        total += sums
    assert total[2] == mesh.nu
    assert np.sqrt(total[0] / bb) < 1e-11 and abs(total[1] - whole[1]) <= 1e-10 * abs(whole[1])
    # band sample in the middle of the mesh: incomplete first/last row excluded
    g0, rows = ny // 4, max(2, ny // 3)
    sl = slice(2 * g0, 2 * (g0 + rows) + 1)
    band, _, _ = bench.slab_operator_sums(d, nx, ny, g0, rows, U[sl], B[sl], XI[g0:g0 + rows + 1])
    assert band[2] == (2 * rows - 1) * (2 * nx + 1) * 2 and np.sqrt(band[0] / bb) < 1e-11
    # negative control: a relative perturbation of 1e-6 of one interior value is seen
    U2 = U.copy()
    U2[ny, nx, 1] *= 1.0 + 1e-6
    bad, _, _ = bench.slab_operator_sums(d, nx, ny, 0, ny, U2, B, XI)
    assert np.sqrt(bad[0] / bb) > 1e-9


def test_bench_reference_arm_line(repo_root):
    """`bench.py --impl reference` prints exactly one JSON line on stdout whose config names the workload it
    really ran, with the steps / warm-up it was asked for."""
    import json
    import subprocess
    import sys

    out = subprocess.run([sys.executable, os.path.join(repo_root, "bench.py"), "--impl", "reference", "--design",
                          "triangle", "--N", "16", "--steps", "2", "--warmup", "1", "--no_omp_baseline"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["steps"] == 2 and line["warmup"] == 1
    assert "N=16" in line["config"]["workload"] and line["config"]["same_config_as_cuda_arm"] is True
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["value"] > 0 and line["unit"] == "iter/s" and line["metric"] == "mirror_descent_iters_per_sec"


def test_committed_bench_lines_keep_the_contract(repo_root):
    """The JSON lines bench.py printed on the B200 (committed under profiles/) carry every key of the
    measurement contract; guards the schema against edits of bench.py made without a GPU at hand."""
    import json

    one = json.loads(open(os.path.join(repo_root, "profiles", "r2w_bench.json")).read().strip().splitlines()[-1])
    eight = json.loads(open(os.path.join(repo_root, "profiles", "r2n_bench_8gpu.json")).read().strip().splitlines()[-1])
    for line in (one, eight):
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                    "scaling", "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline",
                    "parity", "cpu_baseline"):
            assert key in line, key
        assert line["metric"] == "mirror_descent_iters_per_sec" and line["higher_is_better"] is True
        assert "workload" in line["config"] and "model" not in line["config"]
        r = line["roofline"]
        for key in ("bound", "achieved", "peak", "unit", "frac", "traffic", "kernel", "step", "survey_pcg_figure"):
            assert key in r, key
        assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
        assert 0.0 < r["step"]["frac"] < 1.0 and line["gpu_launches"] > 0
        assert set(line["e2e"]) == {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
        assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
        assert set(line["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
        assert line["parity"]["ok"] is True
    assert (one["n_gpus"], eight["n_gpus"]) == (1, 8)
    assert "bridge.json N=2048" in one["config"]["workload"] and "cantilever.json N=16384" in eight["config"]["workload"]
    cb = one["cpu_baseline"]
    assert set(cb) >= {"value", "unit", "cores", "kind", "sample", "same_config_pair"} and cb["kind"] == "port"
    assert one["secondary"]["parity"]["ok"] is True and "N=512" in one["secondary"]["config"]["workload"]
    # round 2, second half: multigrid cycle window + attainable-accuracy stop are reported with the PCG figures,
    # and the latency-bound line meets the bar VERDICT round 1 set (>= 55 iterations/s)
    assert one["pcg"]["multigrid_cycle_window"] == [6, 9, 2] and one["secondary"]["multigrid_cycle_window"] == [5, 6, 2]
    assert 0.0 < one["pcg"]["fp_floor_estimate_last_solve"] < 1e-8
    assert one["pcg"]["rtol_used_last_solve"] == 0.5 * one["pcg"]["fp_floor_estimate_last_solve"]
    assert one["parity"]["relative_residual"] <= one["parity"]["relative_residual_bound"]
    assert one["secondary"]["value"] >= 55.0 and one["value"] > 3.0
    assert eight["cpu_baseline"] is None and eight["roofline"]["phases_by_rank_ms"] and len(eight["roofline"]["phases_by_rank_ms"]) == 8
    # the north-star line: every lattice row of the 6.4e9-dof mesh went through the independent CPU operator
    assert eight["parity"]["coverage"] == "every lattice row" and eight["parity"]["compliance_rel_diff"] < 1e-6
    assert eight["roofline"]["step"]["frac"] >= 0.6


def test_fluid_solver_variants_stay_off_unless_named(repo_root):
    """run.py / bench.py pass the ELASTICITY options "preconditioner" and "warm_start"; they must not
    switch on the fluid solver's opt-in variants, which have their own keys."""
    src = open(os.path.join(repo_root, "topomax_b200", "fem_solver.py")).read()
    assert '"fluid_preconditioner": "preconditioner"' in src and '"fluid_warm_start": "warm_start"' in src
    assert '"preconditioner": "preconditioner"' not in src and '"warm_start": "warm_start"' not in src


def test_function_files_carry_their_ordering_and_reference_files_load(golden_dir):
    """save_function / load_function data layer (reference: FEM_src/utils.py:47-109): files written here are
    tagged row-major; a file WITHOUT the tag is the reference's (dolfin dof order) and a P1 design is mapped
    through the dolfin permutation -- checked on the reference's own golden design; ordering='dolfin' writes a
    file the reference reads; other spaces in dolfin order are refused, not silently permuted."""
    import json

    from topomax_b200.fem_solver import dolfin_p1_permutation, pack_function_data, unpack_function_data

    gold = json.load(open(os.path.join(golden_dir, "triangle_N10_reference.json")))
    dolfin_vec, lex = np.array(gold["rho_dolfin_order"]), np.array(gold["rho_lex"])
    ref_file = {"N": 10, "domain_size": (1.0, 1.0), "problem": "design", "vector": dolfin_vec}  # as the reference writes it
    vec, nx, ny, kind = unpack_function_data(ref_file)
    assert (nx, ny, kind) == (10, 10, "P1") and np.array_equal(vec, lex)
    ours = pack_function_data(lex, 10, (1.0, 1.0), "design")
    assert ours["ordering"] == "row_major" and np.array_equal(unpack_function_data(ours)[0], lex)
    theirs = pack_function_data(lex, 10, (1.0, 1.0), "design", ordering="dolfin")
    assert "ordering" not in theirs and np.array_equal(theirs["vector"], dolfin_vec)
    # a non-square mesh: the permutation is a bijection and round-trips
    perm = dolfin_p1_permutation(7, 3)
    assert sorted(perm.tolist()) == list(range(8 * 4))
    v = np.arange(32.0)
    assert np.array_equal(unpack_function_data(pack_function_data(v, 1, (7.0, 3.0), "design", "dolfin"))[0], v)
    for problem, n in (("elasticity", 2 * 21 * 21), ("fluid", 2 * 21 * 21 + 11 * 11)):
        tagged = pack_function_data(np.zeros(n), 10, (1.0, 1.0), problem)
        assert unpack_function_data(tagged)[3] in ("P2", "TH")
        untagged = {k: v for k, v in tagged.items() if k != "ordering"}
        with pytest.raises(ValueError):
            unpack_function_data(untagged)
        with pytest.raises(ValueError):
            pack_function_data(np.zeros(n), 10, (1.0, 1.0), problem, ordering="dolfin")
    with pytest.raises(ValueError):
        unpack_function_data({"N": 2, "domain_size": (1.0, 1.0), "problem": "nope", "vector": np.zeros(9)})


def test_result_records_pickle_from_another_working_directory(tmp_path):
    """IterationData / SolverResult pickle as src.utils.<name> (the reference's path); without the repo-root
    src/ shim on sys.path the package registers itself under that name."""
    import subprocess
    import sys

    code = (
        "import sys, pickle\n"
        f"sys.path = [p for p in sys.path if p not in ('', {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})]\n"
        "import importlib.util, types\n"
        f"root = {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r}\n"
        "pkg = types.ModuleType('topomax_b200'); pkg.__path__ = [root + '/topomax_b200']; sys.modules['topomax_b200'] = pkg\n"
        "spec = importlib.util.spec_from_file_location('topomax_b200.utils', root + '/topomax_b200/utils.py')\n"
        "m = importlib.util.module_from_spec(spec); sys.modules['topomax_b200.utils'] = m; spec.loader.exec_module(m)\n"
        "d = m.IterationData((1.0, 2.0), 0.5, 3, 'x_rho.dat', 3.0)\n"
        "blob = pickle.dumps(d)\n"
        "assert b'src.utils' in blob\n"
        "back = pickle.loads(blob)\n"
        "assert back == d and type(back) is m.IterationData\n"
        "print('ok')\n")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=str(tmp_path))
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stderr[-2000:]
