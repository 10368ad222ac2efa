"""Measurement of the fluid row -- a study tool, not the contract bench (bench.py keeps
the elasticity headline).  One step = one mirror-descent iteration of a fluid design (state solve by
MINRES, objective, sensitivity + projection, update); prints one JSON line with the step time, the
MINRES iteration counts and the achieved bandwidth of ``fluid_apply_kernel`` against its algorithmic
bytes per launch: read x and write y (2 n), read the 21 stored mass entries per triangle:
(2 (nu + n1) + 21 * 2 cells) * 8 bytes.

    python tools/fluid_bench.py --design designs/diffuser.json --N 128 --steps 3 --warmup 2
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from topomax_b200 import _lib  # noqa: E402
from topomax_b200.fem_solver import FEMSolver  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--design", default=os.path.join(ROOT, "designs", "diffuser.json"))
    ap.add_argument("--N", type=int, default=128)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--state_rtol", type=float, default=1e-10)
    ap.add_argument("--preconditioner", choices=("diagonal", "multigrid"), default="diagonal")
    ap.add_argument("--device_scalars", action="store_true", help="MINRES scalars resident on the device")
    ap.add_argument("--warm_start", action="store_true")
    ap.add_argument("--deterministic", action="store_true", help="gather kernels instead of scatter + atomics")
    ap.add_argument("--graph", action="store_true", help="six MINRES iterations replayed from a CUDA graph")
    args = ap.parse_args()
    lib = _lib.load_library()
    tmp = tempfile.mkdtemp(prefix="tm_fluid_bench_")
    solver = FEMSolver(args.N, args.design, data_path=tmp, verbose=False,
                       problem_options={"state_rtol": args.state_rtol, "fluid_preconditioner": args.preconditioner,
                                        "fluid_device_scalars": args.device_scalars,
                                        "fluid_warm_start": args.warm_start,
                                        "fluid_deterministic": args.deterministic, "fluid_graph": args.graph})
    problem = solver.problem
    problem.set_penalization(solver.parameters.penalties[-1])
    rho = solver.rho.tensor
    psi = torch.log(rho / (1.0 - rho))
    prev = torch.empty_like(psi)
    problem.calculate_objective(solver.rho)
    objectives = []

    def step(k):
        prev.copy_(psi)
        solver.step_device(prev, solver.step_size_at_iter(k), psi, rho)
        objectives.append(problem.calculate_objective(solver.rho))

    for k in range(args.warmup):
        step(k)
    torch.cuda.synchronize()
    n_log = len(problem.solve_log)
    launches0 = lib.tm_launch_count()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    start.record()
    for k in range(args.warmup, args.warmup + args.steps):
        step(k)
    stop.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = start.elapsed_time(stop) / args.steps
    its = [e["iterations"] for e in problem.solve_log[n_log:]]

    # the operator kernel alone, timed with events around repeated applications
    n = problem.nu + problem.n1
    x = torch.randn(n, dtype=torch.float64, device=problem.device)
    x[: problem.nu][problem.boundary_velocity != 0] = 0.0
    for _ in range(3):
        problem.apply_operator(x)
    reps = 20
    start.record()
    for _ in range(reps):
        problem.apply_operator(x)
    stop.record()
    torch.cuda.synchronize()
    apply_ms = start.elapsed_time(stop) / reps      # includes the memset of y
    cells = solver.mesh.nx * solver.mesh.ny
    alg_bytes = (2 * n + 21 * 2 * cells) * 8
    peak = None
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(peaks):
        peak = json.load(open(peaks)).get("hbm_gbs")
    achieved = alg_bytes / (apply_ms * 1e-3) / 1e9
    print(json.dumps({
        "metric": "fluid_mirror_descent_iters_per_sec", "value": 1e3 / ms, "unit": "iter/s",
        "ms_per_step": ms, "wall_s": wall, "steps": args.steps, "warmup": args.warmup, "dtype": "f64",
        "config": {"workload": f"{os.path.basename(args.design)} N={args.N} (nx={solver.mesh.nx}, ny={solver.mesh.ny}; "
                               "Taylor-Hood P2/P1, fp64)", "velocity_dofs": problem.nu, "pressure_dofs": problem.n1,
                   "state_rtol": args.state_rtol, "preconditioner": args.preconditioner,
                   "device_scalars": args.device_scalars, "warm_start": args.warm_start,
                   "deterministic": args.deterministic, "graph": args.graph},
        "minres_iterations_by_solve": its, "gpu_launches": lib.tm_launch_count() - launches0,
        "objective_trace": objectives,
        "roofline": {"bound": "hbm", "kernel": "fluid_apply_kernel (+ memset of y)", "achieved": achieved,
                     "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None,
                     "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": apply_ms, "traffic": None},
    }))


if __name__ == "__main__":
    main()
