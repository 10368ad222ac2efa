"""Benchmark of the hot path: mirror-descent iterations per second (BASELINE.json metric).

A "step" is one full mirror-descent iteration of the reference's optimiser loop
(src/solver.py:247-294): filtered sensitivity (1 Helmholtz solve) -> latent step + volume
projection -> Helmholtz filter -> SIMP state solve -> compliance.  Steps W .. W+K-1 of the real
optimisation are timed, so the density field evolves exactly as in a run.

    python bench.py [--gpus N] [--steps K] [--warmup W]              # CUDA arm
    python bench.py --impl reference [--steps K] [--warmup W]        # CPU arm (oracle port)

Workload at N=1: designs/short_cantilever.json at N=512 (configs[1] of BASELINE.json; the
reference's N truncation makes it nx=1020, ny=510, 4 167 722 displacement dofs), fp64.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DESIGN = "short_cantilever"
FULL_N = 512
METRIC = "mirror_descent_iters_per_sec"
UNIT = "iter/s"


def workload_description(design, full_n, nx, ny):
    n_u = 2 * (2 * nx + 1) * (2 * ny + 1)
    return {
        "workload": f"designs/{design}.json N={full_n} (nx={nx}, ny={ny}; vector-P2 / P1, fp64)",
        "n_cells": nx * ny, "n_density_dofs": (nx + 1) * (ny + 1), "n_displacement_dofs": n_u,
    }


def weak_scaling_n(design_path, full_n, world):
    """Resolution of the weak-scaling run on `world` GPUs: N * sqrt(world) (~world x the cells), with
    the elements per unit length rounded down to a multiple of `world` so that every strip gets the
    same number of cell rows (unchanged for 2 and 4 GPUs on the default design; 1448 -> 1440 on 8)."""
    from topomax_b200.designs.design_parser import parse_design
    dom, _ = parse_design(design_path)
    shortest = min(dom.width, dom.height)
    per_unit = int(int(round(full_n * world ** 0.5)) / shortest)
    per_unit -= per_unit % world
    return int(per_unit * shortest)


def mesh_of(design_path, full_n):
    from topomax_b200.designs.design_parser import parse_design
    dom, _ = parse_design(design_path)
    n = int(full_n / min(dom.width, dom.height))
    return int(dom.width * n), int(dom.height * n)


class ClockSampler:
    """nvidia-smi sampling of SM clocks and throttle reasons during the timed region."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.device_index = device_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-i", str(self.device_index), "-lms", "100"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons, smax = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    smax = float(f[2])
                except ValueError:
                    continue
                for name, flag in zip(names, f[5:9]):
                    if flag.lower().startswith("active"):
                        reasons.add(name)
        finally:
            try:
                os.remove(self.path)
            except OSError:
                pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=smax, reasons=sorted(reasons),
                       samples=len(sm))
        return out


class Function_like:
    """minimal stand-in so that FEMSolver.to_array can read an arbitrary device tensor"""

    def __init__(self, tensor, solver):
        self.tensor = tensor


def measured_traffic(design, full_n, dtype, kernel):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as fh:
            entry = json.load(fh).get(f"{design}:{full_n}:{dtype}:{kernel}")
        return (entry["bytes_per_launch"], entry["source"]) if entry else (None, None)
    except Exception:
        return None, None


def measured_peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference algorithm (assembly + sparse direct solves)
# --------------------------------------------------------------------------------------
def oracle_iteration_seconds(design_path, sample_n, steps, warmup):
    """Times `steps` mirror-descent iterations of the oracle at resolution sample_n."""
    from oracle.md_oracle import OracleSolver, expit, logit

    s = OracleSolver(sample_n, design_path)
    s.problem.set_penalization(s.design["penalties"][0])
    psi = logit(s.rho)
    s.problem.calculate_objective(s.rho)
    times = []
    for k in range(warmup + steps):
        t0 = time.perf_counter()
        psi = s.step(psi.copy(), s.step_size_at_iter(k))
        s.rho = expit(psi)
        s.problem.calculate_objective(s.rho)
        dt = time.perf_counter() - t0
        if k >= warmup:
            times.append(dt)
    return times, s


def omp_kernel_baseline(design_path, full_n, max_iterations=60):
    """Like-for-like CPU *kernel* figure: the C + OpenMP matrix-free operator with
    Jacobi-PCG (oracle/c/elast_omp.c) on all host cores, at the FULL resolution of the GPU
    workload, for a bounded number of PCG iterations on the initial (uniform) design."""
    import numpy as np
    from oracle.fem_oracle import StructuredMesh, lame
    from oracle.md_oracle import read_design
    from oracle.omp_kernels import OmpElasticity

    d = read_design(design_path)
    n = int(full_n / min(d["width"], d["height"]))
    nx, ny = int(d["width"] * n), int(d["height"] * n)
    lda, mu = lame(d["E"], d["nu"])
    op = OmpElasticity(d["width"], d["height"], nx, ny, lda, mu, d["fixed_sides"], p=d["penalties"][0])
    xi = np.full(op.n1, d["volume_fraction"])
    if d["body_force"] is None:
        b = StructuredMesh(d["width"], d["height"], nx, ny).load_vector(None, d["tractions"])
    else:  # the P2 mass matrix of the body-force load is not worth assembling for a timing
        b = np.zeros(op.nu)
        b[1::2] = -1.0
    _, its, rel, sec = op.jacobi_pcg(xi, b, rtol=1e-10, maxit=max_iterations)
    return {"what": "C + OpenMP matrix-free P2 elasticity operator (quadrature) + Jacobi-PCG, fp64",
            "dof_iters_per_sec": op.nu * its / sec if sec > 0 else None, "cores": op.threads,
            "sample": f"{its} PCG iterations at the full workload resolution (nx={nx}, ny={ny}, {op.nu} dofs)",
            "seconds": sec}


def pick_sample_n(full_n, steps_total, budget_s):
    # sparse LU with nested dissection on a 2-D mesh: ~17 s at N=256 here, flops ~ N^3
    for n in (full_n, 384, 256, 192, 128, 96, 64):
        if n > full_n:
            continue
        est = 22.0 * (n / 256.0) ** 3 + 3.0 * (n / 256.0) ** 2
        if est * steps_total <= budget_s:
            return n
    return 64


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    design_path = os.path.join(ROOT, "designs", f"{args.design}.json")
    nx, ny = mesh_of(design_path, args.N)
    warmup = min(args.warmup, 1)
    sample_n = args.sample_n or pick_sample_n(args.N, warmup + args.steps, 240.0)
    times, s = oracle_iteration_seconds(design_path, sample_n, args.steps, warmup)
    total = sum(times)
    value = len(times) / total
    sample = (f"{len(times)} mirror-descent iteration(s) of the scipy oracle (CSR assembly + SuperLU with "
              f"nested-dissection ordering in MUMPS' role) on designs/{args.design}.json at N={sample_n} "
              f"(nx={s.mesh.nx}, ny={s.mesh.ny}, {s.mesh.nu} displacement dofs"
              + ("" if sample_n == args.N else f"; bounded sample: the metric workload is N={args.N}") + ")")
    cfg = workload_description(args.design, args.N, nx, ny)
    cfg["reference_sample_N"] = sample_n
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(times), "warmup": warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
                         "host_cores_available": os.cpu_count(),
                         "split_seconds": {k: round(v, 3) for k, v in s.problem.timings.items()}},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    try:
        line["cpu_baseline"]["state_solve_kernel"] = omp_kernel_baseline(design_path, args.N)
    except Exception as exc:  # the direct-solver figure above is the arm's value either way
        line["cpu_baseline"]["state_solve_kernel"] = {"unavailable": repr(exc)}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------
# CUDA arm
# --------------------------------------------------------------------------------------
def run_cuda_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # stdout carries exactly one JSON line: keep NCCL's version banner off it
    os.environ["NCCL_DEBUG"] = os.environ.get("TM_NCCL_DEBUG", "WARN")
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (CUDA arm) needs a GPU; use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from topomax_b200 import _lib
    from topomax_b200.fem_solver import FEMSolver
    from topomax_b200.solver import expit, logit

    design_path = args.design if os.path.isfile(args.design) else os.path.join(ROOT, "designs", f"{args.design}.json")
    tmp = tempfile.mkdtemp(prefix="tm_bench_")
    # weak scaling: the same design at N * sqrt(world), i.e. ~world x the cells, cut into one
    # strip of cell rows per GPU (NCCL halo exchange per operator application, all-reduced dots)
    run_n = args.N if (world == 1 or args.exact_N) else weak_scaling_n(design_path, args.N, world)
    base_nx, base_ny = mesh_of(design_path, args.N)
    solver = FEMSolver(run_n, design_path, data_path=tmp, verbose=False, dtype=args.dtype,
                       distributed=world > 1, dist_levels=args.dist_levels,
                       problem_options={"preconditioner": args.preconditioner, "state_rtol": args.state_rtol,
                                        "mixed_precision": args.mixed, "warm_start": not args.no_warm_start})
    problem, engine = solver.problem, solver.problem.engine
    for kv in args.engine_option:
        key, val = kv.split("=")
        engine.set_option(int(key), float(val))
    nx, ny = solver.mesh.nx, solver.mesh.ny
    n1, nu = engine.n1, engine.nu  # rank-local sizes (owned + halo rows)
    nu_global = 2 * (2 * nx + 1) * (2 * ny + 1)
    nu_base = 2 * (2 * base_nx + 1) * (2 * base_ny + 1)
    size_factor = nu_global / nu_base  # 1 on one GPU
    esize = 8 if args.dtype == "float64" else 4
    problem.set_penalization(solver.parameters.penalties[0])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident loop: W warm-up + K timed mirror-descent iterations
    rho = solver.rho.tensor
    psi = torch.log(rho / (1.0 - rho))
    prev = torch.empty_like(psi)
    objectives = [problem.calculate_objective(solver.rho)]
    k = 0
    for _ in range(args.warmup):
        prev.copy_(psi)
        solver.step_device(prev, solver.step_size_at_iter(k), psi, rho)
        objectives.append(problem.calculate_objective(solver.rho))
        k += 1

    engine.set_option(_lib.OPT_PROFILE, 1)
    engine.profile_read()
    counts0 = engine.last_solve_stats()["fine_launches_total"]
    log0 = len(problem.solve_log)
    launches0 = engine.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    in_profiler = bool(os.environ.get("TM_PROFILER_RANGE"))  # ncu --profile-from-start off
    if in_profiler:
        torch.cuda.profiler.start()
    psi_at_start, rho_at_start, k_at_start = psi.clone(), rho.clone(), k  # the e2e leg repeats these steps
    start.record()
    for _ in range(args.steps):
        prev.copy_(psi)
        solver.step_device(prev, solver.step_size_at_iter(k), psi, rho)
        objectives.append(problem.calculate_objective(solver.rho))
        k += 1
    stop.record()
    barrier()
    if in_profiler:
        torch.cuda.profiler.stop()
    elapsed_ms = start.elapsed_time(stop)
    clocks = sampler.stop()
    prof = engine.profile_read()
    counts1 = engine.last_solve_stats()["fine_launches_total"]
    true_counts = {n: counts1[n] - counts0[n] for n in counts1}
    launches = engine.launch_count() - launches0
    # one extra, untimed iteration with every multigrid level instrumented (diagnostic only)
    engine.set_option(_lib.OPT_PROFILE, 2)
    prev.copy_(psi)
    solver.step_device(prev, solver.step_size_at_iter(k), psi, rho)
    objectives.append(problem.calculate_objective(solver.rho))
    k += 1
    engine.profile_read()
    level_profile = getattr(engine, "last_level_profile", None)
    engine.set_option(_lib.OPT_PROFILE, 0)
    solves = problem.solve_log[log0:log0 + args.steps]
    pcg_iters = sum(s["iterations"] for s in solves)
    fine_applies = sum(s["fine_applies"] for s in solves)

    t = torch.tensor([elapsed_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    raw_rate = args.steps / (elapsed_ms * 1e-3)
    # whole-job aggregate: iterations/s normalised to the 1-GPU mesh (x global dofs / 1-GPU dofs)
    value = size_factor * raw_rate

    # ---- end to end through the reference-facing hooks with HOST buffers (numpy in/out):
    # Solver.step + calculate_objective of src/solver.py.
    # FEMSolver's numpy hooks count their own PCIe traffic (pinned staging inside _h2d/_d2h).
    # Per step: psi goes up, psi_new comes down, the objective comes down; rho = expit(psi_new) is
    # left on the device by FEMSolver.step (the reference loop's host-side expit + upload is not
    # needed to evaluate the next objective).
    # The leg repeats exactly the mirror-descent iterations of the device-timed region (same
    # starting design, same PCG work).  Untimed first: the state at that point is re-established
    # and the pinned staging buffers are allocated.
    psi_host = None
    if not args.no_e2e:
        k = k_at_start
        rho.copy_(rho_at_start)
        objectives.append(problem.calculate_objective(solver.rho))
        psi_host = solver.to_array(Function_like(psi_at_start, solver))
        solver._h2d(psi_host)
    solver.h2d_bytes = solver.d2h_bytes = 0
    barrier()
    t0 = time.perf_counter()
    for _ in range(0 if args.no_e2e else args.steps):
        psi_host = solver.step(psi_host, solver.step_size_at_iter(k))
        objectives.append(problem.calculate_objective(solver.rho))
        k += 1
    barrier()
    e2e_s = time.perf_counter() - t0
    traffic = {"h2d": solver.h2d_bytes, "d2h": solver.d2h_bytes}
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = None if args.no_e2e else size_factor * args.steps / float(t.item())

    # ---- reported separately: the same iterations with the multigrid preconditioner in fp32 inside
    # the fp64 PCG (TM option 117).  Residual test, search directions and the converged displacement
    # stay fp64 (parity: test_mixed_precision_preconditioner_keeps_fp64_accuracy); the headline
    # `value` above is the all-fp64 run.
    mixed_leg = None
    if world == 1 and esize == 8 and not args.mixed and not args.no_mixed_leg and args.preconditioner == "multigrid":
        engine.set_option(117, 1)
        k = k_at_start
        psi.copy_(psi_at_start)
        rho.copy_(rho_at_start)
        objectives_mixed = [problem.calculate_objective(solver.rho)]  # untimed: builds the fp32 hierarchy
        log1 = len(problem.solve_log)
        barrier()
        start.record()
        for _ in range(args.steps):
            prev.copy_(psi)
            solver.step_device(prev, solver.step_size_at_iter(k), psi, rho)
            objectives_mixed.append(problem.calculate_objective(solver.rho))
            k += 1
        stop.record()
        barrier()
        t = torch.tensor([start.elapsed_time(stop)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        engine.set_option(117, 0)
        msolves = problem.solve_log[log1:log1 + args.steps]
        ref_obj = objectives[args.warmup:args.warmup + args.steps + 1]
        mixed_leg = {
            "value": size_factor * args.steps / (float(t.item()) * 1e-3), "unit": UNIT,
            "ms_per_step": float(t.item()) / args.steps,
            "pcg_iterations_by_solve": [s_["iterations"] for s_ in msolves],
            "last_relative_residual": msolves[-1]["relative_residual"] if msolves else None,
            "max_relative_objective_difference_vs_fp64_run": max(
                abs(a - b) / abs(b) for a, b in zip(objectives_mixed, ref_obj)) if ref_obj else None,
            "note": "fp32 V-cycle inside the fp64 PCG; NOT the headline value",
        }

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel: the fine-level operator with a fused epilogue.
    # Algorithmic bytes per launch (every lattice vector once), with the
    # default fine smoother degree 1: Chebyshev step  reads x, b, D^-1, xi, writes x_new (the
    # direction d is neither read, c1 = 0, nor stored, last step); residual with the fused first
    # step  reads b, D^-1, xi, writes x, r;  p.Ap  reads p, xi, writes Ap.
    peak, peak_src = measured_peak_hbm()
    es_mg = 4 if (args.mixed and esize == 8) else esize  # multigrid-side launches run in fp32 when mixed
    alg_bytes = {
        "cheb": (4 * nu + n1) * es_mg, "resid": (4 * nu + n1) * es_mg,
        "dot": (2 * nu + n1) * esize, "plain": (2 * nu + n1) * es_mg,
    }
    # Event-timed launches are a sample (V-cycles replayed from a CUDA graph are not individually
    # timed): time per epilogue = sampled mean x true launch count in the timed region
    est_ms = {n: (p["ms"] / p["launches"] * true_counts[n]) if p["launches"] else 0.0 for n, p in prof.items()}
    dominant = max(est_ms, key=est_ms.get)
    d = prof[dominant]
    avg_ms = d["ms"] / max(d["launches"], 1)
    achieved = alg_bytes[dominant] / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
    kernel_ms = sum(est_ms.values())
    traffic_bytes, traffic_src = (None, None)
    if world == 1 and not args.mixed:
        traffic_bytes, traffic_src = measured_traffic(os.path.splitext(os.path.basename(design_path))[0], run_n,
                                                      args.dtype, dominant)
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": traffic_bytes, "traffic_source": traffic_src, "peak_source": peak_src,
        "kernel": f"elast_apply_kernel<{'double' if esize == 8 else 'float'},xi,EP_{dominant.upper()}>",
        "algorithmic_bytes_per_launch": alg_bytes[dominant], "avg_launch_ms": avg_ms,
        "launches_in_timed_region": true_counts[dominant], "launches_event_timed": d["launches"],
        "share_of_step_time": kernel_ms / elapsed_ms,
        "operator_ms_by_multigrid_level_one_untimed_step": level_profile,
        "per_epilogue": {n: {"ms_sampled": p["ms"], "launches_sampled": p["launches"],
                             "launches_true": true_counts[n], "ms_estimated": est_ms[n],
                             "GBps": (alg_bytes[n] * p["launches"] / (p["ms"] * 1e-3) / 1e9) if p["ms"] > 0 else None}
                         for n, p in prof.items()},
    }

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        sample_n = args.sample_n or 256
        times, s = oracle_iteration_seconds(design_path, sample_n, 1, 0)
        cpu_baseline = {
            "value": 1.0 / times[0], "unit": UNIT, "cores": 1, "kind": "port",
            "host_cores_available": os.cpu_count(),
            "sample": (f"1 mirror-descent iteration of the scipy oracle (CSR assembly + SuperLU, nested "
                       f"dissection) on designs/{args.design}.json at N={sample_n} (nx={s.mesh.nx}, "
                       f"ny={s.mesh.ny}, {s.mesh.nu} displacement dofs): a bounded sample, "
                       f"{(nx * ny) / (s.mesh.nx * s.mesh.ny):.1f}x fewer cells than the GPU workload"),
            "split_seconds": {k2: round(v, 3) for k2, v in s.problem.timings.items()},
        }
        try:
            cpu_baseline["state_solve_kernel"] = omp_kernel_baseline(design_path, args.N)
        except Exception as exc:
            cpu_baseline["state_solve_kernel"] = {"unavailable": repr(exc)}

    cfg = workload_description(args.design, run_n, nx, ny)
    cfg.update({
        "preconditioner": args.preconditioner, "state_rtol": args.state_rtol,
        "parallelism": "1 GPU" if world == 1 else
        f"{world} GPUs, row strips of cells, " + (
            "halo rows and scalar sums by the library's own kernels over peer-mapped NVLink windows (TM_OPT_P2P)"
            if engine.peer_memory_active else "NCCL halo exchange + all-reduce") +
        f", {engine.dist_levels} sharded multigrid levels",
        "weak_scaling_N": run_n, "value_normalisation": f"iter/s x (global dofs / dofs of the N={args.N} mesh) = x{size_factor:.3f}",
        "l2": f"working set of a state solve ~{10 * nu * esize / 1e6:.0f} MB of lattice vectors, larger than the 126 MB L2",
        "md_iterations_timed": [args.warmup, args.warmup + args.steps],
    })
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": ("f64 (fp32 multigrid preconditioner)" if args.mixed else "f64") if esize == 8 else "f32",
        "data": "synthetic", "config": cfg, "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT,
                "h2d_bytes_per_step": traffic["h2d"] // args.steps,
                "d2h_bytes_per_step": traffic["d2h"] // args.steps},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "raw_iters_per_sec": raw_rate,
        "mixed_precision_preconditioner": mixed_leg,
        "pcg": {"iterations_per_step": pcg_iters / args.steps,
                "dof_iters_per_sec": pcg_iters * nu_global / (elapsed_ms * 1e-3),
                "fine_operator_applies_per_step": fine_applies / args.steps,
                "state_solves": len(solves),
                "warm_starts_kept": sum(1 for s in solves if s.get("warm_start_used")),
                "iterations_by_solve": [s["iterations"] for s in solves],
                "last_relative_residual": solves[-1]["relative_residual"] if solves else None},
        "objective_trace": objectives[: args.warmup + args.steps + 1],
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=("cuda", "reference"), default="cuda")
    ap.add_argument("--design", default=DESIGN)
    ap.add_argument("--N", type=int, default=FULL_N)
    ap.add_argument("--dtype", choices=("float64", "float32"), default="float64")
    ap.add_argument("--preconditioner", choices=("multigrid", "jacobi"), default="multigrid")
    ap.add_argument("--state_rtol", type=float, default=1e-10)
    ap.add_argument("--sample_n", type=int, default=0, help="resolution of the CPU baseline sample")
    ap.add_argument("--no_cpu_baseline", action="store_true")
    ap.add_argument("--dist_levels", type=int, default=0, help="sharded multigrid levels (0 = automatic)")
    ap.add_argument("--mixed", action="store_true", help="fp32 multigrid preconditioner inside the fp64 PCG (reported separately)")
    ap.add_argument("--engine_option", action="append", default=[], help="KEY=VALUE passed to tm_set_option (tuning studies)")
    ap.add_argument("--exact_N", action="store_true", help="multi-GPU: run exactly --N (strong scaling of a named config)")
    ap.add_argument("--no_warm_start", action="store_true", help="state solves start from zero (study)")
    ap.add_argument("--no_mixed_leg", action="store_true", help="skip the separately reported fp32-preconditioner leg")
    ap.add_argument("--no_e2e", action="store_true", help="skip the host-buffer end-to-end leg (very large meshes)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "cuda":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_cuda_arm(args)


if __name__ == "__main__":
    main()
