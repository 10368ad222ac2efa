"""GPU study: PCG/multigrid behaviour along a real optimisation run (not a test, not a bench).
    python tools/mg_study.py design N iterations [key=value ...]   -> JSON lines on stdout
keys: degree, ratio, safety, rtol, precond"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from topomax_b200 import _lib  # noqa: E402
from topomax_b200.fem_solver import FEMSolver  # noqa: E402


def main():
    design, N, iters = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    opts = dict(kv.split("=") for kv in sys.argv[4:])
    s = FEMSolver(N, os.path.join(ROOT, "designs", f"{design}.json"), data_path="/tmp/tm_study",
                  verbose=False, problem_options={"preconditioner": opts.get("precond", "multigrid"),
                                                  "state_rtol": float(opts.get("rtol", 1e-10)),
                                                  "warm_start": opts.get("warm", "1") == "1"})
    e = s.problem.engine
    if "degree" in opts:
        e.set_option(_lib.OPT_CHEB_DEGREE, float(opts["degree"]))
    if "ratio" in opts:
        e.set_option(_lib.OPT_CHEB_RATIO, float(opts["ratio"]))
    if "safety" in opts:
        e.set_option(_lib.OPT_EIG_SAFETY, float(opts["safety"]))
    s.problem.set_penalization(3.0)
    rho = s.rho.tensor
    psi = torch.log(rho / (1 - rho))
    prev = torch.empty_like(psi)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    obj = s.problem.calculate_objective(s.rho)
    torch.cuda.synchronize()
    rows = [dict(k=0, obj=obj, t=time.perf_counter() - t0, **s.problem.solve_log[-1],
                 filter_its=s.problem.filter.last_info.iterations)]
    for k in range(iters):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        prev.copy_(psi)
        d = s.step_device(prev, s.step_size_at_iter(k), psi, rho)
        fit_g = s.problem.filter.last_info.iterations
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        obj = s.problem.calculate_objective(s.rho)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        rows.append(dict(k=k + 1, obj=obj, delta=d, t_step=t1 - t0, t_obj=t2 - t1, **s.problem.solve_log[-1],
                         filter_its=s.problem.filter.last_info.iterations, filter_its_grad=fit_g,
                         rho_min=float(rho.min()), rho_max=float(rho.max())))
    print(json.dumps(dict(design=design, N=N, nx=s.mesh.nx, ny=s.mesh.ny, opts=opts, rows=rows)))


if __name__ == "__main__":
    main()
