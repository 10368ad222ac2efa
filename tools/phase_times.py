"""GPU timing of the phases of one mirror-descent iteration (wall clock around synchronised
calls; a study tool, not the bench).  python tools/phase_times.py design N iters [opt=value ...]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from topomax_b200 import _lib  # noqa: E402
from topomax_b200.fem_solver import FEMSolver  # noqa: E402
from topomax_b200.filter import AssembledP1Form  # noqa: E402
from topomax_b200.solver import find_volume_shift  # noqa: E402


def timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return out, (time.perf_counter() - t0) * 1e3


def main():
    design, N, iters = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    opts = dict(kv.split("=") for kv in sys.argv[4:])
    s = FEMSolver(N, os.path.join(ROOT, "designs", f"{design}.json"), data_path="/tmp/tm_study", verbose=False,
                  problem_options={"state_rtol": float(opts.pop("rtol", 1e-10))})
    pr, e = s.problem, s.problem.engine
    for k, v in opts.items():
        e.set_option(int(k), float(v))
    pr.set_penalization(3.0)
    rho = s.rho.tensor
    psi = torch.log(rho / (1 - rho))
    prev = torch.empty_like(psi)
    pr.calculate_objective(s.rho)
    rows = []
    for k in range(iters):
        prev.copy_(psi)
        rhs, t_sens = timed(lambda: e.sens_rhs(pr.filtered_rho.tensor, pr.u.tensor, 3.0))
        G, t_fg = timed(lambda: pr.filter.apply(AssembledP1Form(s.control_space, rhs)))
        half = e.md_halfstep(prev, G.tensor, s.step_size_at_iter(k))
        c, t_proj = timed(lambda: find_volume_shift(lambda c: e.md_volume(half, c)[0] - s.volume,
                                                    lambda c: e.md_volume(half, c)[1]))
        _, t_apply = timed(lambda: e.md_apply(half, c, prev, psi, rho))
        xi, t_f = timed(lambda: pr.filter.apply(s.rho))
        pr.filtered_rho = xi
        u, t_solve = timed(lambda: pr.forward(xi))
        pr.u = u
        obj, t_dot = timed(lambda: e.dot_p2(u.tensor, pr.load))
        st = pr.solve_log[-1]
        rows.append(dict(k=k + 1, obj=obj, sens=t_sens, filt_grad=t_fg, proj=t_proj, apply=t_apply, filt=t_f,
                         solve=t_solve, dot=t_dot, its=st["iterations"], fits=pr.filter.last_info.iterations,
                         ms_per_it=t_solve / max(st["iterations"], 1)))
    for r in rows:
        print(json.dumps({k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items()}))


if __name__ == "__main__":
    main()
