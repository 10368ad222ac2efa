"""Wall-clock split of mirror-descent iterations into their host-visible phases (every phase synchronised: a
diagnostic, not a bench).   python tools/step_phases.py design N warmup steps [opt=value ...]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from topomax_b200.fem_solver import FEMSolver  # noqa: E402


def main():
    design, N, warmup, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    s = FEMSolver(N, os.path.join(ROOT, "designs", f"{design}.json"), data_path="/tmp/tm_study", verbose=False)
    pr, e = s.problem, s.problem.engine
    for kv in sys.argv[5:]:
        k, v = kv.split("=")
        e.set_option(int(k), float(v))
    pr.set_penalization(3.0)
    rho = s.rho.tensor
    psi = torch.log(rho / (1 - rho))
    prev = torch.empty_like(psi)
    pr.calculate_objective(s.rho)
    acc = {}

    def tick(name, t0):
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        acc[name] = acc.get(name, 0.0) + (t1 - t0) * 1e3
        return t1

    for k in range(warmup + steps):
        if k == warmup:
            acc.clear()
            torch.cuda.synchronize()
            t_all = time.perf_counter()
        torch.cuda.synchronize()
        t = time.perf_counter()
        prev.copy_(psi)
        s.step_device(prev, s.step_size_at_iter(k), psi, rho)
        t = tick("step_device (sensitivity, its filter, projection)", t)
        pr.filtered_rho = pr.filter.apply(s.rho)
        t = tick("density filter", t)
        pr.u = pr.forward(pr.filtered_rho)
        t = tick("state solve", t)
        e.dot_p2(pr.u.tensor, pr.load)
        t = tick("compliance", t)
    total = (time.perf_counter() - t_all) * 1e3
    print(json.dumps({"design": design, "N": N, "steps": steps, "ms_per_step": round(total / steps, 3),
                      "phases_ms_per_step": {k: round(v / steps, 3) for k, v in acc.items()},
                      "pcg_iterations": [d["iterations"] for d in pr.solve_log[-steps:]],
                      "cycle_window": e.last_solve_stats().get("cycle_window")}))


if __name__ == "__main__":
    main()
