"""GPU study: PCG iterations and time of ONE state solve on the design a real run reaches after `iters`
mirror-descent iterations, for multigrid cycle windows (engine options 133-135) -- the same warm start for all.
    python tools/cycle_study.py design N iters "133=5,134=7,135=2" "119=0;133=5,135=2" ...
An option set is a comma-separated list of option=value; the first run is always the library default."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from topomax_b200.fem_solver import FEMSolver  # noqa: E402


def main():
    design, N, iters = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    sets = [""] + sys.argv[4:]
    s = FEMSolver(N, os.path.join(ROOT, "designs", f"{design}.json"), data_path="/tmp/tm_study", verbose=False)
    pr = s.problem
    e = pr.engine
    pr.set_penalization(3.0)
    rho = s.rho.tensor
    psi = torch.log(rho / (1 - rho))
    prev = torch.empty_like(psi)
    pr.calculate_objective(s.rho)
    u_prev = None
    for k in range(iters):
        prev.copy_(psi)
        if k == iters - 1:
            u_prev = pr.u.tensor.clone()
        s.step_device(prev, s.step_size_at_iter(k), psi, rho)
        pr.calculate_objective(s.rho)
    xi = pr.filtered_rho.tensor.clone()
    stats = e.last_solve_stats()
    print(json.dumps({"design": design, "N": N, "iters": iters, "levels": stats.get("levels"),
                      "tail_first_level": stats.get("tail_first_level"), "cycle_window": stats.get("cycle_window"), "run_pcg_iterations": [d["iterations"] for d in pr.solve_log]}))
    for spec in sets:
        applied = []
        for kv in filter(None, spec.replace(";", ",").split(",")):
            k, v = kv.split("=")
            e.set_option(int(k), float(v))
            applied.append((int(k), float(v)))
        row = {"options": spec or "default"}
        for rep in range(2):
            u0 = u_prev.clone()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            t0.record()
            u, info = e.state_solve(xi, pr.load, 3.0, rtol=pr.state_rtol, maxit=2000, u=u0, warm_start=True)
            t1.record()
            torch.cuda.synchronize()
            row[f"ms_{rep}"] = round(t0.elapsed_time(t1), 3)
            row["iterations"] = info.iterations
            row["relres"] = info.relative_residual
        row["compliance"] = float(e.dot_p2(u, pr.load))
        print(json.dumps(row), flush=True)
        # back to the defaults for the next set
        for k, _ in applied:
            e.set_option(k, {119: 2304, 133: 1, 134: 1 << 20, 135: 0, 109: 3, 2: 1, 4: 4, 136: -1}.get(k, 0))


if __name__ == "__main__":
    main()
