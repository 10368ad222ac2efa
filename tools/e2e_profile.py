"""Host-side profile of the end-to-end step (numpy in/out through FEMSolver.step).  Study tool.
python tools/e2e_profile.py design N steps"""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from topomax_b200.fem_solver import FEMSolver  # noqa: E402


def main():
    design, N, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    s = FEMSolver(N, os.path.join(ROOT, "designs", f"{design}.json"), data_path="/tmp/tm_study", verbose=False)
    pr = s.problem
    pr.set_penalization(s.parameters.penalties[0])
    pr.calculate_objective(s.rho)
    rho = s.rho.tensor
    psi = torch.log(rho / (1 - rho))
    prev = torch.empty_like(psi)
    for k in range(3):
        prev.copy_(psi)
        s.step_device(prev, s.step_size_at_iter(k), psi, rho)
        pr.calculate_objective(s.rho)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(3, 3 + steps):
        prev.copy_(psi)
        s.step_device(prev, s.step_size_at_iter(k), psi, rho)
        pr.calculate_objective(s.rho)
    torch.cuda.synchronize()
    print("device loop ms/step", (time.perf_counter() - t0) / steps * 1e3)
    print("device loop PCG iterations", [e["iterations"] for e in pr.solve_log[-steps:]], "xi ptr",
          hex(pr.filtered_rho.tensor.data_ptr()), "u ptr", hex(pr.u.tensor.data_ptr()))
    psi_host = s.to_array(type("F", (), {"tensor": psi})())
    psi_host = s.step(psi_host, s.step_size_at_iter(10))
    pr.calculate_objective(s.rho)
    torch.cuda.synchronize()
    prof = cProfile.Profile()
    t0 = time.perf_counter()
    prof.enable()
    for k in range(11, 11 + steps):
        psi_host = s.step(psi_host, s.step_size_at_iter(k))
        pr.calculate_objective(s.rho)
        print("   xi", hex(pr.filtered_rho.tensor.data_ptr()), "its", pr.solve_log[-1]["iterations"])
    torch.cuda.synchronize()
    prof.disable()
    print("host loop ms/step", (time.perf_counter() - t0) / steps * 1e3)
    print("host loop PCG iterations", [e["iterations"] for e in pr.solve_log[-steps:]], "xi ptr",
          hex(pr.filtered_rho.tensor.data_ptr()), "u ptr", hex(pr.u.tensor.data_ptr()))
    pstats.Stats(prof).sort_stats("cumulative").print_stats(28)


if __name__ == "__main__":
    main()
