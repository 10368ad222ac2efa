"""Total state-solve time over the first K mirror-descent iterations for a few smoother settings."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from topomax_b200 import _lib
from topomax_b200.fem_solver import FEMSolver

design, N, K = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
for spec in sys.argv[4:]:
    opts = dict(kv.split("=") for kv in spec.split(","))
    s = FEMSolver(N, os.path.join(ROOT, "designs", f"{design}.json"), data_path="/tmp/tm_study", verbose=False)
    e = s.problem.engine
    for k, v in opts.items():
        e.set_option(int(k), float(v))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = s.solve(fixed_iterations=K)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    its = [x["iterations"] for x in s.problem.solve_log]
    print(json.dumps(dict(spec=spec, total_s=round(dt, 3), pcg_total=sum(its), pcg_last=its[-5:], obj=r["objectives"][-1])))
