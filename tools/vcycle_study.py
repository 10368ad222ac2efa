"""Where one V-cycle's time goes, measured as it runs inside a solve (graph replay): the V-cycle
is truncated at depth d (levels >= d replaced by smoothing only, engine option 121), so the
difference between consecutive depths is the cost of one level.  Study tool, not the bench.
python tools/vcycle_study.py design N [opt=value ...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from topomax_b200.fem_solver import FEMSolver  # noqa: E402


def main():
    design, N = sys.argv[1], int(sys.argv[2])
    opts = dict(kv.split("=") for kv in sys.argv[3:])
    s = FEMSolver(N, os.path.join(ROOT, "designs", f"{design}.json"), data_path="/tmp/tm_study", verbose=False)
    e = s.problem.engine
    for k, v in opts.items():
        e.set_option(int(k), float(v))
    xi = torch.full((e.n1,), 0.5, dtype=torch.float64, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(1)
    r = torch.rand(e.nu, dtype=torch.float64, device="cuda", generator=g)
    nl = len(e.mg_levels())
    rows = []
    for tail in (1, 0):
        if not tail:
            e.set_option(119, 0)
        prev = 0.0
        for depth in list(range(1, nl)) + [0]:
            e.set_option(121, depth)
            out = e.mg_debug(xi, 6, 0, r, 2).cpu().numpy()
            rows.append({"tail": tail, "depth": depth or nl, "vcycle_ms": round(float(out[0]), 4),
                         "delta_ms": round(float(out[0]) - prev, 4), "setup_ms": round(float(out[1]), 4)})
            prev = float(out[0])
    for row in rows:
        print(json.dumps(row))


if __name__ == "__main__":
    main()
