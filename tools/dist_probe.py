"""2-GPU probe: PCG iteration counts of the sharded vs unsharded multigrid solve at several tolerances."""
import json, os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from topomax_b200.designs.definitions import Side, CircularRegion, Force
from topomax_b200.engine import Engine
from topomax_b200 import sharding as sh

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl")
nx, ny, ld = 96, 64, 3
W, H = 0.25 * nx, 0.25 * ny
kw = dict(lame_lambda=1.3, lame_mu=0.8, filter_radius=0.3, fixed_sides=[Side.LEFT, Side.RIGHT])
eng = Engine(nx, ny, W, H, rank=rank, nranks=world, dist_levels=ld, **kw); eng.init_comm()
ref = Engine(nx, ny, W, H, **kw)
rng = np.random.default_rng(102)
xi_g = 0.05 + 0.9 * rng.random((nx + 1) * (ny + 1))
force = Force(CircularRegion((0.6 * W, 0.5 * H), 0.2 * H), (0.0, -1.0))
b, b_ref = eng.load_vector(force, None), ref.load_vector(force, None)
out = {}
for opts in ({}, {2: 2}, {100: 10}, {110: 0}):
    for k, v in opts.items():
        eng.set_option(k, v); ref.set_option(k, v)
    pairs = []
    for rtol in (1e-4, 1e-6, 1e-8, 1e-10, 1e-12):
        u, i1 = eng.state_solve(sh.local_p1(eng, xi_g), b, rtol=rtol)
        u2, i2 = ref.state_solve(torch.as_tensor(xi_g).cuda(), b_ref, rtol=rtol)
        pairs.append((i1.iterations, i2.iterations, eng.last_solve_stats()["lambda_max"], ref.last_solve_stats()["lambda_max"]))
    out[str(opts)] = pairs
if rank == 0:
    print("PROBE " + json.dumps(out))
dist.barrier(); dist.destroy_process_group()
