"""Sharded path vs single-GPU path on the same inputs.  Launched by tests/test_gpu_sharded.py as
    python -m torch.distributed.run --nproc-per-node {2,4,8} tests/dist_check.py
Every rank builds the global inputs from the same seed, runs the sharded engine on its strip and
the gathered result is compared with an unsharded engine on rank 0's GPU."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from topomax_b200 import _lib  # noqa: E402
from topomax_b200.designs.definitions import Side  # noqa: E402
from topomax_b200.engine import Engine  # noqa: E402
from topomax_b200 import sharding as sh  # noqa: E402


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl")
    report = {}
    want_p2p = os.environ.get("TM_P2P") == "1"
    p2p_seen = []
    # cell rows scale with the rank count so that every strip keeps >= 2^dist_levels rows (world 2: the
    # meshes of round 1: 40x32, 70x52, 96x64)
    for case, (nx, ny, fixed, ld) in enumerate([(40, 16 * world, [Side.LEFT], 2),
                                                (70, 26 * world, [Side.BOTTOM, Side.TOP], 1),
                                                (96, 32 * world, [Side.LEFT, Side.RIGHT], 3)]):
        W, H = 0.25 * nx, 0.25 * ny
        kw = dict(lame_lambda=1.3, lame_mu=0.8, filter_radius=0.3, fixed_sides=fixed)
        eng = Engine(nx, ny, W, H, rank=rank, nranks=world, dist_levels=ld, **kw)
        eng.init_comm()
        p2p_seen.append(eng.peer_memory_active)
        ref = Engine(nx, ny, W, H, **kw)
        rng = np.random.default_rng(100 + case)
        n1, nu = (nx + 1) * (ny + 1), 2 * (2 * nx + 1) * (2 * ny + 1)
        xi_g, x_g, rho_g = 0.05 + 0.9 * rng.random(n1), rng.standard_normal(nu), rng.random(n1)
        rhs_g = rng.standard_normal(n1)
        t = lambda a: torch.as_tensor(a, dtype=torch.float64).cuda()
        tag = f"case{case}"

        # operator and diagonal
        y = sh.gather_p2(eng, eng.elast_matvec(sh.local_p1(eng, xi_g), sh.local_p2(eng, x_g)))
        y_ref = ref.elast_matvec(t(xi_g), t(x_g)).cpu().numpy()
        report[tag + "_matvec"] = rel(y, y_ref)
        d = sh.gather_p2(eng, eng.elast_diag_inverse(sh.local_p1(eng, xi_g)))
        report[tag + "_diag"] = rel(d, ref.elast_diag_inverse(t(xi_g)).cpu().numpy())

        # filter, both right-hand-side kinds
        f, info = eng.filter_apply(sh.local_p1(eng, rho_g), assembled=False, rtol=1e-13)
        f_ref, _ = ref.filter_apply(t(rho_g), assembled=False, rtol=1e-13)
        report[tag + "_filter0"] = rel(sh.gather_p1(eng, f), f_ref.cpu().numpy())
        g, _ = eng.filter_apply(sh.local_p1(eng, rhs_g), assembled=True, rtol=1e-13)
        g_ref, _ = ref.filter_apply(t(rhs_g), assembled=True, rtol=1e-13)
        report[tag + "_filter1"] = rel(sh.gather_p1(eng, g), g_ref.cpu().numpy())

        # loads, state solve (multigrid and Jacobi), compliance, sensitivity
        from topomax_b200.designs.definitions import CircularRegion, Force, Traction
        force = Force(CircularRegion((0.6 * W, 0.5 * H), 0.2 * H), (0.0, -1.0))
        tr = [Traction(Side.TOP, 0.5 * W, 0.3 * W, (0.0, -3.0))]
        b = eng.load_vector(force, tr)
        b_ref = ref.load_vector(force, tr)
        report[tag + "_load"] = rel(sh.gather_p2(eng, b), b_ref.cpu().numpy())
        for name, pre in (("mg", _lib.PRECOND_MULTIGRID), ("jacobi", _lib.PRECOND_JACOBI)):
            eng.set_option(_lib.OPT_PRECOND, pre)
            ref.set_option(_lib.OPT_PRECOND, pre)
            u, info = eng.state_solve(sh.local_p1(eng, xi_g), b, rtol=1e-11)
            u_ref, info_ref = ref.state_solve(t(xi_g), b_ref, rtol=1e-11)
            ug = sh.gather_p2(eng, u)
            report[f"{tag}_solve_{name}"] = float(np.linalg.norm(ug - u_ref.cpu().numpy()) /
                                                  np.linalg.norm(u_ref.cpu().numpy()))
            report[f"{tag}_iters_{name}"] = [info.iterations, info_ref.iterations]
        c, c_ref = eng.dot_p2(u, b), ref.dot_p2(u_ref, b_ref)
        report[tag + "_compliance"] = abs(c - c_ref) / abs(c_ref)
        s = eng.sens_rhs(sh.local_p1(eng, xi_g), u)
        s_ref = ref.sens_rhs(t(xi_g), u_ref)
        report[tag + "_sens"] = rel(sh.gather_p1(eng, s), s_ref.cpu().numpy())

        # general SIMP exponent (level-0 stored moments on the local strip): operator, multigrid
        # solve and sensitivity, then back to p = 3 on the same engines
        if case == 2:
            eng.set_option(_lib.OPT_PRECOND, _lib.PRECOND_MULTIGRID)
            ref.set_option(_lib.OPT_PRECOND, _lib.PRECOND_MULTIGRID)
            y = sh.gather_p2(eng, eng.elast_matvec(sh.local_p1(eng, xi_g), sh.local_p2(eng, x_g), penalty=2.0))
            report[tag + "_matvec_p2"] = rel(y, ref.elast_matvec(t(xi_g), t(x_g), penalty=2.0).cpu().numpy())
            u2, info2 = eng.state_solve(sh.local_p1(eng, xi_g), b, 2.5, rtol=1e-11)
            u2_ref, info2_ref = ref.state_solve(t(xi_g), b_ref, 2.5, rtol=1e-11)
            report[tag + "_solve_mg_p25"] = float(np.linalg.norm(sh.gather_p2(eng, u2) - u2_ref.cpu().numpy()) /
                                                  np.linalg.norm(u2_ref.cpu().numpy()))
            report[tag + "_iters_mg_p25"] = [info2.iterations, info2_ref.iterations]
            s2 = eng.sens_rhs(sh.local_p1(eng, xi_g), u2, penalty=2.5)
            report[tag + "_sens_p25"] = rel(sh.gather_p1(eng, s2), ref.sens_rhs(t(xi_g), u2_ref, penalty=2.5).cpu().numpy())
            y = sh.gather_p2(eng, eng.elast_matvec(sh.local_p1(eng, xi_g), sh.local_p2(eng, x_g)))
            report[tag + "_matvec_back_to_p3"] = rel(y, y_ref)

        # mirror-descent reductions
        half = sh.local_p1(eng, rhs_g)
        v, dv = eng.md_volume(half, 0.2)
        v_ref, dv_ref = ref.md_volume(t(rhs_g), 0.2)
        report[tag + "_volume"] = max(abs(v - v_ref) / abs(v_ref), abs(dv - dv_ref) / abs(dv_ref))
        report[tag + "_integrate"] = abs(eng.integrate(half) - ref.integrate(t(rhs_g))) / abs(ref.integrate(t(rhs_g)))
        del eng, ref
    # the whole optimiser, sharded vs unsharded (device loop and the numpy-hook loop)
    from topomax_b200.fem_solver import FEMSolver
    design = os.path.join(ROOT, "designs", "cantilever.json")
    N = 40 if world <= 4 else 64  # 40 cell rows cannot be cut into 8 strips of multiples of 4 rows
    report["solver_N"] = N
    sharded = FEMSolver(N, design, data_path=f"/tmp/tm_dist_{rank}", verbose=False, distributed=True, dist_levels=2)
    rs = sharded.solve(fixed_iterations=4)
    rho_s = sharded.to_array(sharded.rho)
    single = FEMSolver(N, design, data_path=f"/tmp/tm_single_{rank}", verbose=False)
    r1 = single.solve(fixed_iterations=4)
    report["solver_objectives"] = float(max(abs(a - b) / abs(b) for a, b in zip(rs["objectives"], r1["objectives"])))
    report["solver_rho"] = float(np.abs(rho_s - single.to_array(single.rho)).max())
    hooks = FEMSolver(N, design, data_path=f"/tmp/tm_hooks_{rank}", skip_multiple=999, verbose=False, distributed=True,
                      dist_levels=2)
    hooks.solve_generic()
    full = FEMSolver(N, design, data_path=f"/tmp/tm_full_{rank}", skip_multiple=999, verbose=False)
    full.solve()
    report["hooks_k"] = [hooks.last_result["k_final"], full.last_result["k_final"]]
    report["hooks_rho"] = float(np.abs(hooks.to_array(hooks.rho) - full.to_array(full.rho)).max())
    if want_p2p:  # the peer-memory path must really have been the one that ran
        assert all(p2p_seen), p2p_seen
    else:
        assert not any(p2p_seen), p2p_seen
    if rank == 0:
        report["files"] = sorted(os.listdir(f"/tmp/tm_hooks_0/FEM/cantilever/data"))[:3]
        print("DIST_REPORT " + json.dumps(report))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
