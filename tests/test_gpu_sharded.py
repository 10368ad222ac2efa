"""Row-strip sharded path (2 GPUs, NCCL halos + all-reduce + replicated coarse multigrid levels)
against the single-GPU path.  Skipped when fewer than 2 GPUs are visible."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def report_k(report):
    return 36  # cantilever N=40 converges at iteration 36 (oracle anchor)


@pytest.mark.parametrize("peer_memory", [False, True], ids=["nccl", "peer_memory"])
def test_sharded_matches_single_gpu(repo_root, peer_memory):
    """peer_memory: halo rows and scalar sums through the library's own kernels over peer-mapped
    windows (csrc/tm_p2p.cuh, TM_P2P=1) instead of NCCL calls; same checks, same tolerances."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    if peer_memory and os.environ.get("TM_TEST_P2P") != "1":
        # the kernels pass the single-GPU loop-back (test_gpu_z_p2p_loopback.py); the IPC mapping and
        # the engine's dispatch have not run on 2 GPUs yet (the multi-GPU budget of round 1 was spent)
        pytest.skip("cross-GPU peer-memory transport not yet run on hardware: opt in with TM_TEST_P2P=1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29534" if peer_memory else "29533",
           os.path.join(repo_root, "tests", "dist_check.py")]
    env = dict(os.environ, TM_P2P="1" if peer_memory else "0")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=repo_root, env=env)
    line = [l for l in out.stdout.splitlines() if l.startswith("DIST_REPORT ")]
    assert line, out.stdout[-2000:] + out.stderr[-4000:]
    report = json.loads(line[0][len("DIST_REPORT "):])
    print(report)
    assert report.pop("hooks_k")[0] == report_k(report)
    assert len(report.pop("files")) == 3
    for key, val in report.items():
        if key.startswith("solver_") or key.startswith("hooks_"):
            assert val < 1e-7, (key, val)
        elif "_iters_" in key:
            # same preconditioner up to round-off; PCG counts may drift by a few near the tolerance
            assert abs(val[0] - val[1]) <= max(3, 0.12 * val[1]), (key, val)
        elif "_solve_" in key or "compliance" in key or "filter" in key or "sens" in key:
            assert val < 1e-8, (key, val)
        else:
            assert val < 1e-12, (key, val)
