"""Host logic of the row-strip sharded path on CPU (gloo, world_size 2): the partition
arithmetic mirrored from the library, slicing global fields into strips with halo rows, and the
gather of owned rows.  The CUDA side of sharding is covered by tests/test_gpu_sharded.py."""
import os
import socket
from types import SimpleNamespace

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from topomax_b200 import sharding as sh  # noqa: E402


def fake_engine(nx, ny, rank, nranks, dist_levels):
    starts = sh.partition_rows(ny, nranks, dist_levels)
    c0, c1 = starts[rank], starts[rank + 1]
    last = rank == nranks - 1
    return SimpleNamespace(nx=nx, ny=ny, rank=rank, nranks=nranks, c0=c0, c1=c1, cl0=max(0, c0 - 2),
                           cl1=min(ny, c1 + 1), owns_top=last, dtype=torch.float64, device=torch.device("cpu"),
                           owned_p1_rows=lambda: (c0 - max(0, c0 - 2), c1 - max(0, c0 - 2) + (1 if last else 0), c0),
                           owned_p2_rows=lambda: (2 * (c0 - max(0, c0 - 2)),
                                                  2 * (c1 - max(0, c0 - 2)) + (1 if last else 0), 2 * c0))


@pytest.mark.parametrize("ny,nranks,ld", [(32, 2, 2), (510, 2, 3), (52, 4, 1), (16384, 8, 5), (2048, 8, 4)])
def test_partition_is_aligned_and_covers(ny, nranks, ld):
    starts = sh.partition_rows(ny, nranks, ld)
    assert starts[0] == 0 and starts[-1] == ny and len(starts) == nranks + 1
    assert all(b > a for a, b in zip(starts, starts[1:]))
    assert all(s % (1 << ld) == 0 for s in starts[:-1])  # strips nest across the sharded levels
    # every middle rank owns at least two cell rows on every sharded level (halo depth)
    for l in range(ld):
        for r in range(nranks - 1):
            assert (starts[r + 1] >> l) - (starts[r] >> l) >= 2


def test_partition_rejects_too_many_ranks():
    with pytest.raises(ValueError):
        sh.partition_rows(8, 4, 2)


def _worker(rank, world, port, nx, ny, ld, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        eng = fake_engine(nx, ny, rank, world, ld)
        rng = np.random.default_rng(7)
        g1 = rng.random((ny + 1) * (nx + 1))
        g2 = rng.random(2 * (2 * ny + 1) * (2 * nx + 1))
        l1, l2 = sh.local_p1(eng, g1), sh.local_p2(eng, g2)
        assert l1.numel() == (eng.cl1 - eng.cl0 + 1) * (nx + 1)
        assert l2.numel() == 2 * (2 * (eng.cl1 - eng.cl0) + 1) * (2 * nx + 1)
        # halo rows carry the neighbour's values
        assert np.array_equal(l1.numpy().reshape(-1, nx + 1)[0], g1.reshape(ny + 1, nx + 1)[eng.cl0])
        back1, back2 = sh.gather_p1(eng, l1), sh.gather_p2(eng, l2)
        ok = np.array_equal(back1, g1) and np.array_equal(back2, g2)
        # a collective scalar, as the engine's all-reduced integrals behave
        lo, hi, _ = eng.owned_p1_rows()
        part = torch.tensor([l1.reshape(-1, nx + 1)[lo:hi].sum().item()], dtype=torch.float64)
        dist.all_reduce(part)
        ok = ok and abs(part.item() - g1.sum()) < 1e-9
        if rank == 0:
            out.put(bool(ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nx,ny,ld", [(12, 32, 2), (9, 27, 1)])
def test_strip_roundtrip_world_size_2(nx, ny, ld):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, nx, ny, ld, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=5) is True
