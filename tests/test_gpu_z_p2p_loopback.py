"""The peer-memory halo-exchange / all-reduce kernels (csrc/tm_p2p.cuh) on ONE GPU: the "ranks" are
streams of one process and their windows plain allocations of the same device (``tm_p2p_selftest``),
so the hand-shake, the epoch bookkeeping, the last-block publication and the acquire/release idioms
run on real hardware without a second GPU.  The ranks run asynchronously for hundreds of epochs;
every received row and every sum is checked on the device.  (Across GPUs the same kernels are
compared with the NCCL transport in tests/test_gpu_sharded.py.)"""
import ctypes

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.mark.parametrize("nranks,row_elems", [(2, 1000), (2, 1001), (3, 257), (4, 4096)])
def test_loopback(nranks, row_elems):
    from topomax_b200 import _lib

    lib = _lib.load_library()
    torch.cuda.init()
    epochs, every = 300, 3
    report = (ctypes.c_double * 4)()
    with torch.cuda.device(0):
        _lib.check(lib.tm_p2p_selftest(nranks, epochs, row_elems, every, report))
    mismatches, timeouts, halo_epoch, red_epoch = list(report)
    assert timeouts == 0, list(report)
    assert mismatches == 0, list(report)
    assert halo_epoch == epochs and red_epoch == epochs // every
