"""The hand-shake of the peer-memory halo exchange / all-reduce kernels (csrc/tm_p2p.cuh), modelled
with host threads and std::atomic release/acquire (tests/hostcheck/p2p_model.cpp).  CPU only.

What it pins: two mailboxes per direction (epoch parity) are enough WITHOUT acknowledgements, for
the halo exchange between neighbours and for the all-to-all reduction slots, under randomised
interleavings; a single mailbox is not (negative control).  The kernels themselves are compared
with the NCCL path on 2 GPUs in tests/test_gpu_sharded.py."""
import ctypes
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def model():
    build = os.path.join(HERE, "_build")
    os.makedirs(build, exist_ok=True)
    so = os.path.join(build, "libp2p_model.so")
    src = os.path.join(HERE, "hostcheck", "p2p_model.cpp")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-shared", "-fPIC", src, "-o", so])
    lib = ctypes.CDLL(so)
    lib.p2p_model_run.restype = ctypes.c_int
    lib.p2p_model_run.argtypes = [ctypes.c_int] * 5
    return lib


@pytest.mark.parametrize("ranks", [2, 3, 8])
def test_double_buffered_mailboxes_need_no_acknowledgement(model, ranks):
    for seed in range(4):
        assert model.p2p_model_run(ranks, 3000, seed, 2, 3) == 0


def test_single_mailbox_is_not_enough(model):
    """Negative control: with one mailbox per direction a fast neighbour overwrites rows that
    have not been unpacked yet (some seed shows it)."""
    bad = sum(model.p2p_model_run(4, 3000, seed, 1, 0) for seed in range(6))
    assert bad > 0
