"""Argument checking of the peer all-gather entry points (no GPU needed: every call fails before a launch)."""

import ctypes as C

import pytest


def test_peer_entry_points_check_their_arguments():
    from jaxoplanet_b200 import _lib

    lib = _lib.load()
    assert lib.jxp_peer_buffer_bytes(0) == 256 and lib.jxp_peer_buffer_bytes(4096) == 256 + 2 * 4096 * 8
    assert lib.jxp_peer_buffer_bytes(-1) == 0
    bufs = (C.c_uint64 * 2)(0, 0)
    src = C.c_void_p(8)
    for args in ((None, 1, 0, 2, C.addressof(bufs), 0, 2, 1, None, None),          # null source
                 (src, 1, 0, 2, None, 0, 2, 1, None, None),                        # null buffer table
                 (src, 1, 0, 2, C.addressof(bufs), 2, 2, 1, None, None),           # rank outside the world
                 (src, 1, 0, 2, C.addressof(bufs), 0, 17, 1, None, None),          # more ranks than the flag block holds
                 (src, 3, 0, 2, C.addressof(bufs), 0, 2, 1, None, None),           # block outside the total
                 (src, 1, 0, 2, C.addressof(bufs), 0, 2, 1, None, None)):          # no buffer mapped for rank 0
        with pytest.raises(_lib.JxpError):
            _lib.call("jxp_peer_allgather_f64", *args)
    with pytest.raises(_lib.JxpError):
        _lib.call("jxp_peer_open", None, None)
    assert lib.jxp_peer_close(None) == 0 and lib.jxp_peer_free(None) == 0


def test_graph_helper_and_peer_gather_are_importable_without_a_gpu():
    """Host-side modules of the multi-GPU / graph paths import on a machine without CUDA; nothing in them falls back
    to the CPU: peer_gather_for answers None outside an NCCL process group, capture needs a CUDA stream."""
    import torch

    from jaxoplanet_b200 import distributed, graphs

    assert callable(graphs.capture)
    assert distributed.peer_gather_for(16, torch.device("cpu")) is None
    x = torch.arange(5, dtype=torch.float64)
    assert distributed.gather_sets(x, 5) is x  # (no process group: the block is the whole)
