"""A step of the hot path as a CUDA graph.

The launches of a C-ABI call are fixed by its arguments -- everything that depends on the data is decided on the
device (which sets have function tables, which samples are in transit, whether the fallback kernels have anything
to do) -- and a call never synchronises, so a call whose buffers stay where they are can be captured ONCE by stream
capture and replayed: one driver call per step instead of ~15, and the dependencies between its kernels (the table
launch on the side stream included) become graph edges.  On config 3: 0.267 -> 0.251 ms per step on one GPU, 0.207
-> 0.166 ms on two, where the host no longer keeps up with stream-ordered launches.  Under XLA the
custom calls sit in command buffers the same way; this helper is for callers that drive the C ABI from torch.
"""

from __future__ import annotations

import torch


def capture(fn, after_capture=None) -> torch.cuda.CUDAGraph | None:
    """Run ``fn()`` once on a side stream (warm-up: lazy initialisation must not happen inside a capture), then
    capture a second call into a graph.  ``fn`` may only launch work on the current stream (every entry point of
    libjxp_b200 qualifies; host-side staging of new inputs does not).  Returns the graph -- ``graph.replay()``
    repeats the call on the buffers it was captured with -- or None if the capture failed.  ``after_capture()``
    runs in either case (e.g. PeerGather.sync_epoch: the captured launch was counted on the host but has not run)."""
    cs = torch.cuda.Stream()
    cs.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(cs):
        fn()
    torch.cuda.current_stream().wait_stream(cs)
    torch.cuda.synchronize()
    graph = None
    try:
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=cs):
            fn()
    except Exception:  # noqa: BLE001 -- the caller falls back to stream-ordered launches
        graph = None
    if after_capture is not None:
        torch.cuda.synchronize()
        after_capture()
    return graph
