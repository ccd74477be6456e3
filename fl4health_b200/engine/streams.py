"""Side-stream fork/join for branch-level parallelism inside one training step.

A CIFAR-scale ResNet step on a B200 is a chain of ~130 short kernels that each occupy a fraction of the 148 SMs; the
weight-gradient GEMMs of the backward pass are off the critical path (only the optimizer consumes them), so they are
issued on a side stream and overlap the dgrad -> BN-backward chain.  Under CUDA-graph capture the fork/join becomes
parallel branches of the captured graph.

Protocol: ``fork()`` makes the side stream wait for everything issued so far on the current stream and returns it;
``defer_join(keepalive)`` records the side stream's progress; ``join()`` makes the current stream wait for all
deferred work.  A join is queued automatically at the end of the running backward pass (the autograd engine's
final-callback hook, the mechanism DDP uses), so ``loss.backward()`` returns with every gradient ordered on the
caller's stream no matter which client drives the step.
"""

from __future__ import annotations

import os
import threading
from typing import Any

import torch

_LOCK = threading.Lock()


class _State:
    """Process-wide (NOT thread-local): backward nodes run on the autograd engine's device thread while the engine's
    final callback -- and the optimizer -- run on the thread that called ``backward()``."""

    def __init__(self) -> None:
        self.side: dict[int, torch.cuda.Stream] = {}
        self.pending: list[tuple[torch.device, torch.cuda.Event, tuple[Any, ...]]] = []
        self.callback_queued = False


_STATE = _State()


def overlap_enabled() -> bool:
    return os.environ.get("FL4H_OVERLAP_WGRAD", "1") != "0"


def _state() -> _State:
    return _STATE


def side_stream(device: torch.device) -> torch.cuda.Stream:
    st = _state()
    index = device.index if device.index is not None else torch.cuda.current_device()
    with _LOCK:
        if index not in st.side:
            st.side[index] = torch.cuda.Stream(device=device)
        return st.side[index]


def branch_stream(device: torch.device) -> torch.cuda.Stream:
    """Second side stream for *forward* branch parallelism (e.g. a ResNet block's downsample path).  Autograd runs
    each backward node on the stream its forward ran on and inserts the cross-stream syncs itself, so a branch issued
    here overlaps the main path in both directions."""
    st = _state()
    index = device.index if device.index is not None else torch.cuda.current_device()
    with _LOCK:
        key = -1 - index
        if key not in st.side:
            st.side[key] = torch.cuda.Stream(device=device)
        return st.side[key]


def fork(device: torch.device) -> torch.cuda.Stream:
    side = side_stream(device)
    side.wait_stream(torch.cuda.current_stream(device))
    return side


def defer_join(device: torch.device, *keepalive: Any) -> None:
    """Call right after issuing side-stream work.  ``keepalive`` tensors are held until the join so the caching
    allocator cannot hand their memory to later main-stream kernels while the side stream still reads them."""
    st = _state()
    event = torch.cuda.Event()
    event.record(side_stream(device))
    with _LOCK:
        st.pending.append((device, event, keepalive))
        need_callback = not st.callback_queued
        st.callback_queued = True
    if need_callback:
        try:
            torch.autograd.Variable._execution_engine.queue_callback(join)  # runs when this backward pass finishes
        except RuntimeError:  # not inside a backward pass: the caller joins explicitly
            with _LOCK:
                st.callback_queued = False


def join() -> None:
    st = _state()
    with _LOCK:
        st.callback_queued = False
        pending, st.pending = st.pending, []
    for device, event, _ in pending:
        torch.cuda.current_stream(device).wait_event(event)


def pending_count() -> int:
    return len(_state().pending)
