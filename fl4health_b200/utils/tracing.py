"""Phase tracing: NVTX ranges + device-side timing of the FL round phases (the reference only has
whole-second wall-clock deltas).

``with phase("local_train"):`` always pushes an NVTX range on CUDA builds (free when no profiler is attached).  With
``FL4H_TRACE=1`` each phase is also bracketed by CUDA events; ``phase_report()`` returns the accumulated device
milliseconds per phase (resolved lazily, so tracing never adds a synchronisation to the round) and the clients / server
attach it to their round reports.
"""

from __future__ import annotations

import contextlib
import os
from collections import defaultdict
from collections.abc import Iterator

import torch

_PENDING: dict[str, list[tuple[torch.cuda.Event, torch.cuda.Event]]] = defaultdict(list)
_TOTALS: dict[str, float] = defaultdict(float)
_COUNTS: dict[str, int] = defaultdict(int)


def tracing_enabled() -> bool:
    return os.environ.get("FL4H_TRACE", "0") == "1" and torch.cuda.is_available()


@contextlib.contextmanager
def phase(name: str) -> Iterator[None]:
    cuda = torch.cuda.is_available()
    timed = tracing_enabled()
    annotate = torch.autograd.profiler.record_function(f"fl4h:{name}") if torch.autograd._profiler_enabled() else None
    if annotate is not None:  # shows up as a user annotation on torch.profiler / kineto timelines
        annotate.__enter__()
    if cuda:
        torch.cuda.nvtx.range_push(f"fl4h:{name}")
    if timed:
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
    try:
        yield
    finally:
        if timed:
            end.record()
            _PENDING[name].append((start, end))
        if cuda:
            torch.cuda.nvtx.range_pop()
        if annotate is not None:
            annotate.__exit__(None, None, None)


def phase_report(reset: bool = False) -> dict[str, dict[str, float]]:
    """``{phase: {"device_ms": total, "count": n, "mean_ms": ...}}`` over everything recorded so far."""
    for name, pairs in list(_PENDING.items()):
        remaining = []
        for start, end in pairs:
            if end.query():
                _TOTALS[name] += start.elapsed_time(end)
                _COUNTS[name] += 1
            else:
                remaining.append((start, end))
        _PENDING[name] = remaining
    report = {name: {"device_ms": _TOTALS[name], "count": float(_COUNTS[name]), "mean_ms": _TOTALS[name] / max(_COUNTS[name], 1)}
              for name in _TOTALS}
    if reset:
        _TOTALS.clear()
        _COUNTS.clear()
    return report
