"""CUDA-graph capture of a client's per-batch step.

The reference's hot loop (``basic_client.py:726-758``) launches ~10^2-10^3 tiny kernels per batch from Python and
syncs the device every step to read the loss.  On a B200 a ResNet-18/CIFAR batch-32 step is ~50 µs of math, so the
step is entirely launch-bound.  ``GraphStepRunner`` runs the *same Python hooks* (``train_step`` and friends — any
subclass override included) under ``torch.cuda.graph`` once, after a few eager warm-up steps, and from then on a step
is: two async copies into static input buffers + one ``cudaGraphLaunch``.

Rules of the road:
* eager for the first ``warmup`` steps of each input signature (shapes/dtypes), so lazily-created state (optimizer
  moments, loss-meter accumulators, metric counters) exists before capture;
* everything the step reads that changes between replays must live in device memory at a fixed address (inputs are
  copied into static buffers; hyper-parameters go through the fused optimizer's device block);
* if capture fails (a hook syncs, allocates pinned memory, calls ``.item()`` ...) the runner permanently falls back
  to eager execution for that signature and logs why — correctness never depends on capture.
"""

from __future__ import annotations

from collections.abc import Callable
from dataclasses import dataclass, field
from logging import INFO, WARNING
from typing import Any

import torch

from fl4health_b200.common.logger import log

StepFn = Callable[[Any, Any], Any]


def _signature(obj: Any) -> tuple:
    if isinstance(obj, torch.Tensor):
        return ("t", tuple(obj.shape), obj.dtype)
    if isinstance(obj, dict):
        return ("d",) + tuple((k, _signature(v)) for k, v in sorted(obj.items()))
    if obj is None:
        return ("n",)
    raise TypeError(f"unsupported step input type {type(obj)}")


def _alloc_like(obj: Any, device: torch.device) -> Any:
    if isinstance(obj, torch.Tensor):
        # same physical layout as the live batch (channels-last images stay channels-last: layout-specialised kernels
        # — the NHWC stem convolution — must see during capture what they saw during the eager warm-up)
        if obj.dim() == 4 and not obj.is_contiguous() and obj.is_contiguous(memory_format=torch.channels_last):
            return torch.empty(obj.shape, dtype=obj.dtype, device=device, memory_format=torch.channels_last)
        return torch.empty(obj.shape, dtype=obj.dtype, device=device)
    if isinstance(obj, dict):
        return {k: _alloc_like(v, device) for k, v in obj.items()}
    return None


def _copy_into(dst: Any, src: Any) -> None:
    if isinstance(dst, torch.Tensor):
        dst.copy_(src, non_blocking=True)
    elif isinstance(dst, dict):
        for key, value in dst.items():
            _copy_into(value, src[key])


_STEP_STREAMS: dict[int, torch.cuda.Stream] = {}


def step_stream(device: torch.device) -> torch.cuda.Stream:
    """The ONE stream per device on which steps are warmed up and captured.  Autograd's gradient accumulators remember
    the stream they were created on; if the eager warm-up ran on the default stream and the capture on torch's private
    capture stream, every accumulator kept alive across iterations (a client holding on to a loss tensor is enough)
    makes the captured backward copy each parameter gradient instead of adopting it -- one extra elementwise kernel per
    parameter per step.  Same stream for both, shared by every runner of the device: no mismatch."""
    index = device.index if device.index is not None else torch.cuda.current_device()
    if index not in _STEP_STREAMS:
        _STEP_STREAMS[index] = torch.cuda.Stream(device=device)
    return _STEP_STREAMS[index]


def _record_on(obj: Any, stream: torch.cuda.Stream) -> None:
    """Tell the caching allocator that tensors produced on the step stream are consumed on ``stream``."""
    if isinstance(obj, torch.Tensor):
        if obj.is_cuda:
            obj.record_stream(stream)
    elif isinstance(obj, dict):
        for value in obj.values():
            _record_on(value, stream)
    elif isinstance(obj, (list, tuple)):
        for value in obj:
            _record_on(value, stream)
    elif hasattr(obj, "__dict__") and not isinstance(obj, torch.nn.Module):
        for value in vars(obj).values():
            if isinstance(value, (torch.Tensor, dict, list, tuple)):
                _record_on(value, stream)


@dataclass
class _Captured:
    graph: torch.cuda.CUDAGraph
    static_input: Any
    static_target: Any
    outputs: Any
    kernel_launches: int = 0


@dataclass
class GraphStepRunner:
    """Runs ``fn(input, target)`` eagerly ``warmup`` times per input signature, then through a captured graph."""

    fn: StepFn
    device: torch.device
    warmup: int = 3
    name: str = "step"
    before_replay: Callable[[], None] | None = None  # e.g. push changed LR into the device hyper-parameter block
    max_signatures: int = 64  # captured graphs pin memory: past this many input signatures, new ones run eagerly
    _seen: dict[tuple, int] = field(default_factory=dict)
    _graphs: dict[tuple, _Captured] = field(default_factory=dict)
    _disabled: set = field(default_factory=set)
    replays: int = 0
    eager_steps: int = 0

    def __call__(self, input: Any, target: Any) -> Any:
        if self.device.type != "cuda":
            self.eager_steps += 1
            return self.fn(input, target)
        sig = (_signature(input), _signature(target))
        captured = self._graphs.get(sig)
        if captured is not None:
            _copy_into(captured.static_input, input)
            _copy_into(captured.static_target, target)
            if self.before_replay is not None:
                self.before_replay()
            captured.graph.replay()
            self.replays += 1
            if captured.kernel_launches:
                from fl4health_b200.ops import _lib

                _lib.count_launches(captured.kernel_launches)
            return captured.outputs
        count = self._seen.get(sig, 0)
        # lazily created state (optimizer moments, meters) exists once ANY signature has been captured: a further
        # signature (Poisson-sampled batch sizes, a ragged last batch) needs one eager pass, for library autotuning
        needed = self.warmup if not self._graphs else min(self.warmup, 1)
        if sig not in self._disabled and count >= needed and len(self._graphs) >= self.max_signatures:
            self._disabled.add(sig)
        if sig in self._disabled or count < needed:
            self._seen[sig] = count + 1
            self.eager_steps += 1
            return self._eager_on_step_stream(input, target)
        return self._capture_and_run(sig, input, target)

    def _eager_on_step_stream(self, input: Any, target: Any) -> Any:
        caller, stream = torch.cuda.current_stream(self.device), step_stream(self.device)
        stream.wait_stream(caller)
        with torch.cuda.stream(stream):
            outputs = self.fn(input, target)
        caller.wait_stream(stream)
        _record_on(outputs, caller)
        return outputs

    def _capture_and_run(self, sig: tuple, input: Any, target: Any) -> Any:
        from fl4health_b200.ops import _lib

        static_input = _alloc_like(input, self.device)
        static_target = _alloc_like(target, self.device)
        _copy_into(static_input, input)
        _copy_into(static_target, target)
        if self.before_replay is not None:
            self.before_replay()
        graph = torch.cuda.CUDAGraph()
        torch.cuda.synchronize(self.device)
        launches_before = _lib.launch_count()
        try:
            with torch.cuda.graph(graph, stream=step_stream(self.device)):
                outputs = self.fn(static_input, static_target)
        except Exception as exc:  # noqa: BLE001 - any capture failure means "run eagerly"
            torch.cuda.synchronize(self.device)
            self._disabled.add(sig)
            log(WARNING, f"[{self.name}] CUDA-graph capture failed ({type(exc).__name__}: {exc}); running eagerly.")
            self.eager_steps += 1
            return self._eager_on_step_stream(input, target)
        launched = _lib.launch_count() - launches_before  # our kernels recorded into the graph (not executed yet)
        _lib.count_launches(-launched)
        captured = _Captured(graph, static_input, static_target, outputs, launched)
        self._graphs[sig] = captured
        log(INFO, f"[{self.name}] captured CUDA graph for signature {hash(sig) & 0xFFFF:04x} ({launched} fl4h kernels)")
        graph.replay()  # capture does not execute: run the step for real
        self.replays += 1
        _lib.count_launches(launched)
        return outputs

    def reset(self) -> None:
        self._graphs.clear()
        self._seen.clear()
        self._disabled.clear()
