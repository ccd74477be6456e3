"""Execution of one client step (train or eval) as a replayable unit.

A client supplies *what* a step is (its ``train_step`` / ``val_step`` hooks and the meters they feed); the
``StepExecutor`` decides *how* it runs: batch staging (device placement, channels-last), autocast, and — on CUDA with
``EngineOptions.cuda_graphs`` — capture of the whole unit (hook body, device-side loss accumulation, metric update) into
a CUDA graph that later batches replay.  One graph is kept per ``(variant, input signature)``; ``variant`` is whatever
the client says changes the launched kernels besides tensor shapes (FedRep's head / representation phase, ...).

Nothing here corresponds to reference code: the reference runs every hook eagerly
(``fl4health/clients/basic_client.py:700-760``).
"""

from __future__ import annotations

import contextlib
from collections import OrderedDict
from collections.abc import Callable, Hashable
from typing import Any

import torch

from fl4health_b200.engine.fused_optim import _FlatOptimizer
from fl4health_b200.engine.graph_runner import GraphStepRunner
from fl4health_b200.engine.options import EngineOptions
from fl4health_b200.utils.client import move_data_to_device

StepFn = Callable[[Any, Any], tuple[Any, Any]]


class RunnerCache(OrderedDict):
    """Insertion-ordered ``key -> GraphStepRunner`` with oldest-first eviction.  Captured graphs pin device memory, and
    variants that re-bind tensors every round (MOON's frozen models) would otherwise pile up; keeping the newest few
    lets alternating variants (FedRep phases) stay resident."""

    def __init__(self, capacity: int = 6) -> None:
        super().__init__()
        self.capacity = capacity

    def obtain(self, key: Hashable, build: Callable[[], GraphStepRunner]) -> GraphStepRunner:
        runner = self.get(key)
        if runner is None:
            while len(self) >= self.capacity:
                self.popitem(last=False)
            runner = self[key] = build()
        return runner


class StepExecutor:
    def __init__(self, engine: EngineOptions, device: torch.device, label: str) -> None:
        self.engine, self.device, self.label = engine, device, label
        self.train_runners = RunnerCache()
        self.eval_runners = RunnerCache()
        self.latest_train_runner: GraphStepRunner | None = None

    # -- policy ------------------------------------------------------------------------------------------------
    @property
    def graphs_on(self) -> bool:
        return self.engine.cuda_graphs and self.device.type == "cuda"

    def autocast(self) -> contextlib.AbstractContextManager:
        dtype = self.engine.amp_dtype
        if dtype is not None and (self.device.type == "cuda" or self.engine.master_weights):
            return torch.autocast(device_type=self.device.type, dtype=dtype)
        return contextlib.nullcontext()

    def stage(self, batch_input: Any, batch_target: Any) -> tuple[Any, Any]:
        """Host/device placement of one batch (+ NHWC for image tensors when the engine runs channels-last)."""
        staged_input = move_data_to_device(batch_input, self.device)
        staged_target = move_data_to_device(batch_target, self.device)
        if self.engine.channels_last and isinstance(staged_input, torch.Tensor) and staged_input.dim() == 4:
            staged_input = staged_input.contiguous(memory_format=torch.channels_last)
        return staged_input, staged_target

    def reset(self) -> None:
        """Forget every captured graph (after anything re-binds tensors a step reads: new model, new optimizer)."""
        self.train_runners.clear()
        self.eval_runners.clear()
        self.latest_train_runner = None

    # -- execution ---------------------------------------------------------------------------------------------
    @staticmethod
    def push_hyperparameters(optimizers: dict[str, Any]) -> None:
        """lr / mu edits made on the host reach the optimizers' device-side blocks (read by replayed kernels)."""
        for optimizer in optimizers.values():
            if isinstance(optimizer, _FlatOptimizer):
                optimizer.sync_hyperparams()

    def run_train(self, unit: StepFn, variant: Hashable, before_replay: Callable[[], None], batch_input: Any,
                  batch_target: Any) -> tuple[Any, Any]:
        if not self.graphs_on:
            return unit(batch_input, batch_target)
        runner = self.train_runners.obtain(variant, lambda: GraphStepRunner(
            unit, self.device, warmup=self.engine.graph_warmup_steps, name=f"{self.label}/train[{variant}]",
            before_replay=before_replay))
        self.latest_train_runner = runner
        return runner(batch_input, batch_target)

    def run_eval(self, unit: StepFn, key: Hashable, batch_input: Any, batch_target: Any) -> tuple[Any, Any]:
        if not self.graphs_on:
            return unit(batch_input, batch_target)
        runner = self.eval_runners.obtain(key, lambda: GraphStepRunner(
            unit, self.device, warmup=self.engine.graph_warmup_steps, name=f"{self.label}/eval"))
        return runner(batch_input, batch_target)
