"""``BasicClient``: the template-method FL client every algorithm client derives from.

API parity: ``fl4health/clients/basic_client.py:43-1321`` — same constructor, the same ~15 overridable hooks
(``get_model``, ``get_optimizer``, ``get_data_loaders``, ``get_criterion``, ``predict``,
``compute_loss_and_additional_losses``, ``transform_gradients``, ``update_before_train`` ...), the same
``fit``/``evaluate``/``get_parameters``/``get_properties`` protocol, report keys and config keys.

What is different is *how the hooks are executed* (``EngineOptions``):

* the model's state is re-homed into a flat ``ParameterArena`` so exchange/optimizer/penalties are single kernels;
* stock ``torch.optim`` optimizers are swapped for one-launch flat equivalents;
* the per-batch step (hooks included) can be captured in a CUDA graph and replayed;
* losses and metrics accumulate on the device and are read back once per round, not once per step
  (reference: ``losses.as_dict()`` → ``.item()`` every step, ``basic_client.py:751``).
"""

from __future__ import annotations

import contextlib
import datetime
from collections.abc import Iterator, Sequence
from logging import INFO, WARNING
from pathlib import Path
from typing import Any

import torch
from torch import nn
from torch.nn.modules.loss import _Loss
from torch.optim import Optimizer
from torch.optim.lr_scheduler import LRScheduler
from torch.utils.data import DataLoader

from fl4health_b200.checkpointing.client_module import CheckpointMode, ClientCheckpointAndStateModule
from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import Config, NDArrays, Scalar
from fl4health_b200.engine.fused_optim import _FlatOptimizer, translate_optimizer
from fl4health_b200.engine.graph_runner import GraphStepRunner
from fl4health_b200.engine.modes import set_training
from fl4health_b200.engine.options import EngineOptions
from fl4health_b200.metrics.base_metrics import TEST_LOSS_KEY, TEST_NUM_EXAMPLES_KEY, Metric
from fl4health_b200.metrics.metric_managers import MetricManager
from fl4health_b200.parallel.arena import ParameterArena, arena_of, attach_arena
from fl4health_b200.utils import tracing
from fl4health_b200.parameter_exchange.full_exchanger import FullParameterExchanger
from fl4health_b200.parameter_exchange.parameter_exchanger_base import ParameterExchanger
from fl4health_b200.reporting.base_reporter import BaseReporter
from fl4health_b200.reporting.reports_manager import ReportsManager
from fl4health_b200.utils.client import (
    check_if_batch_is_empty_and_verify_input,
    fold_loss_dict_into_metrics,
    maybe_progress_bar,
    move_data_to_device,
    process_and_check_validation_steps,
    set_pack_losses_with_val_metrics,
)
from fl4health_b200.utils.config import narrow_dict_type
from fl4health_b200.utils.early_stopper import EarlyStopper
from fl4health_b200.utils.logging import LoggingMode
from fl4health_b200.utils.losses import EvaluationLosses, LossMeter, LossMeterType, TrainingLosses
from fl4health_b200.utils.random import generate_hash
from fl4health_b200.utils.typing import LogLevel, TorchFeatureType, TorchInputType, TorchPredType, TorchTargetType

EXPECTED_OUTPUT_TUPLE_SIZE = 2  # a model may return (predictions, features)


class BasicClient:
    def __init__(
        self,
        data_path: Path,
        metrics: Sequence[Metric],
        device: torch.device,
        loss_meter_type: LossMeterType = LossMeterType.AVERAGE,
        checkpoint_and_state_module: ClientCheckpointAndStateModule | None = None,
        reporters: Sequence[BaseReporter] | None = None,
        progress_bar: bool = False,
        client_name: str | None = None,
        engine_options: EngineOptions | None = None,
    ) -> None:
        self.data_path = data_path
        self.device = torch.device(device)
        self.metrics = metrics
        self.progress_bar = progress_bar
        self.client_name = client_name if client_name is not None else generate_hash()
        log(INFO, f"Client Name: {self.client_name}")
        self.engine = engine_options if engine_options is not None else EngineOptions.from_env()

        self.checkpoint_and_state_module = checkpoint_and_state_module or ClientCheckpointAndStateModule(
            pre_aggregation=None, post_aggregation=None, state_checkpointer=None
        )
        self.reports_manager = ReportsManager(reporters)
        self.reports_manager.initialize(id=self.client_name, name=self.client_name)

        self.initialized = False

        self.train_loss_meter = LossMeter[TrainingLosses](loss_meter_type, TrainingLosses)
        self.val_loss_meter = LossMeter[EvaluationLosses](loss_meter_type, EvaluationLosses)
        self.test_loss_meter = LossMeter[EvaluationLosses](loss_meter_type, EvaluationLosses)
        self.train_metric_manager = MetricManager(metrics=self.metrics, metric_manager_name="train")
        self.val_metric_manager = MetricManager(metrics=self.metrics, metric_manager_name="val")
        self.test_metric_manager = MetricManager(metrics=self.metrics, metric_manager_name="test")

        self.initial_weights: NDArrays | None = None
        self.total_steps: int = 0
        self.total_epochs: int = 0

        # set in setup_client
        self.parameter_exchanger: ParameterExchanger
        self.model: nn.Module
        self.optimizers: dict[str, Optimizer]
        self.lr_schedulers: dict[str, LRScheduler]
        self.criterion: _Loss
        self.train_loader: DataLoader
        self.val_loader: DataLoader
        self.test_loader: DataLoader | None
        self.num_train_samples: int
        self.num_val_samples: int
        self.num_test_samples: int | None = None
        self.learning_rate: float | None = None

        self.early_stopper: EarlyStopper | None = None
        self.num_validation_steps: int | None = None
        self.train_iterator: Iterator | None = None
        self.val_iterator: Iterator | None = None

        # engine state
        self._train_runner: GraphStepRunner | None = None
        self._train_runners: dict[Any, GraphStepRunner] = {}
        self._val_runners: dict[Any, GraphStepRunner] = {}

    # ==================================================================================================================
    # protocol: parameters in / out
    # ==================================================================================================================
    def _maybe_checkpoint(self, loss: float, metrics: dict[str, Scalar], checkpoint_mode: CheckpointMode) -> None:
        self.checkpoint_and_state_module.maybe_checkpoint(self.model, loss, metrics, checkpoint_mode)

    def get_parameters(self, config: Config) -> NDArrays:
        """Arrays for the server.  Before the client is set up, this sets it up and returns ALL model state (the
        server uses it to initialise the global model)."""
        if not self.initialized:
            return self.setup_client_and_return_all_model_parameters(config)
        assert self.model is not None and self.parameter_exchanger is not None
        self._maybe_load_saved_best_local_model_state()
        return self.parameter_exchanger.push_parameters(self.model, config=config)

    def _maybe_load_saved_best_local_model_state(self) -> None:
        if self.early_stopper is not None and self.early_stopper.patience is None:
            log(INFO, "Loading saved best model's state before sending model to server.")
            self.early_stopper.load_snapshot(["model"])

    def set_parameters(self, parameters: NDArrays, config: Config, fitting_round: bool) -> None:
        """Round 1 of fitting initialises *all* weights (full exchange) whatever the exchanger; afterwards the
        client's exchanger decides which state is overwritten."""
        assert self.model is not None
        current_server_round = narrow_dict_type(config, "current_server_round", int)
        if current_server_round == 1 and fitting_round:
            self.initialize_all_model_weights(parameters, config)
        else:
            assert self.parameter_exchanger is not None
            self.parameter_exchanger.pull_parameters(parameters, self.model, config)

    def initialize_all_model_weights(self, parameters: NDArrays, config: Config) -> None:
        FullParameterExchanger().pull_parameters(parameters, self.model, config)

    def setup_client_and_return_all_model_parameters(self, config: Config) -> NDArrays:
        log(INFO, "Setting up client and providing full model parameters to the server for initialization")
        if not config:
            log(
                WARNING,
                "This client has not yet been initialized and the config is empty. This may cause unexpected "
                "failures, as setting up a client typically requires several configuration parameters, "
                "including batch_size and current_server_round.",
            )
        self.setup_client(config)
        return FullParameterExchanger().push_parameters(self.model, config=config)

    def initial_parameters_requested(self, config: Config) -> bool:
        """True when ``get_parameters`` is the server's request for INITIAL parameters (round 0 of
        ``on_init_parameters_config_fn``) reaching a client that has already been set up — a properties poll (nnU-Net plan
        negotiation, tabular feature alignment) may have initialised it.  Clients that pack side information into their
        regular payload must answer this request with the plain model state, exactly like an uninitialised client."""
        return self.initialized and config.get("current_server_round") == 0

    def shutdown(self) -> None:
        self.reports_manager.report({"shutdown": str(datetime.datetime.now())})
        self.reports_manager.shutdown()

    # ==================================================================================================================
    # protocol: fit / evaluate
    # ==================================================================================================================
    def process_config(self, config: Config) -> tuple[int | None, int | None, int, bool, bool]:
        current_server_round = narrow_dict_type(config, "current_server_round", int)
        if ("local_epochs" in config) and ("local_steps" in config):
            raise ValueError("Config cannot contain both local_epochs and local_steps. Please specify only one.")
        if "local_epochs" in config:
            local_epochs: int | None = narrow_dict_type(config, "local_epochs", int)
            local_steps: int | None = None
        elif "local_steps" in config:
            local_steps = narrow_dict_type(config, "local_steps", int)
            local_epochs = None
        else:
            raise ValueError("Must specify either local_epochs or local_steps in the Config.")
        evaluate_after_fit = bool(config.get("evaluate_after_fit", False))
        pack_losses_with_val_metrics = set_pack_losses_with_val_metrics(config)
        return local_epochs, local_steps, current_server_round, evaluate_after_fit, pack_losses_with_val_metrics

    def fit(self, parameters: NDArrays, config: Config) -> tuple[NDArrays, int, dict[str, Scalar]]:
        round_start_time = datetime.datetime.now()
        local_epochs, local_steps, current_server_round, evaluate_after_fit, pack_losses_with_val_metrics = (
            self.process_config(config)
        )
        if not self.initialized:
            self.setup_client(config)
            if self.checkpoint_and_state_module.state_checkpointer is not None:
                loaded = self._load_client_state()
                log(INFO, "Successfully loaded client state." if loaded else "Client state was not loaded.")

        with tracing.phase("pull_parameters"):
            self.set_parameters(parameters, config, fitting_round=True)
        with tracing.phase("update_before_train"):
            self.update_before_train(current_server_round)

        fit_start_time = datetime.datetime.now()
        with tracing.phase("local_train"):
            if local_epochs is not None:
                loss_dict, metrics = self.train_by_epochs(local_epochs, current_server_round)
                local_steps = len(self.train_loader) * local_epochs
            elif local_steps is not None:
                loss_dict, metrics = self.train_by_steps(local_steps, current_server_round)
            else:
                raise ValueError("Must specify either local_epochs or local_steps in the Config.")
        fit_end_time = datetime.datetime.now()

        with tracing.phase("update_after_train"):
            self.update_after_train(local_steps, loss_dict, config)

        if self._should_evaluate_after_fit(evaluate_after_fit):
            validation_loss, validation_metrics = self.validate(pack_losses_with_val_metrics)
            metrics.update(validation_metrics)
            self._maybe_checkpoint(validation_loss, validation_metrics, CheckpointMode.PRE_AGGREGATION)

        self.reports_manager.report(
            {
                "fit_round_metrics": metrics,
                "fit_round_losses": loss_dict,
                "round": current_server_round,
                "round_start": str(round_start_time),
                "round_end": str(datetime.datetime.now()),
                "fit_round_start": str(fit_start_time),
                "fit_round_time_elapsed": round((fit_end_time - fit_start_time).total_seconds()),
                "fit_round_end": str(fit_end_time),
                "fit_step": self.total_steps,
                "fit_epoch": self.total_epochs,
                **({"device_phase_ms": tracing.phase_report()} if tracing.tracing_enabled() else {}),
            },
            current_server_round,
        )

        if self.checkpoint_and_state_module.state_checkpointer is not None:
            self._save_client_state()

        with tracing.phase("push_parameters"):
            return self.get_parameters(config), self.num_train_samples, metrics

    def evaluate(self, parameters: NDArrays, config: Config) -> tuple[float, int, dict[str, Scalar]]:
        if not self.initialized:
            self.setup_client(config)
        start_time = datetime.datetime.now()
        current_server_round = narrow_dict_type(config, "current_server_round", int)
        pack_losses_with_val_metrics = set_pack_losses_with_val_metrics(config)

        with tracing.phase("pull_parameters"):
            self.set_parameters(parameters, config, fitting_round=False)
        with tracing.phase("evaluate"):
            loss, metrics = self.validate(pack_losses_with_val_metrics)
        end_time = datetime.datetime.now()

        self._maybe_checkpoint(loss, metrics, CheckpointMode.POST_AGGREGATION)
        self.reports_manager.report(
            {
                "eval_round_metrics": metrics,
                "eval_round_loss": loss,
                "eval_round_start": str(start_time),
                "eval_round_time_elapsed": round((end_time - start_time).total_seconds()),
                "eval_round_end": str(end_time),
                "fit_step": self.total_steps,
                "fit_epoch": self.total_epochs,
                "round": current_server_round,
            },
            current_server_round,
        )
        return loss, self.num_val_samples, metrics

    def _should_evaluate_after_fit(self, evaluate_after_fit: bool) -> bool:
        pre_aggregation = self.checkpoint_and_state_module.pre_aggregation
        return evaluate_after_fit or pre_aggregation is not None

    # ==================================================================================================================
    # logging / reporting helpers
    # ==================================================================================================================
    def _log_header_str(
        self, current_round: int | None = None, current_epoch: int | None = None,
        logging_mode: LoggingMode = LoggingMode.TRAIN,
    ) -> None:
        parts = [f"Current FL Round: {current_round}"] if current_round is not None else []
        if current_epoch is not None:
            parts.append(f"Current Epoch: {current_epoch}")
        log(INFO, f"{logging_mode.value} | " + "\t".join(parts))

    def _log_results(
        self, loss_dict: dict[str, float], metrics_dict: dict[str, Scalar], current_round: int | None = None,
        current_epoch: int | None = None, logging_mode: LoggingMode = LoggingMode.TRAIN,
    ) -> None:
        _, client_logs = self.get_client_specific_logs(current_round, current_epoch, logging_mode)
        lines = [f"Client {logging_mode.value} Losses: " + ", ".join(f"{k}: {v:.6f}" for k, v in loss_dict.items())]
        if metrics_dict:
            lines.append(
                f"Client {logging_mode.value} Metrics: " + ", ".join(f"{k}: {v}" for k, v in metrics_dict.items())
            )
        log(INFO, " | ".join(lines))
        for level, message in client_logs:
            log(level.value, message)

    def get_client_specific_logs(
        self, current_round: int | None, current_epoch: int | None, logging_mode: LoggingMode
    ) -> tuple[str, list[tuple[LogLevel, str]]]:
        """Hook: extra header text + log lines (e.g. current FedProx mu)."""
        return "", []

    def get_client_specific_reports(self) -> dict[str, Any]:
        """Hook: extra key/values merged into reporter payloads."""
        return {}

    def _step_reports_enabled(self) -> bool:
        if self.engine.step_reports is not None:
            return self.engine.step_reports
        return any(getattr(r, "wants_step_reports", False) for r in self.reports_manager.reporters)

    def update_metric_manager(self, preds: TorchPredType, target: TorchTargetType, metric_manager: MetricManager) -> None:
        metric_manager.update(preds, target)

    # ==================================================================================================================
    # the per-batch step
    # ==================================================================================================================
    def _amp(self) -> contextlib.AbstractContextManager:
        if self.engine.amp_dtype is not None and (self.device.type == "cuda" or self.engine.master_weights):
            return torch.autocast(device_type=self.device.type, dtype=self.engine.amp_dtype)
        return contextlib.nullcontext()

    def train_step(self, input: TorchInputType, target: TorchTargetType) -> tuple[TrainingLosses, TorchPredType]:
        self.optimizers["global"].zero_grad()
        with self._amp():
            preds, features = self.predict(input)
            target = self.transform_target(target)
            losses = self.compute_training_loss(preds, features, target)
        losses.backward["backward"].backward()
        self.transform_gradients(losses)
        self.optimizers["global"].step()
        return losses, preds

    def val_step(self, input: TorchInputType, target: TorchTargetType) -> tuple[EvaluationLosses, TorchPredType]:
        with torch.no_grad(), self._amp():
            preds, features = self.predict(input)
            target = self.transform_target(target)
            losses = self.compute_evaluation_loss(preds, features, target)
        return losses, preds

    # --- engine wrappers: (train|val)_step + device-side loss/metric accumulation as one replayable unit ---------
    def _train_unit(self, input: TorchInputType, target: TorchTargetType) -> tuple[TrainingLosses, TorchPredType]:
        losses, preds = self.train_step(input, target)
        self.train_loss_meter.accumulate(losses)
        self.update_metric_manager(preds, target, self.train_metric_manager)
        return losses.detach(), {key: value.detach() for key, value in preds.items()}  # type: ignore[return-value]

    def _sync_optimizer_hyperparams(self) -> None:
        for optimizer in self.optimizers.values():
            if isinstance(optimizer, _FlatOptimizer):
                optimizer.sync_hyperparams()

    def _run_train_unit(self, input: TorchInputType, target: TorchTargetType) -> tuple[TrainingLosses, TorchPredType]:
        if self.engine.cuda_graphs and self.device.type == "cuda":
            variant = self._graph_variant()
            runner = self._train_runners.get(variant)
            if runner is None:
                self._evict_stale_runners(self._train_runners)
                runner = GraphStepRunner(
                    self._train_unit, self.device, warmup=self.engine.graph_warmup_steps,
                    name=f"{self.client_name}/train[{variant}]", before_replay=self._sync_optimizer_hyperparams,
                )
                self._train_runners[variant] = runner
                self._train_runner = runner
            losses, preds = runner(input, target)
        else:
            losses, preds = self._train_unit(input, target)
        self.train_loss_meter.mark_step()
        return losses, preds

    def _run_val_unit(
        self, input: TorchInputType, target: TorchTargetType, loss_meter: LossMeter, metric_manager: MetricManager
    ) -> tuple[EvaluationLosses, TorchPredType]:
        def unit(inp: TorchInputType, tgt: TorchTargetType) -> tuple[EvaluationLosses, TorchPredType]:
            losses, preds = self.val_step(inp, tgt)
            loss_meter.accumulate(losses)
            self.update_metric_manager(preds, tgt, metric_manager)
            return losses, preds

        if self.engine.cuda_graphs and self.device.type == "cuda":
            key = (id(loss_meter), self._graph_variant())
            runner = self._val_runners.get(key)
            if runner is None:
                self._evict_stale_runners(self._val_runners)
                runner = GraphStepRunner(unit, self.device, warmup=self.engine.graph_warmup_steps,
                                         name=f"{self.client_name}/eval")
                self._val_runners[key] = runner
            losses, preds = runner(input, target)
        else:
            losses, preds = unit(input, target)
        loss_meter.mark_step()
        return losses, preds

    @staticmethod
    def _evict_stale_runners(runners: dict, keep: int = 6) -> None:
        """Captured graphs pin device memory; variants that re-bind tensors every round (MOON's frozen models) would
        otherwise accumulate.  Oldest-first eviction keeps alternating variants (FedRep phases) resident."""
        while len(runners) >= keep:
            runners.pop(next(iter(runners)))

    def _graph_variant(self) -> Any:
        """Hashable tag of everything *besides input shapes* that changes what ``train_step`` launches (e.g. FedRep's
        head/representation phase).  One captured graph is kept per (variant, input signature)."""
        return "default"

    def _invalidate_graphs(self) -> None:
        """Drop captured graphs (call after anything that re-binds tensors the step reads: new model, new optimizer)."""
        self._train_runner = None
        self._train_runners = {}
        self._val_runners = {}

    def _prepare_batch(self, input: TorchInputType, target: TorchTargetType) -> tuple[TorchInputType, TorchTargetType]:
        input = move_data_to_device(input, self.device)
        target = move_data_to_device(target, self.device)
        if self.engine.channels_last and isinstance(input, torch.Tensor) and input.dim() == 4:
            input = input.contiguous(memory_format=torch.channels_last)
        return input, target

    # ==================================================================================================================
    # training loops
    # ==================================================================================================================
    def train_by_epochs(
        self, epochs: int, current_round: int | None = None
    ) -> tuple[dict[str, float], dict[str, Scalar]]:
        set_training(self.model, True)
        steps_this_round = 0
        report_data: dict[str, Any] = {"round": current_round}
        step_reports = self._step_reports_enabled()
        continue_training = True
        loss_dict: dict[str, float] = {}
        metrics: dict[str, Scalar] = {}
        for local_epoch in range(epochs):
            self.train_metric_manager.clear()
            self.train_loss_meter.clear()
            self._log_header_str(current_round, local_epoch)
            self.update_before_epoch(epoch=local_epoch)
            report_data.update({"fit_epoch": self.total_epochs})
            for input, target in maybe_progress_bar(self.train_loader, self.progress_bar):
                self.update_before_step(steps_this_round, current_round)
                if check_if_batch_is_empty_and_verify_input(input):
                    log(INFO, "Empty batch generated by data loader. Skipping step.")
                    continue
                input, target = self._prepare_batch(input, target)
                losses, _ = self._run_train_unit(input, target)
                self.update_after_step(steps_this_round, current_round)
                self.update_lr_schedulers(epoch=local_epoch)
                if step_reports:
                    report_data.update({"fit_step_losses": losses.as_dict(), "fit_step": self.total_steps})
                    report_data.update(self.get_client_specific_reports())
                    self.reports_manager.report(report_data, current_round, self.total_epochs, self.total_steps)
                self.total_steps += 1
                steps_this_round += 1
                if self.early_stopper is not None and self.early_stopper.should_stop(steps_this_round):
                    log(INFO, "Early stopping criterion met. Stopping training.")
                    self.early_stopper.load_snapshot()
                    continue_training = False
                    break
            metrics = self.train_metric_manager.compute()
            loss_dict = self.train_loss_meter.compute().as_dict()
            report_data.update({"fit_epoch_metrics": metrics, "fit_epoch_losses": loss_dict})
            report_data.update(self.get_client_specific_reports())
            self.reports_manager.report(report_data, current_round, self.total_epochs)
            self._log_results(loss_dict, metrics, current_round, local_epoch)
            self.total_epochs += 1
            if not continue_training:
                break
        return loss_dict, metrics

    def _next_train_batch(self) -> tuple[TorchInputType, TorchTargetType]:
        if self.train_iterator is None:
            self.train_iterator = iter(self.train_loader)
        try:
            return next(self.train_iterator)
        except StopIteration:
            self.train_iterator = iter(self.train_loader)
            return next(self.train_iterator)

    def train_by_steps(
        self, steps: int, current_round: int | None = None
    ) -> tuple[dict[str, float], dict[str, Scalar]]:
        set_training(self.model, True)
        self.train_loss_meter.clear()
        self.train_metric_manager.clear()
        self._log_header_str(current_round)
        report_data: dict[str, Any] = {"round": current_round}
        step_reports = self._step_reports_enabled()
        for step in maybe_progress_bar(range(steps), self.progress_bar):
            self.update_before_step(step, current_round)
            input, target = self._next_train_batch()
            if check_if_batch_is_empty_and_verify_input(input):
                log(INFO, "Empty batch generated by data loader. Skipping step.")
                continue
            input, target = self._prepare_batch(input, target)
            losses, _ = self._run_train_unit(input, target)
            self.update_after_step(step, current_round)
            self.update_lr_schedulers(step=step)
            if step_reports:
                report_data.update({"fit_step_losses": losses.as_dict(), "fit_step": self.total_steps})
                report_data.update(self.get_client_specific_reports())
                self.reports_manager.report(report_data, current_round, None, self.total_steps)
            self.total_steps += 1
            if self.early_stopper is not None and self.early_stopper.should_stop(step):
                log(INFO, "Early stopping criterion met. Stopping training.")
                self.early_stopper.load_snapshot()
                break
        loss_dict = self.train_loss_meter.compute().as_dict()
        metrics = self.train_metric_manager.compute()
        self._log_results(loss_dict, metrics, current_round)
        return loss_dict, metrics

    # ==================================================================================================================
    # validation
    # ==================================================================================================================
    def _finish_validation(
        self, loss_meter: LossMeter, metric_manager: MetricManager, logging_mode: LoggingMode,
        include_losses_in_metrics: bool,
    ) -> tuple[float, dict[str, Scalar]]:
        loss_dict = loss_meter.compute().as_dict()
        metrics = metric_manager.compute()
        self._log_results(loss_dict, metrics, logging_mode=logging_mode)
        if include_losses_in_metrics:
            fold_loss_dict_into_metrics(metrics, loss_dict, logging_mode)
        return loss_dict["checkpoint"], metrics

    def _validate_by_steps(
        self, loss_meter: LossMeter, metric_manager: MetricManager, include_losses_in_metrics: bool = False
    ) -> tuple[float, dict[str, Scalar]]:
        assert self.num_validation_steps is not None, "num_validation_steps must be defined to use this function"
        set_training(self.model, False)
        metric_manager.clear()
        loss_meter.clear()
        if self.val_iterator is None:
            self.val_iterator = iter(self.val_loader)
        with torch.no_grad():
            for _ in maybe_progress_bar(range(self.num_validation_steps), self.progress_bar):
                try:
                    input, target = next(self.val_iterator)
                except StopIteration:
                    self.val_iterator = iter(self.val_loader)
                    input, target = next(self.val_iterator)
                input, target = self._prepare_batch(input, target)
                self._run_val_unit(input, target, loss_meter, metric_manager)
        return self._finish_validation(loss_meter, metric_manager, LoggingMode.VALIDATION, include_losses_in_metrics)

    def _fully_validate_or_test(
        self, loader: DataLoader, loss_meter: LossMeter, metric_manager: MetricManager,
        logging_mode: LoggingMode = LoggingMode.VALIDATION, include_losses_in_metrics: bool = False,
    ) -> tuple[float, dict[str, Scalar]]:
        assert logging_mode in (LoggingMode.VALIDATION, LoggingMode.TEST, LoggingMode.EARLY_STOP_VALIDATION)
        set_training(self.model, False)
        metric_manager.clear()
        loss_meter.clear()
        with torch.no_grad():
            for input, target in maybe_progress_bar(loader, self.progress_bar):
                input, target = self._prepare_batch(input, target)
                self._run_val_unit(input, target, loss_meter, metric_manager)
        return self._finish_validation(loss_meter, metric_manager, logging_mode, include_losses_in_metrics)

    def validate(self, include_losses_in_metrics: bool = False) -> tuple[float, dict[str, Scalar]]:
        if self.num_validation_steps is None:
            val_loss, val_metrics = self._fully_validate_or_test(
                self.val_loader, self.val_loss_meter, self.val_metric_manager,
                include_losses_in_metrics=include_losses_in_metrics,
            )
        else:
            val_loss, val_metrics = self._validate_by_steps(
                self.val_loss_meter, self.val_metric_manager, include_losses_in_metrics=include_losses_in_metrics
            )
        if self.test_loader:
            test_loss, test_metrics = self._fully_validate_or_test(
                self.test_loader, self.test_loss_meter, self.test_metric_manager, LoggingMode.TEST,
                include_losses_in_metrics=include_losses_in_metrics,
            )
            if self.num_test_samples is not None:
                val_metrics[TEST_NUM_EXAMPLES_KEY] = self.num_test_samples
            val_metrics[TEST_LOSS_KEY] = test_loss
            val_metrics.update(test_metrics)
        return val_loss, val_metrics

    def get_properties(self, config: Config) -> dict[str, Scalar]:
        if not self.initialized:
            self.setup_client(config)
        return {"num_train_samples": self.num_train_samples, "num_val_samples": self.num_val_samples}

    # ==================================================================================================================
    # setup
    # ==================================================================================================================
    def _place_model(self, model: nn.Module, with_grad: bool = True) -> nn.Module:
        """Move a model to the device and (engine option) re-home its state into a flat arena."""
        model = model.to(self.device)
        if self.engine.arena:
            arena = attach_arena(model, self.device, with_grad=with_grad, channels_last=self.engine.channels_last,
                                 allocator=self._arena_allocator())
            if self.engine.master_weights and self.engine.amp_dtype is not None and self.engine.fused_optimizer and with_grad:
                arena.enable_compute_shadow(self.engine.amp_dtype)
        elif self.engine.channels_last:
            model = model.to(memory_format=torch.channels_last)
        return model

    def _arena_allocator(self) -> Any:
        """Hook for the SPMD runtime: allocate arenas from peer-mapped symmetric memory."""
        return getattr(self, "arena_allocator", None)

    def _maybe_fuse_optimizers(self) -> None:
        if not self.engine.fused_optimizer:
            return
        for key, optimizer in list(self.optimizers.items()):
            arena = self._arena_for_optimizer(optimizer)
            if arena is not None:
                self.optimizers[key] = translate_optimizer(optimizer, arena)

    def _candidate_modules(self) -> list[nn.Module]:
        """Modules whose arenas an optimizer may be operating on (clients with extra models extend this)."""
        return [self.model]

    def _arena_for_optimizer(self, optimizer: Optimizer) -> ParameterArena | None:
        opt_params = {id(p) for group in optimizer.param_groups for p in group["params"]}
        for module in self._candidate_modules():
            arena = arena_of(module)
            if arena is None:
                continue
            if opt_params and opt_params <= {id(p) for p in module.parameters()}:
                return arena
        return None

    def setup_client(self, config: Config) -> None:
        self.model = self._place_model(self.get_model(config))
        train_loader, val_loader = self.get_data_loaders(config)
        self.train_loader = train_loader
        self.val_loader = val_loader
        self.test_loader = self.get_test_data_loader(config)

        self.num_validation_steps = process_and_check_validation_steps(config, self.val_loader)
        self.num_train_samples = len(self.train_loader.dataset)  # type: ignore[arg-type]
        self.num_val_samples = len(self.val_loader.dataset)  # type: ignore[arg-type]
        if self.num_validation_steps is not None:
            assert self.val_loader.batch_size is not None, (
                "Validation batch size must be defined if we want to limit the number of validation steps"
            )
            self.num_val_samples = self.num_validation_steps * self.val_loader.batch_size
        if self.test_loader:
            self.num_test_samples = len(self.test_loader.dataset)  # type: ignore[arg-type]

        self.set_optimizer(config)
        self._maybe_fuse_optimizers()

        self.lr_schedulers = {}
        for optimizer_key in self.optimizers:
            lr_scheduler = self.get_lr_scheduler(optimizer_key, config)
            if lr_scheduler is not None:
                self.lr_schedulers[optimizer_key] = lr_scheduler

        self.criterion = self.get_criterion(config).to(self.device)
        self.parameter_exchanger = self.get_parameter_exchanger(config)

        self.reports_manager.report({"host_type": "client", "initialized": str(datetime.datetime.now())})
        self.initialized = True

    def get_parameter_exchanger(self, config: Config) -> ParameterExchanger:
        return FullParameterExchanger()

    # ==================================================================================================================
    # model forward + losses (hooks)
    # ==================================================================================================================
    def predict(self, input: TorchInputType) -> tuple[TorchPredType, TorchFeatureType]:
        """Forward pass.  The model may return a tensor, a dict of predictions, or ``(preds_dict, features_dict)``."""
        if isinstance(input, torch.Tensor):
            output = self.model(input)
        elif isinstance(input, dict):
            output = self.model(**input)
        else:
            raise TypeError('"input" must be of type torch.Tensor or dict[str, torch.Tensor].')
        if isinstance(output, dict):
            return output, {}
        if isinstance(output, torch.Tensor):
            return {"prediction": output}, {}
        if isinstance(output, tuple):
            if len(output) != 2:
                raise ValueError(f"Output tuple should have length 2 but has length {len(output)}")
            preds, features = output
            return preds, features
        raise ValueError("Model forward did not return a tensor, dictionary of tensors, or tuple of tensors")

    def compute_loss_and_additional_losses(
        self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType
    ) -> tuple[torch.Tensor, dict[str, torch.Tensor] | None]:
        return self.criterion(preds["prediction"], target), None

    def compute_training_loss(
        self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType
    ) -> TrainingLosses:
        loss, additional_losses = self.compute_loss_and_additional_losses(preds, features, target)
        return TrainingLosses(backward=loss, additional_losses=additional_losses)

    def compute_evaluation_loss(
        self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType
    ) -> EvaluationLosses:
        loss, additional_losses = self.compute_loss_and_additional_losses(preds, features, target)
        return EvaluationLosses(checkpoint=loss, additional_losses=additional_losses)

    def set_optimizer(self, config: Config) -> None:
        optimizer = self.get_optimizer(config)
        assert not isinstance(optimizer, dict), "get_optimizer returned a dict: override set_optimizer to route it"
        self.optimizers = {"global": optimizer}

    # ==================================================================================================================
    # user-supplied pieces
    # ==================================================================================================================
    def get_data_loaders(self, config: Config) -> tuple[DataLoader, DataLoader]:
        raise NotImplementedError

    def get_test_data_loader(self, config: Config) -> DataLoader | None:
        return None

    def transform_target(self, target: TorchTargetType) -> TorchTargetType:
        return target

    def get_criterion(self, config: Config) -> _Loss:
        raise NotImplementedError

    def get_optimizer(self, config: Config) -> Optimizer | dict[str, Optimizer]:
        raise NotImplementedError

    def get_model(self, config: Config) -> nn.Module:
        raise NotImplementedError

    def get_lr_scheduler(self, optimizer_key: str, config: Config) -> LRScheduler | None:
        return None

    def update_lr_schedulers(self, step: int | None = None, epoch: int | None = None) -> None:
        """Step-mode training steps schedulers every step; epoch-mode at the last batch of each epoch."""
        assert (step is None) ^ (epoch is None)
        if step is not None:
            for scheduler in self.lr_schedulers.values():
                scheduler.step()
        elif self.lr_schedulers:
            assert epoch is not None
            seen = getattr(self, "_last_scheduler_epoch", None)
            if seen != (self.total_epochs, epoch):
                # advance once per epoch (first batch of a new epoch marks the boundary of the previous one)
                if seen is not None:
                    for scheduler in self.lr_schedulers.values():
                        scheduler.step()
                self._last_scheduler_epoch = (self.total_epochs, epoch)

    # ------------------------------------------------------------------ lifecycle hooks (no-ops by default)
    def update_before_train(self, current_server_round: int) -> None:
        pass

    def update_after_train(self, local_steps: int, loss_dict: dict[str, float], config: Config) -> None:
        pass

    def update_before_step(self, step: int, current_round: int | None = None) -> None:
        pass

    def update_after_step(self, step: int, current_round: int | None = None) -> None:
        pass

    def update_before_epoch(self, epoch: int) -> None:
        pass

    def transform_gradients(self, losses: TrainingLosses) -> None:
        pass

    # ------------------------------------------------------------------ state
    def _save_client_state(self) -> None:
        self.checkpoint_and_state_module.save_state(self)

    def _load_client_state(self) -> bool:
        return self.checkpoint_and_state_module.maybe_load_state(self)

    # ------------------------------------------------------------------ Flower-style conversion shim
    def to_client(self) -> BasicClient:
        return self
