"""``BasicClient``: the template-method FL client every algorithm client derives from.

API parity: ``fl4health/clients/basic_client.py:43-1321`` — same constructor, the same ~15 overridable hooks
(``get_model``, ``get_optimizer``, ``get_data_loaders``, ``get_criterion``, ``predict``,
``compute_loss_and_additional_losses``, ``transform_gradients``, ``update_before_train`` ...), the same
``fit``/``evaluate``/``get_parameters``/``get_properties`` protocol, report keys and config keys.

The class is a thin API shell: the round protocol is parsed by ``engine/round_protocol.RoundPlan``, the loops live in
``engine/local_loop`` (one schedule-driven driver instead of four hand-written loops) and the per-batch step runs through
``engine/step_executor.StepExecutor``.  What is different is *how the hooks are executed* (``EngineOptions``):

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
from fl4health_b200.engine import outputs as model_outputs
from fl4health_b200.engine.companions import Companion, build_companions, companion_modules
from fl4health_b200.engine.fused_optim import translate_optimizer
from fl4health_b200.engine.graph_runner import GraphStepRunner
from fl4health_b200.engine.local_loop import BatchCycler, EpochSchedule, StepSchedule, run_evaluation, run_training
from fl4health_b200.engine.options import EngineOptions
from fl4health_b200.engine.round_protocol import RoundPlan, Stopwatch
from fl4health_b200.engine.step_executor import StepExecutor
from fl4health_b200.metrics.base_metrics import TEST_LOSS_KEY, TEST_NUM_EXAMPLES_KEY, Metric
from fl4health_b200.metrics.metric_managers import MetricManager
from fl4health_b200.parallel.arena import ParameterArena, arena_of, attach_arena
from fl4health_b200.utils import tracing
from fl4health_b200.parameter_exchange.full_exchanger import FullParameterExchanger
from fl4health_b200.parameter_exchange.parameter_exchanger_base import ParameterExchanger
from fl4health_b200.reporting.base_reporter import BaseReporter
from fl4health_b200.reporting.reports_manager import ReportsManager
from fl4health_b200.utils.client import (
    fold_loss_dict_into_metrics,
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

_SPLITS = ("train", "val", "test")


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
        self.data_path, self.metrics, self.progress_bar = data_path, metrics, progress_bar
        self.device = torch.device(device)
        self.client_name = client_name or generate_hash()
        log(INFO, f"Client Name: {self.client_name}")
        self._executor = StepExecutor(engine_options or EngineOptions.from_env(), self.device, self.client_name)

        self.checkpoint_and_state_module = checkpoint_and_state_module or ClientCheckpointAndStateModule(
            pre_aggregation=None, post_aggregation=None, state_checkpointer=None
        )
        self.reports_manager = ReportsManager(reporters)
        self.reports_manager.initialize(id=self.client_name, name=self.client_name)

        # one loss meter + metric manager per data split (train_loss_meter, val_metric_manager, ...)
        for split in _SPLITS:
            losses_type = TrainingLosses if split == "train" else EvaluationLosses
            setattr(self, f"{split}_loss_meter", LossMeter[losses_type](loss_meter_type, losses_type))
            setattr(self, f"{split}_metric_manager", MetricManager(metrics=self.metrics, metric_manager_name=split))

        self.initialized = False
        self._answering_fit = False
        self.initial_weights: NDArrays | None = None
        self.total_steps = self.total_epochs = 0
        self.num_test_samples: int | None = None
        self.learning_rate: float | None = None
        self.early_stopper: EarlyStopper | None = None
        self.num_validation_steps: int | None = None
        self._train_batches: BatchCycler | None = None
        self._val_batches: BatchCycler | None = None

        # bound by setup_client (declared for type checkers and subclasses)
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

    # ------------------------------------------------------------------------------------------------------------------
    # set-up: user factories -> placed model, loaders, (fused) optimizers, schedulers, criterion, exchanger
    # ------------------------------------------------------------------------------------------------------------------
    companions: dict[str, Companion] = {}  # extra models kept next to ``self.model`` (engine/companions.py)

    @property
    def engine(self) -> EngineOptions:
        """How the hooks are executed.  One object, shared with the step executor: replacing it (a client variant
        turning graphs or table gradients off) is seen by both."""
        return self._executor.engine

    @engine.setter
    def engine(self, options: EngineOptions) -> None:
        self._executor.engine = options

    def setup_client(self, config: Config) -> None:
        build_companions(self, config)
        self.model = self._place_model(self.get_model(config))
        self.train_loader, self.val_loader = self.get_data_loaders(config)
        self.test_loader = self.get_test_data_loader(config)
        self._train_batches = self._val_batches = None
        self._count_samples(config)

        self.set_optimizer(config)
        self._maybe_fuse_optimizers()
        candidates = ((key, self.get_lr_scheduler(key, config)) for key in self.optimizers)
        self.lr_schedulers = {key: scheduler for key, scheduler in candidates if scheduler is not None}

        self.criterion = self.get_criterion(config).to(self.device)
        self.parameter_exchanger = self.get_parameter_exchanger(config)
        self.reports_manager.report({"host_type": "client", "initialized": str(datetime.datetime.now())})
        self.initialized = True

    def _count_samples(self, config: Config) -> None:
        def size(loader: DataLoader) -> int:
            return len(loader.dataset)  # type: ignore[arg-type]

        self.num_train_samples, self.num_val_samples = size(self.train_loader), size(self.val_loader)
        self.num_validation_steps = process_and_check_validation_steps(config, self.val_loader)
        if self.num_validation_steps is not None:  # capped validation: the sample count follows the cap
            batch = self.val_loader.batch_size
            assert batch is not None, "Validation batch size must be defined if we want to limit the number of validation steps"
            self.num_val_samples = self.num_validation_steps * batch
        if self.test_loader:
            self.num_test_samples = size(self.test_loader)

    def _place_model(self, model: nn.Module, with_grad: bool = True) -> nn.Module:
        """Move a model to the device and (engine option) re-home its state into a flat arena."""
        model = model.to(self.device)
        options = self.engine
        if not options.arena:
            return model.to(memory_format=torch.channels_last) if options.channels_last else model
        arena = attach_arena(model, self.device, with_grad=with_grad, channels_last=options.channels_last,
                             allocator=self._arena_allocator())
        if with_grad and options.master_weights and options.fused_optimizer and options.amp_dtype is not None:
            arena.enable_compute_shadow(options.amp_dtype)
        elif with_grad and options.table_grads and options.fused_optimizer:
            arena.use_table_gradients()
        return model

    def _arena_allocator(self) -> Any:
        """Hook for the SPMD runtime: allocate arenas from peer-mapped symmetric memory."""
        return getattr(self, "arena_allocator", None)

    def _candidate_modules(self) -> list[nn.Module]:
        """Modules whose arenas an optimizer may be operating on: the model and the trainable companions."""
        return [self.model, *companion_modules(self, trainable_only=True)]

    def _arena_for_optimizer(self, optimizer: Optimizer) -> ParameterArena | None:
        """The arena of the candidate module that owns every parameter of ``optimizer`` (None: keep the stock one)."""
        wanted = {id(p) for group in optimizer.param_groups for p in group["params"]}
        if not wanted:
            return None
        for module in self._candidate_modules():
            arena = arena_of(module)
            if arena is not None and wanted <= {id(p) for p in module.parameters()}:
                return arena
        return None

    def _maybe_fuse_optimizers(self) -> None:
        if self.engine.fused_optimizer:
            arenas = {key: self._arena_for_optimizer(opt) for key, opt in self.optimizers.items()}
            self.optimizers.update({key: translate_optimizer(self.optimizers[key], arena)
                                    for key, arena in arenas.items() if arena is not None})

    def set_optimizer(self, config: Config) -> None:
        optimizer = self.get_optimizer(config)
        assert not isinstance(optimizer, dict), "get_optimizer returned a dict: override set_optimizer to route it"
        self.optimizers = {"global": optimizer}

    def get_parameter_exchanger(self, config: Config) -> ParameterExchanger:
        return FullParameterExchanger()

    # ------------------------------------------------------------------------------------------------------------------
    # parameters in / out
    # ------------------------------------------------------------------------------------------------------------------
    def get_parameters(self, config: Config) -> NDArrays:
        """Arrays for the server.  Before the client is set up, this sets it up and returns ALL model state (the
        server uses it to initialise the global model)."""
        if not self.initialized:
            return self.setup_client_and_return_all_model_parameters(config)
        self._maybe_load_saved_best_local_model_state()
        return self.parameter_exchanger.push_parameters(self.model, config=config)

    def setup_client_and_return_all_model_parameters(self, config: Config) -> NDArrays:
        log(INFO, "Setting up client and providing full model parameters to the server for initialization")
        if not config:
            log(WARNING, "This client has not yet been initialized and the config is empty. This may cause unexpected "
                         "failures, as setting up a client typically requires several configuration parameters, "
                         "including batch_size and current_server_round.")
        self.setup_client(config)
        return FullParameterExchanger().push_parameters(self.model, config=config)

    def initial_parameters_requested(self, config: Config) -> bool:
        """True when ``get_parameters`` is the server's request for INITIAL parameters (round 0 of
        ``on_init_parameters_config_fn``) reaching a client that has already been set up — a properties poll (nnU-Net plan
        negotiation, tabular feature alignment) may have initialised it.  Clients that pack side information into their
        regular payload must answer this request with the plain model state, exactly like an uninitialised client."""
        return self.initialized and config.get("current_server_round") == 0 and not getattr(self, "_answering_fit", False)

    def set_parameters(self, parameters: NDArrays, config: Config, fitting_round: bool) -> None:
        """The very first fit installs *all* weights (full exchange) whatever the exchanger; afterwards the client's
        exchanger decides which state is overwritten."""
        first_fit = fitting_round and narrow_dict_type(config, "current_server_round", int) == 1
        if first_fit:
            self.initialize_all_model_weights(parameters, config)
        else:
            self.parameter_exchanger.pull_parameters(parameters, self.model, config)

    def initialize_all_model_weights(self, parameters: NDArrays, config: Config) -> None:
        FullParameterExchanger().pull_parameters(parameters, self.model, config)

    def _maybe_load_saved_best_local_model_state(self) -> None:
        stopper = self.early_stopper
        if stopper is not None and stopper.patience is None:
            log(INFO, "Loading saved best model's state before sending model to server.")
            stopper.load_snapshot(["model"])

    def get_properties(self, config: Config) -> dict[str, Scalar]:
        if not self.initialized:
            self.setup_client(config)
        return {"num_train_samples": self.num_train_samples, "num_val_samples": self.num_val_samples}

    def shutdown(self) -> None:
        self.reports_manager.report({"shutdown": str(datetime.datetime.now())})
        self.reports_manager.shutdown()

    # ------------------------------------------------------------------------------------------------------------------
    # the round protocol
    # ------------------------------------------------------------------------------------------------------------------
    def process_config(self, config: Config) -> tuple[int | None, int | None, int, bool, bool]:
        return RoundPlan.from_config(config).as_tuple()

    def fit(self, parameters: NDArrays, config: Config) -> tuple[NDArrays, int, dict[str, Scalar]]:
        clock = Stopwatch()
        clock.mark("round")
        epochs, steps, server_round, evaluate_after_fit, pack_losses = self.process_config(config)
        state_io = self.checkpoint_and_state_module.state_checkpointer is not None
        if not self.initialized:
            self.setup_client(config)
            if state_io:
                log(INFO, "Successfully loaded client state." if self._load_client_state() else "Client state was not loaded.")

        with tracing.phase("pull_parameters"):
            self.set_parameters(parameters, config, fitting_round=True)
        with tracing.phase("update_before_train"):
            self.update_before_train(server_round)

        clock.mark("train")
        with tracing.phase("local_train"):
            if epochs is not None:
                loss_dict, metrics = self.train_by_epochs(epochs, server_round)
                steps = len(self.train_loader) * epochs
            else:
                assert steps is not None
                loss_dict, metrics = self.train_by_steps(steps, server_round)
        clock.mark("trained")
        with tracing.phase("update_after_train"):
            self.update_after_train(steps, loss_dict, config)

        if self._should_evaluate_after_fit(evaluate_after_fit):
            validation_loss, validation_metrics = self.validate(pack_losses)
            metrics.update(validation_metrics)
            self._maybe_checkpoint(validation_loss, validation_metrics, CheckpointMode.PRE_AGGREGATION)

        report = {"fit_round_metrics": metrics, "fit_round_losses": loss_dict, "round": server_round,
                  "round_start": str(clock.marks["round"]), "round_end": str(clock.mark("reported")),
                  **clock.span("fit_round", "train", "trained"), "fit_step": self.total_steps, "fit_epoch": self.total_epochs}
        if tracing.tracing_enabled():
            report["device_phase_ms"] = tracing.phase_report()
        self.reports_manager.report(report, server_round)
        if state_io:
            self._save_client_state()
        with tracing.phase("push_parameters"):
            self._answering_fit = True  # round 0 can be a FIT too (SCAFFOLD's warm start): its answer is the regular payload
            try:
                return self.get_parameters(config), self.num_train_samples, metrics
            finally:
                self._answering_fit = False

    def evaluate(self, parameters: NDArrays, config: Config) -> tuple[float, int, dict[str, Scalar]]:
        if not self.initialized:
            self.setup_client(config)
        clock = Stopwatch()
        clock.mark("begin")
        server_round = narrow_dict_type(config, "current_server_round", int)
        with tracing.phase("pull_parameters"):
            self.set_parameters(parameters, config, fitting_round=False)
        with tracing.phase("evaluate"):
            loss, metrics = self.validate(set_pack_losses_with_val_metrics(config))
        clock.mark("end")
        self._maybe_checkpoint(loss, metrics, CheckpointMode.POST_AGGREGATION)
        self.reports_manager.report(
            {"eval_round_metrics": metrics, "eval_round_loss": loss, **clock.span("eval_round", "begin", "end"),
             "fit_step": self.total_steps, "fit_epoch": self.total_epochs, "round": server_round},
            server_round,
        )
        return loss, self.num_val_samples, metrics

    def _should_evaluate_after_fit(self, evaluate_after_fit: bool) -> bool:
        return evaluate_after_fit or self.checkpoint_and_state_module.pre_aggregation is not None

    def _maybe_checkpoint(self, loss: float, metrics: dict[str, Scalar], checkpoint_mode: CheckpointMode) -> None:
        self.checkpoint_and_state_module.maybe_checkpoint(self.model, loss, metrics, checkpoint_mode)

    # ------------------------------------------------------------------------------------------------------------------
    # one batch: the hook bodies, and the engine units wrapped around them
    # ------------------------------------------------------------------------------------------------------------------
    def train_step(self, input: TorchInputType, target: TorchTargetType) -> tuple[TrainingLosses, TorchPredType]:
        optimizer = self.optimizers["global"]
        optimizer.zero_grad()
        with self._amp():
            preds, features = self.predict(input)
            losses = self.compute_training_loss(preds, features, self.transform_target(target))
        losses.backward["backward"].backward()
        self.transform_gradients(losses)
        optimizer.step()
        return losses, preds

    def val_step(self, input: TorchInputType, target: TorchTargetType) -> tuple[EvaluationLosses, TorchPredType]:
        with torch.no_grad(), self._amp():
            preds, features = self.predict(input)
            losses = self.compute_evaluation_loss(preds, features, self.transform_target(target))
        return losses, preds

    def _amp(self) -> contextlib.AbstractContextManager:
        return self._executor.autocast()

    def _prepare_batch(self, input: TorchInputType, target: TorchTargetType) -> tuple[TorchInputType, TorchTargetType]:
        return self._executor.stage(input, target)

    def _graph_variant(self) -> Any:
        """Hashable tag of everything *besides input shapes* that changes what ``train_step`` launches (e.g. FedRep's
        head/representation phase).  One captured graph is kept per (variant, input signature)."""
        return "default"

    def _invalidate_graphs(self) -> None:
        """Drop captured graphs (call after anything that re-binds tensors the step reads: new model, new optimizer)."""
        self._executor.reset()

    def _sync_optimizer_hyperparams(self) -> None:
        StepExecutor.push_hyperparameters(self.optimizers)

    def _train_unit(self, input: TorchInputType, target: TorchTargetType) -> tuple[TrainingLosses, TorchPredType]:
        """``train_step`` + device-side loss / metric accumulation: the unit that is captured and replayed."""
        losses, preds = self.train_step(input, target)
        self.train_loss_meter.accumulate(losses)
        self.update_metric_manager(preds, target, self.train_metric_manager)
        return losses.detach(), {name: tensor.detach() for name, tensor in preds.items()}  # type: ignore[return-value]

    def _run_train_unit(self, input: TorchInputType, target: TorchTargetType) -> tuple[TrainingLosses, TorchPredType]:
        outcome = self._executor.run_train(self._train_unit, self._graph_variant(), self._sync_optimizer_hyperparams, input, target)
        self.train_loss_meter.mark_step(losses=outcome[0])
        return outcome

    def _run_val_unit(
        self, input: TorchInputType, target: TorchTargetType, loss_meter: LossMeter, metric_manager: MetricManager
    ) -> tuple[EvaluationLosses, TorchPredType]:
        def unit(batch_input: TorchInputType, batch_target: TorchTargetType) -> tuple[EvaluationLosses, TorchPredType]:
            losses, preds = self.val_step(batch_input, batch_target)
            loss_meter.accumulate(losses)
            self.update_metric_manager(preds, batch_target, metric_manager)
            return losses, preds

        outcome = self._executor.run_eval(unit, (id(loss_meter), self._graph_variant()), input, target)
        loss_meter.mark_step(losses=outcome[0])
        return outcome

    # captured-graph bookkeeping, exposed for tests / profiling scripts
    @property
    def _train_runner(self) -> GraphStepRunner | None:
        return self._executor.latest_train_runner

    @property
    def _train_runners(self) -> dict[Any, GraphStepRunner]:
        return self._executor.train_runners

    @property
    def _val_runners(self) -> dict[Any, GraphStepRunner]:
        return self._executor.eval_runners

    def update_metric_manager(self, preds: TorchPredType, target: TorchTargetType, metric_manager: MetricManager) -> None:
        metric_manager.update(preds, target)

    # ------------------------------------------------------------------------------------------------------------------
    # local training / evaluation (schedules walked by engine/local_loop)
    # ------------------------------------------------------------------------------------------------------------------
    def train_by_epochs(self, epochs: int, current_round: int | None = None) -> tuple[dict[str, float], dict[str, Scalar]]:
        return run_training(self, EpochSchedule(self.train_loader, epochs), current_round)

    def train_by_steps(self, steps: int, current_round: int | None = None) -> tuple[dict[str, float], dict[str, Scalar]]:
        return run_training(self, StepSchedule(self._train_source(), steps), current_round)

    def _train_source(self) -> BatchCycler:
        if self._train_batches is None or self._train_batches.loader is not self.train_loader:
            self._train_batches = BatchCycler(self.train_loader)
        return self._train_batches

    def _next_train_batch(self) -> tuple[TorchInputType, TorchTargetType]:
        return self._train_source().draw()

    @property
    def train_iterator(self) -> Iterator | None:
        """Position inside the training loader (step mode); assigning None rewinds it."""
        return None if self._train_batches is None else self._train_batches._it

    @train_iterator.setter
    def train_iterator(self, value: Iterator | None) -> None:
        if value is None:
            self._train_batches = None
        else:
            self._train_source()._it = value

    def validate(self, include_losses_in_metrics: bool = False) -> tuple[float, dict[str, Scalar]]:
        if self.num_validation_steps is None:
            batches: Any = self.val_loader
        else:
            if self._val_batches is None or self._val_batches.loader is not self.val_loader:
                self._val_batches = BatchCycler(self.val_loader)
            batches = self._val_batches.take(self.num_validation_steps)
        val_loss, val_metrics = self._fully_validate_or_test(
            batches, self.val_loss_meter, self.val_metric_manager, LoggingMode.VALIDATION, include_losses_in_metrics)
        if self.test_loader:
            test_loss, test_metrics = self._fully_validate_or_test(
                self.test_loader, self.test_loss_meter, self.test_metric_manager, LoggingMode.TEST, include_losses_in_metrics)
            if self.num_test_samples is not None:
                val_metrics[TEST_NUM_EXAMPLES_KEY] = self.num_test_samples
            val_metrics[TEST_LOSS_KEY] = test_loss
            val_metrics.update(test_metrics)
        return val_loss, val_metrics

    def _fully_validate_or_test(
        self, loader: Any, loss_meter: LossMeter, metric_manager: MetricManager,
        logging_mode: LoggingMode = LoggingMode.VALIDATION, include_losses_in_metrics: bool = False,
    ) -> tuple[float, dict[str, Scalar]]:
        """Evaluate every batch of ``loader`` (any iterable of batches); returns (checkpoint loss, metrics)."""
        assert logging_mode in (LoggingMode.VALIDATION, LoggingMode.TEST, LoggingMode.EARLY_STOP_VALIDATION)
        run_evaluation(self, loader, loss_meter, metric_manager)
        loss_dict, metrics = loss_meter.compute().as_dict(), metric_manager.compute()
        self._log_results(loss_dict, metrics, logging_mode=logging_mode)
        if include_losses_in_metrics:
            fold_loss_dict_into_metrics(metrics, loss_dict, logging_mode)
        return loss_dict["checkpoint"], metrics

    # ------------------------------------------------------------------------------------------------------------------
    # logging / reporting
    # ------------------------------------------------------------------------------------------------------------------
    def _log_header_str(self, current_round: int | None = None, current_epoch: int | None = None,
                        logging_mode: LoggingMode = LoggingMode.TRAIN) -> None:
        where = [text for text, value in ((f"Current FL Round: {current_round}", current_round),
                                          (f"Current Epoch: {current_epoch}", current_epoch)) if value is not None]
        log(INFO, f"{logging_mode.value} | " + "\t".join(where))

    def _log_results(self, loss_dict: dict[str, float], metrics_dict: dict[str, Scalar], current_round: int | None = None,
                     current_epoch: int | None = None, logging_mode: LoggingMode = LoggingMode.TRAIN) -> None:
        mode = logging_mode.value
        summary = [f"Client {mode} Losses: " + ", ".join(f"{name}: {value:.6f}" for name, value in loss_dict.items())]
        if metrics_dict:
            summary.append(f"Client {mode} Metrics: " + ", ".join(f"{name}: {value}" for name, value in metrics_dict.items()))
        log(INFO, " | ".join(summary))
        for level, message in self.get_client_specific_logs(current_round, current_epoch, logging_mode)[1]:
            log(level.value, message)

    def _step_reports_enabled(self) -> bool:
        if self.engine.step_reports is not None:
            return self.engine.step_reports
        return any(getattr(reporter, "wants_step_reports", False) for reporter in self.reports_manager.reporters)

    def get_client_specific_logs(self, current_round: int | None, current_epoch: int | None,
                                 logging_mode: LoggingMode) -> tuple[str, list[tuple[LogLevel, str]]]:
        """Hook: extra header text + log lines (e.g. current FedProx mu)."""
        return "", []

    def get_client_specific_reports(self) -> dict[str, Any]:
        """Hook: extra key/values merged into reporter payloads."""
        return {}

    # ------------------------------------------------------------------------------------------------------------------
    # model forward + losses
    # ------------------------------------------------------------------------------------------------------------------
    def predict(self, input: TorchInputType) -> tuple[TorchPredType, TorchFeatureType]:
        """Forward pass.  The model may return a tensor, a dict of predictions, or ``(preds_dict, features_dict)``."""
        return model_outputs.forward(self.model, input)

    def compute_loss_and_additional_losses(
        self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType
    ) -> tuple[torch.Tensor, dict[str, torch.Tensor] | None]:
        return self.criterion(preds["prediction"], target), None

    def compute_training_loss(self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType) -> TrainingLosses:
        main, extra = self.compute_loss_and_additional_losses(preds, features, target)
        return TrainingLosses(backward=main, additional_losses=extra)

    def compute_evaluation_loss(self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType) -> EvaluationLosses:
        main, extra = self.compute_loss_and_additional_losses(preds, features, target)
        return EvaluationLosses(checkpoint=main, additional_losses=extra)

    def update_lr_schedulers(self, step: int | None = None, epoch: int | None = None) -> None:
        """Called after every batch in both training modes (exactly one of ``step`` / ``epoch`` is passed): every
        scheduler advances once per call, as in the reference; override for per-epoch stepping."""
        assert (step is None) != (epoch is None)
        for scheduler in self.lr_schedulers.values():
            scheduler.step()

    # ------------------------------------------------------------------------------------------------------------------
    # user-supplied factories and lifecycle hooks
    # ------------------------------------------------------------------------------------------------------------------
    def get_model(self, config: Config) -> nn.Module:
        raise NotImplementedError

    def get_data_loaders(self, config: Config) -> tuple[DataLoader, DataLoader]:
        raise NotImplementedError

    def get_test_data_loader(self, config: Config) -> DataLoader | None:
        return None

    def get_criterion(self, config: Config) -> _Loss:
        raise NotImplementedError

    def get_optimizer(self, config: Config) -> Optimizer | dict[str, Optimizer]:
        raise NotImplementedError

    def get_lr_scheduler(self, optimizer_key: str, config: Config) -> LRScheduler | None:
        return None

    def transform_target(self, target: TorchTargetType) -> TorchTargetType:
        return target

    def transform_gradients(self, losses: TrainingLosses) -> None:
        pass

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

    def stop_after_epoch(self, epoch: int) -> bool:
        """Hook (epoch-based training only): return True to end local training after this epoch."""
        return False

    # ------------------------------------------------------------------------------------------------------------------
    # persisted client state, Flower-style conversion shim
    # ------------------------------------------------------------------------------------------------------------------
    def _save_client_state(self) -> None:
        self.checkpoint_and_state_module.save_state(self)

    def _load_client_state(self) -> bool:
        return self.checkpoint_and_state_module.maybe_load_state(self)

    def to_client(self) -> BasicClient:
        return self
