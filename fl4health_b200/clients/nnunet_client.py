"""nnU-Net client (parity: ``fl4health/clients/nnunet_client.py:71-935``).

The FL-facing behaviour is the reference's: plans negotiated through ``get_properties``, deep-supervision outputs
and targets carried as keyed dicts, mixed precision with a ``GradScaler`` on CUDA, gradient clipping at 12, poly LR,
ignore-label masking and one-hot targets for the metrics.  What differs is the seam to nnU-Net itself: everything that
touches ``nnunetv2`` (fingerprint extraction, experiment planning, preprocessing, trainer / dataloader construction)
sits behind ``NnunetBackend``.  ``Nnunetv2Backend`` imports the optional dependency lazily and raises a clear error if
it is missing; tests (and users with their own segmentation stacks) inject another backend.
"""

from __future__ import annotations

import gc
import pickle
from collections.abc import Sequence
from dataclasses import dataclass
from logging import DEBUG, INFO
from pathlib import Path
from typing import Any, Protocol

import torch
from torch import nn
from torch.nn.modules.loss import _Loss
from torch.optim import Optimizer
from torch.optim.lr_scheduler import _LRScheduler

from fl4health_b200.checkpointing.client_module import ClientCheckpointAndStateModule
from fl4health_b200.clients.basic_client import BasicClient
from fl4health_b200.common.logger import FLOWER_LOGGER as LOGGER
from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import Config, Scalar
from fl4health_b200.engine.options import EngineOptions
from fl4health_b200.metrics.base_metrics import Metric
from fl4health_b200.metrics.metric_managers import MetricManager
from fl4health_b200.reporting.base_reporter import BaseReporter
from fl4health_b200.utils.config import narrow_dict_type
from fl4health_b200.utils.losses import LossMeterType, TrainingLosses
from fl4health_b200.utils.nnunet_utils import (
    NNUNET_N_SPATIAL_DIMS,
    Module2LossWrapper,
    NnunetConfig,
    PolyLRSchedulerWrapper,
    StreamToLogger,
    convert_deep_supervision_dict_to_list,
    convert_deep_supervision_list_to_dict,
    prepare_loss_arg,
)
from fl4health_b200.utils.typing import TorchInputType, TorchPredType, TorchTargetType


@dataclass
class LabelInfo:
    ignore_label: int | None
    has_regions: bool
    num_segmentation_heads: int


@dataclass
class PreparedExperiment:
    """Everything the client needs from the segmentation stack for one (plans, config, fold)."""

    network: nn.Module
    loss: nn.Module
    train_loader: Any
    val_loader: Any
    labels: LabelInfo
    num_input_channels: int
    enable_deep_supervision: bool
    initial_lr: float = 1e-2
    weight_decay: float = 3e-5
    make_optimizer: Any = None  # optional: callable(params) -> Optimizer (defaults to nnU-Net's SGD recipe)


class NnunetBackend(Protocol):
    dataset_name: str

    def plan(self) -> dict[str, Any]:
        """Plan an experiment on the local dataset (used when the server has no global plans)."""

    def prepare(self, plans: dict[str, Any], config: NnunetConfig, fold: int | str, batch_size: int | None, device: torch.device) -> PreparedExperiment:
        """Localise ``plans``, preprocess if needed and build network / loss / loaders."""


class Nnunetv2Backend:
    """Default backend over the optional ``nnunetv2`` package."""

    def __init__(self, dataset_id: int, data_identifier: str | None = None, plans_identifier: str | None = None,
                 always_preprocess: bool = False, n_dataload_processes: int | None = None, trainer_kwargs: dict | None = None,
                 trainer_class: type | None = None) -> None:
        try:
            import nnunetv2  # type: ignore[import-not-found]  # noqa: F401
        except ImportError as exc:
            raise ImportError(
                "NnunetClient's default backend needs the optional 'nnunetv2' (and 'batchgenerators') packages plus the "
                "nnUNet_raw / nnUNet_preprocessed / nnUNet_results environment variables. Install them or pass backend=..."
            ) from exc
        from nnunetv2.utilities.dataset_name_id_conversion import convert_id_to_dataset_name  # type: ignore[import-not-found]

        self.dataset_id = dataset_id
        self.dataset_name = convert_id_to_dataset_name(dataset_id)
        self.data_identifier, self.plans_identifier = data_identifier, plans_identifier
        self.always_preprocess, self.n_dataload_processes = always_preprocess, n_dataload_processes
        self.trainer_kwargs = trainer_kwargs or {}
        self.trainer_class = trainer_class  # an nnUNetTrainer subclass; None = the stock trainer

    # The four steps below carry the names of the reference client's methods (``nnunet_client.py:388-560``); the client
    # exposes them too and simply forwards to its backend.
    def maybe_extract_fingerprint(self) -> None:
        from nnunetv2.experiment_planning.plan_and_preprocess_api import extract_fingerprints  # type: ignore[import-not-found]
        from nnunetv2.paths import nnUNet_preprocessed  # type: ignore[import-not-found]

        fingerprint = Path(nnUNet_preprocessed) / self.dataset_name / "dataset_fingerprint.json"
        if self.always_preprocess or not fingerprint.exists():
            extract_fingerprints(dataset_ids=[self.dataset_id])

    def plan(self) -> dict[str, Any]:
        from nnunetv2.experiment_planning.experiment_planners.default_experiment_planner import ExperimentPlanner  # type: ignore[import-not-found]

        self.maybe_extract_fingerprint()
        plans = ExperimentPlanner(dataset_name_or_id=self.dataset_id, plans_name="temp_plans").plan_experiment()
        plans["plans_name"] = self.dataset_name + "_plans"
        return plans

    def create_plans(self, plans: dict[str, Any], batch_size: int | None = None) -> dict[str, Any]:
        """Localise the federation's plans to this client's dataset and save them next to the preprocessed data."""
        from batchgenerators.utilities.file_and_folder_operations import save_json  # type: ignore[import-not-found]
        from nnunetv2.paths import nnUNet_preprocessed  # type: ignore[import-not-found]

        local = dict(plans)
        local["source_plans_name"] = plans.get("plans_name", "plans")
        local["plans_name"] = self.plans_identifier or f"FL-{local['source_plans_name']}-{self.dataset_id:03d}local"
        local["dataset_name"] = self.dataset_name
        if batch_size is not None:
            for cfg in local["configurations"].values():
                if "batch_size" in cfg:
                    cfg["batch_size"] = batch_size
        plans_path = Path(nnUNet_preprocessed) / self.dataset_name / f"{local['plans_name']}.json"
        plans_path.parent.mkdir(parents=True, exist_ok=True)
        save_json(local, str(plans_path), sort_keys=False)
        return local

    def maybe_preprocess(self, local_plans: dict[str, Any], config: NnunetConfig) -> None:
        from nnunetv2.experiment_planning.plan_and_preprocess_api import preprocess_dataset  # type: ignore[import-not-found]
        from nnunetv2.paths import nnUNet_preprocessed  # type: ignore[import-not-found]

        identifier = (self.data_identifier or local_plans["plans_name"]) + "_" + config.value
        if self.always_preprocess or not (Path(nnUNet_preprocessed) / self.dataset_name / identifier).exists():
            preprocess_dataset(dataset_id=self.dataset_id, plans_identifier=local_plans["plans_name"], configurations=[config.value],
                               num_processes=[self.n_dataload_processes or 4])

    def prepare(self, plans: dict[str, Any], config: NnunetConfig, fold: int | str, batch_size: int | None, device: torch.device) -> PreparedExperiment:
        from batchgenerators.utilities.file_and_folder_operations import load_json  # type: ignore[import-not-found]
        from nnunetv2.paths import nnUNet_preprocessed  # type: ignore[import-not-found]
        from nnunetv2.training.nnUNetTrainer.nnUNetTrainer import nnUNetTrainer  # type: ignore[import-not-found]

        from fl4health_b200.utils.nnunet_utils import NnUNetDataLoaderWrapper

        self.maybe_extract_fingerprint()
        local = self.create_plans(plans, batch_size)
        self.maybe_preprocess(local, config)
        plans_path = Path(nnUNet_preprocessed) / self.dataset_name / f"{local['plans_name']}.json"
        dataset_json = load_json(str(plans_path.parent / "dataset.json"))
        trainer_cls = self.trainer_class if self.trainer_class is not None else nnUNetTrainer
        trainer = trainer_cls(plans=local, configuration=config.value, fold=fold, dataset_json=dataset_json, device=device,
                              **self.trainer_kwargs)
        trainer.initialize()
        train_gen, val_gen = trainer.get_dataloaders()
        labels = trainer.label_manager
        shape = local["configurations"][config.value].get("median_image_size_in_voxels")
        return PreparedExperiment(
            network=trainer.network, loss=trainer.loss,
            train_loader=NnUNetDataLoaderWrapper(train_gen, config, ref_image_shape=shape, n_cases=len(trainer.get_tr_and_val_datasets()[0])),
            val_loader=NnUNetDataLoaderWrapper(val_gen, config, ref_image_shape=shape, n_cases=len(trainer.get_tr_and_val_datasets()[1])),
            labels=LabelInfo(labels.ignore_label, labels.has_regions, labels.num_segmentation_heads),
            num_input_channels=trainer.num_input_channels, enable_deep_supervision=trainer.enable_deep_supervision,
            initial_lr=trainer.initial_lr, weight_decay=trainer.weight_decay,
        )


class NnunetClient(BasicClient):
    def __init__(
        self,
        device: torch.device,
        dataset_id: int,
        fold: int | str,
        data_identifier: str | None = None,
        plans_identifier: str | None = None,
        compile: bool = False,  # noqa: A002
        always_preprocess: bool = False,
        max_grad_norm: float = 12,
        n_dataload_processes: int | None = None,
        verbose: bool = True,
        metrics: Sequence[Metric] | None = None,
        progress_bar: bool = False,
        loss_meter_type: LossMeterType = LossMeterType.AVERAGE,
        checkpoint_and_state_module: ClientCheckpointAndStateModule | None = None,
        reporters: Sequence[BaseReporter] | None = None,
        client_name: str | None = None,
        backend: NnunetBackend | None = None,
        engine_options: EngineOptions | None = None,
        nnunet_trainer_class: type | None = None,
        nnunet_trainer_class_kwargs: dict[str, Any] | None = None,
    ) -> None:
        """``nnunet_trainer_class`` / ``nnunet_trainer_class_kwargs``: an ``nnUNetTrainer`` subclass (and extra constructor
        arguments) for the default backend, as in the reference (``nnunet_client.py:71-150``).

        Config keys required from the server: ``nnunet_config`` (str) and — unless this client is asked to create
        them — ``nnunet_plans`` (pickled dict).  ``compile`` is accepted for API parity; the engine's CUDA-graph capture
        replaces ``torch.compile`` here."""
        super().__init__(
            data_path=Path("dummy/path"), metrics=metrics or [], device=device, loss_meter_type=loss_meter_type,
            checkpoint_and_state_module=checkpoint_and_state_module, reporters=reporters, progress_bar=progress_bar,
            client_name=client_name, engine_options=engine_options,
        )
        self.dataset_id, self.fold = dataset_id, fold
        self.max_grad_norm = max_grad_norm
        self.verbose = verbose
        self.compile = compile
        self.backend: NnunetBackend = backend if backend is not None else Nnunetv2Backend(
            dataset_id, data_identifier, plans_identifier, always_preprocess, n_dataload_processes,
            trainer_kwargs=nnunet_trainer_class_kwargs, trainer_class=nnunet_trainer_class)
        self.dataset_name = self.backend.dataset_name
        self.stream2debug = StreamToLogger(LOGGER, DEBUG)
        self.grad_scaler = torch.amp.GradScaler("cuda", enabled=device.type == "cuda")
        self.experiment: PreparedExperiment
        self.nnunet_config: NnunetConfig
        self.plans: dict[str, Any]

    # ------------------------------------------------------------------------------------------ set-up
    defer_global_model_creation = True  # the network only exists once ``setup_client`` has prepared the experiment

    def setup_client(self, config: Config) -> None:
        self.nnunet_config = NnunetConfig(narrow_dict_type(config, "nnunet_config", str))
        self.plans = pickle.loads(narrow_dict_type(config, "nnunet_plans", bytes))
        batch_size = config.get("batch_size")
        self.experiment = self.backend.prepare(self.plans, self.nnunet_config, self.fold,
                                               int(batch_size) if batch_size is not None else None, self.device)
        super().setup_client(config)

    # reference-named steps, forwarded to the backend when it has them (the injected test / example backends do not)
    def maybe_extract_fingerprint(self) -> None:
        step = getattr(self.backend, "maybe_extract_fingerprint", None)
        if step is not None:
            step()

    def create_plans(self, config: Config) -> dict[str, Any]:
        plans = pickle.loads(narrow_dict_type(config, "nnunet_plans", bytes))
        step = getattr(self.backend, "create_plans", None)
        batch_size = config.get("batch_size")
        return step(plans, int(batch_size) if batch_size is not None else None) if step is not None else plans

    def maybe_preprocess(self, nnunet_config: NnunetConfig) -> None:
        step = getattr(self.backend, "maybe_preprocess", None)
        if step is not None:
            step(self.create_plans({"nnunet_plans": pickle.dumps(self.plans)}), nnunet_config)

    def empty_cache(self) -> None:
        if self.device.type == "cuda":
            torch.cuda.empty_cache()

    def get_model(self, config: Config) -> nn.Module:
        return self.experiment.network

    def get_data_loaders(self, config: Config) -> tuple[Any, Any]:
        return self.experiment.train_loader, self.experiment.val_loader

    def get_criterion(self, config: Config) -> _Loss:
        loss = self.experiment.loss
        return loss if isinstance(loss, _Loss) else Module2LossWrapper(loss)

    def get_optimizer(self, config: Config) -> Optimizer:
        if self.experiment.make_optimizer is not None:
            return self.experiment.make_optimizer(self.model.parameters())
        # nnU-Net's recipe: SGD, nesterov momentum 0.99, weight decay 3e-5
        return torch.optim.SGD(self.model.parameters(), lr=self.experiment.initial_lr, weight_decay=self.experiment.weight_decay,
                               momentum=0.99, nesterov=True)

    def get_lr_scheduler(self, optimizer_key: str, config: Config) -> _LRScheduler:
        """Poly decay over the WHOLE federation (rounds x local steps), constant within each window of one local epoch."""
        if optimizer_key not in self.optimizers:
            raise ValueError(f"Could not find optimizer with key {optimizer_key}")
        n_rounds = narrow_dict_type(config, "n_server_rounds", int)
        if "local_steps" in config:
            steps_per_round = narrow_dict_type(config, "local_steps", int)
        else:
            steps_per_round = narrow_dict_type(config, "local_epochs", int) * len(self.train_loader)
        total = max(1, n_rounds * steps_per_round)
        return PolyLRSchedulerWrapper(self.optimizers[optimizer_key], initial_lr=self.experiment.initial_lr, max_steps=total,
                                      exponent=0.9, steps_per_lr=max(1, len(self.train_loader)))

    def update_lr_schedulers(self, step: int | None = None, epoch: int | None = None) -> None:  # noqa: ARG002
        """The poly schedule is defined over optimizer steps: advance it every step in both training modes (the windowing
        inside ``PolyLRSchedulerWrapper`` reproduces nnU-Net's per-epoch decay)."""
        for scheduler in self.lr_schedulers.values():
            scheduler.step()

    # ------------------------------------------------------------------------------------------ step
    def train_step(self, input: TorchInputType, target: TorchTargetType) -> tuple[TrainingLosses, TorchPredType]:
        if self.device.type != "cuda":
            return super().train_step(input, target)
        optimizer = self.optimizers["global"]
        optimizer.zero_grad()
        preds, features = self.predict(input)
        target = self.transform_target(target)
        losses = self.compute_training_loss(preds, features, target)
        self.grad_scaler.scale(losses.backward["backward"]).backward()
        self.grad_scaler.unscale_(optimizer)
        self.transform_gradients(losses)
        self.grad_scaler.step(optimizer)
        self.grad_scaler.update()
        return losses, preds

    def predict(self, input: TorchInputType) -> tuple[TorchPredType, dict[str, torch.Tensor]]:
        if not isinstance(input, torch.Tensor):
            raise TypeError('"input" must be of type torch.Tensor for nnUNetClient')
        with torch.autocast(self.device.type, enabled=self.device.type == "cuda"):
            output = self.model(input)
        if isinstance(output, torch.Tensor):
            return {"prediction": output}, {}
        if isinstance(output, (list, tuple)):  # deep supervision: one prediction per resolution
            return convert_deep_supervision_list_to_dict(output, NNUNET_N_SPATIAL_DIMS[self.nnunet_config]), {}
        raise TypeError("Was expecting nnunet model output to be either a torch.Tensor or a list/tuple of torch.Tensors")

    def compute_loss_and_additional_losses(
        self, preds: TorchPredType, features: dict[str, torch.Tensor], target: TorchTargetType
    ) -> tuple[torch.Tensor, dict[str, torch.Tensor] | None]:
        loss_preds, loss_targets = prepare_loss_arg(preds), prepare_loss_arg(target)
        assert isinstance(loss_preds, type(loss_targets)), (
            f"Got unexpected types for preds and targets: {type(loss_preds)} and {type(loss_targets)}")
        if isinstance(loss_preds, list):
            assert len(loss_preds) == len(loss_targets), (
                f"Got {len(loss_preds)} predictions and {len(loss_targets)} targets: deep supervision must match on both sides")
        with torch.autocast(self.device.type, enabled=self.device.type == "cuda"):
            return self.criterion(loss_preds, loss_targets), None

    def transform_gradients(self, losses: TrainingLosses) -> None:
        nn.utils.clip_grad_norm_(self.model.parameters(), self.max_grad_norm)

    # ------------------------------------------------------------------------------------------ metrics
    def mask_data(self, pred: torch.Tensor, target: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        """Zero predictions where the target carries the ignore label and drop that label from the target."""
        labels = self.experiment.labels
        if labels.has_regions:  # one-hot target whose LAST channel is the ignore label
            mask = ~target[:, -1:] if target.dtype == torch.bool else 1 - target[:, -1:]
            target = target[:, :-1]
        else:
            mask = (target != labels.ignore_label).float()
            target = torch.where(target == labels.ignore_label, torch.zeros_like(target), target)
        return pred * mask.expand(-1, pred.shape[1], *mask.shape[2:]), target

    def update_metric_manager(self, preds: TorchPredType, target: TorchTargetType, metric_manager: MetricManager) -> None:
        # personalised wrappers prefix the twin models' outputs with "global-" / "local-": metrics follow the personal
        # (local) model; without prefixes both steps are no-ops (parity: clients/flexible/nnunet.py:799-803)
        preds = {k.removeprefix("local-"): v for k, v in preds.items() if not k.startswith("global")}
        m_pred = convert_deep_supervision_dict_to_list(preds)[0] if len(preds) > 1 else next(iter(preds.values()))
        if isinstance(target, torch.Tensor):
            m_target = target
        elif isinstance(target, dict):
            m_target = convert_deep_supervision_dict_to_list(target)[0] if len(target) > 1 else next(iter(target.values()))
        else:
            raise TypeError("Was expecting target to be type dict[str, torch.Tensor] or torch.Tensor")
        if m_pred.ndim != m_target.ndim:
            m_target = m_target.view(m_target.shape[0], 1, *m_target.shape[1:])
        labels = self.experiment.labels
        if labels.ignore_label is not None and not labels.has_regions:
            m_pred, m_target = self.mask_data(m_pred, m_target)  # index-encoded target: mask before one-hot encoding
        if m_pred.shape != m_target.shape:
            m_target = torch.zeros(m_pred.shape, device=m_pred.device, dtype=torch.bool).scatter_(1, m_target.long(), 1)
        elif labels.ignore_label is not None and labels.has_regions:
            m_pred, m_target = self.mask_data(m_pred, m_target)
        metric_manager.update({"prediction": m_pred}, m_target)

    # ------------------------------------------------------------------------------------------ protocol
    def get_properties(self, config: Config) -> dict[str, Scalar]:
        if "nnunet_plans" not in config:
            log(INFO, "Initializing the global plans using local dataset")
            config["nnunet_plans"] = pickle.dumps(self.backend.plan())
        properties = super().get_properties(config)
        if not self.initialized:
            self.setup_client(config)
        properties["nnunet_plans"] = config["nnunet_plans"]
        properties["num_input_channels"] = self.experiment.num_input_channels
        properties["num_segmentation_heads"] = self.experiment.labels.num_segmentation_heads
        properties["enable_deep_supervision"] = self.experiment.enable_deep_supervision
        return properties

    def get_client_specific_reports(self) -> dict[str, Any]:
        return {"learning_rate": float(self.optimizers["global"].param_groups[0]["lr"])}

    def update_before_train(self, current_server_round: int) -> None:
        gc.collect()
        if current_server_round == 2:  # after the first round's allocations: freezing makes later collections cheap
            gc.freeze()
        super().update_before_train(current_server_round)

    def shutdown_dataloader(self, dataloader: Any, dl_name: str | None = None) -> None:
        if dataloader is not None and hasattr(dataloader, "shutdown"):
            if self.verbose:
                log(INFO, f"\tShutting down nnunet dataloader: {dl_name}")
            dataloader.shutdown()

    def shutdown(self) -> None:
        gc.unfreeze()
        gc.collect()
        for name in ("train_loader", "val_loader", "test_loader"):
            self.shutdown_dataloader(getattr(self, name, None), name)
        super().shutdown()
