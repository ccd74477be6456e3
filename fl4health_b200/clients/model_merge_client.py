"""One-shot model-merging client (parity: ``fl4health/clients/model_merge_client.py:23-256``): loads a locally trained
model, evaluates it, uploads its weights; later evaluates the merged model."""

from __future__ import annotations

import datetime
from abc import abstractmethod
from collections.abc import Sequence
from pathlib import Path

import torch
from torch import nn
from torch.utils.data import DataLoader

from fl4health_b200.common.typing import Config, NDArrays, Scalar
from fl4health_b200.metrics.base_metrics import Metric
from fl4health_b200.metrics.metric_managers import MetricManager
from fl4health_b200.parameter_exchange.full_exchanger import FullParameterExchanger
from fl4health_b200.parameter_exchange.parameter_exchanger_base import ParameterExchanger
from fl4health_b200.reporting.base_reporter import BaseReporter
from fl4health_b200.reporting.reports_manager import ReportsManager
from fl4health_b200.utils.client import move_data_to_device
from fl4health_b200.utils.random import generate_hash


class ModelMergeClient:
    def __init__(
        self,
        data_path: Path,
        model_path: Path,
        metrics: Sequence[Metric],
        device: torch.device,
        reporters: Sequence[BaseReporter] | None = None,
        client_name: str | None = None,
    ) -> None:
        self.data_path, self.model_path, self.metrics, self.device = data_path, model_path, metrics, torch.device(device)
        self.client_name = client_name if client_name is not None else generate_hash()
        self.initialized = False
        self.test_metric_manager = MetricManager(metrics=self.metrics, metric_manager_name="test")
        self.reports_manager = ReportsManager(reporters)
        self.reports_manager.initialize(id=self.client_name)
        self.model: nn.Module
        self.test_loader: DataLoader
        self.num_test_samples: int

    def setup_client(self, config: Config) -> None:
        self.model = self.get_model(config).to(self.device)
        self.test_loader = self.get_test_data_loader(config)
        self.num_test_samples = len(self.test_loader.dataset)  # type: ignore[arg-type]
        self.parameter_exchanger = self.get_parameter_exchanger(config)
        self.initialized = True

    def get_parameters(self, config: Config) -> NDArrays:
        assert self.model is not None
        return self.parameter_exchanger.push_parameters(self.model, config=config)

    def set_parameters(self, parameters: NDArrays, config: Config) -> None:
        assert self.initialized
        self.parameter_exchanger.pull_parameters(parameters, self.model)

    def fit(self, parameters: NDArrays, config: Config) -> tuple[NDArrays, int, dict[str, Scalar]]:
        assert not self.initialized
        self.setup_client(config)
        self.reports_manager.report({"host_type": "client", "fit_start": str(datetime.datetime.now())})
        val_metrics = self.validate()
        self.reports_manager.report({"fit_metrics": val_metrics, "host_type": "client", "fit_end": str(datetime.datetime.now())})
        return self.get_parameters(config), self.num_test_samples, val_metrics

    def validate(self) -> dict[str, Scalar]:
        self.model.eval()
        self.test_metric_manager.clear()
        with torch.no_grad():
            for input, target in self.test_loader:
                input, target = move_data_to_device(input, self.device), move_data_to_device(target, self.device)
                self.test_metric_manager.update({"predictions": self.model(input)}, target)
        return self.test_metric_manager.compute()

    def evaluate(self, parameters: NDArrays, config: Config) -> tuple[float, int, dict[str, Scalar]]:
        self.set_parameters(parameters, config)
        return 0.0, len(self.test_loader), self.validate()

    def get_properties(self, config: Config) -> dict[str, Scalar]:
        if not self.initialized:
            self.setup_client(config)
        return {"num_test_samples": self.num_test_samples}

    def get_parameter_exchanger(self, config: Config) -> ParameterExchanger:
        return FullParameterExchanger()

    def shutdown(self) -> None:
        self.reports_manager.shutdown()

    @abstractmethod
    def get_model(self, config: Config) -> nn.Module:
        raise NotImplementedError

    @abstractmethod
    def get_test_data_loader(self, config: Config) -> DataLoader:
        raise NotImplementedError
