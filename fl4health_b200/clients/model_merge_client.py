"""One-shot model-merging client (parity: ``fl4health/clients/model_merge_client.py:23-256``): loads a locally trained
model, evaluates it, uploads its weights; later evaluates the merged model."""

from __future__ import annotations

import datetime
from abc import abstractmethod
from collections.abc import Sequence
from pathlib import Path
from typing import Any

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
    """Life cycle: ``fit`` (exactly once) = load the locally trained network, score it, upload it;
    ``evaluate`` (afterwards) = install the merged weights, score again.  There is no training, so the client owns just a
    network, one held-out loader and one metric manager."""

    def __init__(
        self,
        data_path: Path,
        model_path: Path,
        metrics: Sequence[Metric],
        device: torch.device,
        reporters: Sequence[BaseReporter] | None = None,
        client_name: str | None = None,
    ) -> None:
        self.data_path, self.model_path, self.metrics = data_path, model_path, metrics
        self.device = torch.device(device)
        self.client_name = client_name or generate_hash()
        self.test_metric_manager = MetricManager(metrics=self.metrics, metric_manager_name="test")
        self.reports_manager = ReportsManager(reporters)
        self.reports_manager.initialize(id=self.client_name)
        self.initialized = False
        self.model: nn.Module
        self.test_loader: DataLoader
        self.num_test_samples: int

    def setup_client(self, config: Config) -> None:
        self.test_loader = self.get_test_data_loader(config)
        self.num_test_samples = len(self.test_loader.dataset)  # type: ignore[arg-type]
        self.model = self.get_model(config).to(self.device)
        self.parameter_exchanger = self.get_parameter_exchanger(config)
        self.initialized = True

    def _stamp(self, key: str, **extra: Any) -> None:
        self.reports_manager.report({"host_type": "client", key: str(datetime.datetime.now()), **extra})

    # ---------------------------------------------------------------------------------------------- protocol
    def fit(self, parameters: NDArrays, config: Config) -> tuple[NDArrays, int, dict[str, Scalar]]:
        assert not self.initialized
        self.setup_client(config)
        self._stamp("fit_start")
        scores = self.validate()
        self._stamp("fit_end", fit_metrics=scores)
        return self.get_parameters(config), self.num_test_samples, scores

    def evaluate(self, parameters: NDArrays, config: Config) -> tuple[float, int, dict[str, Scalar]]:
        self.set_parameters(parameters, config)
        return 0.0, len(self.test_loader), self.validate()

    def get_parameters(self, config: Config) -> NDArrays:
        return self.parameter_exchanger.push_parameters(self.model, config=config)

    def set_parameters(self, parameters: NDArrays, config: Config) -> None:
        assert self.initialized
        self.parameter_exchanger.pull_parameters(parameters, self.model)

    def get_properties(self, config: Config) -> dict[str, Scalar]:
        if not self.initialized:
            self.setup_client(config)
        return {"num_test_samples": self.num_test_samples}

    def validate(self) -> dict[str, Scalar]:
        """Metrics of the current weights on the held-out loader (predictions keyed ``"predictions"``)."""
        manager = self.test_metric_manager
        manager.clear()
        self.model.eval()
        with torch.no_grad():
            for batch in self.test_loader:
                features, labels = (move_data_to_device(part, self.device) for part in batch)
                manager.update({"predictions": self.model(features)}, labels)
        return manager.compute()

    def shutdown(self) -> None:
        self.reports_manager.shutdown()

    # ---------------------------------------------------------------------------------------------- factories
    def get_parameter_exchanger(self, config: Config) -> ParameterExchanger:
        return FullParameterExchanger()

    @abstractmethod
    def get_model(self, config: Config) -> nn.Module:
        raise NotImplementedError

    @abstractmethod
    def get_test_data_loader(self, config: Config) -> DataLoader:
        raise NotImplementedError
