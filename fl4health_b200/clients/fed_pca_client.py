"""Federated PCA client (parity: ``fl4health/clients/fed_pca_client.py:18-235``): SVD of the local data matrix, ships
``(principal_components, singular_values)``; evaluation = reconstruction error / projected variance of the merged
subspace on local validation data."""

from __future__ import annotations

from collections.abc import Sequence
from logging import INFO
from pathlib import Path

import torch
from torch import Tensor
from torch.utils.data import DataLoader

from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import Config, NDArrays, Scalar
from fl4health_b200.metrics.base_metrics import Metric
from fl4health_b200.model_bases.pca import PcaModule
from fl4health_b200.parameter_exchange.full_exchanger import FullParameterExchanger
from fl4health_b200.parameter_exchange.parameter_exchanger_base import ParameterExchanger
from fl4health_b200.utils.config import narrow_dict_type
from fl4health_b200.utils.random import generate_hash


class _PcaExchanger(FullParameterExchanger):
    """(components, singular values) <-> ``PcaModule`` parameters (shapes change between rounds, so set, not copy)."""

    def push_parameters(self, model, initial_model=None, config=None) -> NDArrays:  # noqa: ANN001
        return NDArrays([model.principal_components.detach(), model.singular_values.detach()])

    def pull_parameters(self, parameters: NDArrays, model, config=None) -> None:  # noqa: ANN001
        from fl4health_b200.common.typing import to_tensor

        device = next(iter(model.buffers()), torch.zeros(0)).device if False else getattr(model, "_device", torch.device("cpu"))
        model.set_principal_components(to_tensor(parameters[0], device).float(), to_tensor(parameters[1], device).float())


class FedPCAClient:
    def __init__(self, data_path: Path, device: torch.device, model_save_dir: Path, client_name: str | None = None,
                 metrics: Sequence[Metric] | None = None) -> None:
        self.client_name = self.generate_hash() if client_name is None else client_name
        self.model: PcaModule
        self.initialized = False
        self.data_path = data_path
        self.model_save_dir = model_save_dir
        self.device = torch.device(device)
        self.train_data_tensor: Tensor
        self.val_data_tensor: Tensor
        self.num_train_samples: int
        self.num_val_samples: int
        self.parameter_exchanger: ParameterExchanger = _PcaExchanger()

    def generate_hash(self, length: int = 8) -> str:
        """Unique id used as the client name when none is given (parity: ``fed_pca_client.py:44-55``)."""
        return generate_hash(length)

    def get_parameters(self, config: Config) -> NDArrays:
        if not self.initialized:
            log(INFO, "Setting up client and providing full model parameters to the server for initialization")
            self.setup_client(config)
            components, values = self.model(self.train_data_tensor, bool(config.get("center_data", True)))
            self.model.set_principal_components(components, values)
        return self.parameter_exchanger.push_parameters(self.model, config=config)

    def set_parameters(self, parameters: NDArrays, config: Config) -> None:
        self.parameter_exchanger.pull_parameters(parameters, self.model, config)
        self.save_model()

    def get_data_loaders(self, config: Config) -> tuple[DataLoader, DataLoader]:
        raise NotImplementedError

    def get_model(self, config: Config) -> PcaModule:
        return PcaModule(narrow_dict_type(config, "low_rank", bool), narrow_dict_type(config, "full_svd", bool),
                         narrow_dict_type(config, "rank_estimation", int))

    def setup_client(self, config: Config) -> None:
        self.model = self.get_model(config).to(self.device)
        self.model._device = self.device  # type: ignore[assignment]
        train_loader, val_loader = self.get_data_loaders(config)
        self.train_data_tensor = self.get_data_tensor(train_loader).to(self.device)
        self.val_data_tensor = self.get_data_tensor(val_loader).to(self.device)
        self.num_train_samples = len(train_loader.dataset)  # type: ignore[arg-type]
        self.num_val_samples = len(val_loader.dataset)  # type: ignore[arg-type]
        self.initialized = True

    def get_data_tensor(self, data_loader: DataLoader) -> Tensor:
        """Whole dataset as one matrix (default: concatenate the loader's batches)."""
        return torch.cat([batch[0] for batch in data_loader], dim=0)

    def fit(self, parameters: NDArrays, config: Config) -> tuple[NDArrays, int, dict[str, Scalar]]:
        if not self.initialized:
            self.setup_client(config)
        center_data = narrow_dict_type(config, "center_data", bool)
        components, values = self.model(self.train_data_tensor, center_data)
        self.model.set_principal_components(components, values)
        ratios = self.model.compute_explained_variance_ratios()
        metrics: dict[str, Scalar] = {
            "cumulative_explained_variance": self.model.compute_cumulative_explained_variance(),
            "top_explained_variance_ratio": ratios[0].item(),
        }
        return self.get_parameters(config), self.num_train_samples, metrics

    def evaluate(self, parameters: NDArrays, config: Config) -> tuple[float, int, dict[str, Scalar]]:
        if not self.initialized:
            self.setup_client(config)
        if not hasattr(self.model, "data_mean"):
            self.model.set_data_mean(self.model.maybe_reshape(self.train_data_tensor))
        self.set_parameters(parameters, config)
        k = narrow_dict_type(config, "num_components_eval", int) if "num_components_eval" in config else None
        val = self.model.center_data(self.model.maybe_reshape(self.val_data_tensor)).to(self.device)
        loss = self.model.compute_reconstruction_error(val, k)
        return loss, self.num_val_samples, {"projection_variance": self.model.compute_projection_variance(val, k)}

    def get_properties(self, config: Config) -> dict[str, Scalar]:
        if not self.initialized:
            self.setup_client(config)
        return {"num_train_samples": self.num_train_samples, "num_val_samples": self.num_val_samples}

    def save_model(self) -> None:
        path = Path(self.model_save_dir) / f"client_{self.client_name}_pca.pt"
        torch.save(self.model, path)
        log(INFO, f"Model parameters saved to {path}.")

    def shutdown(self) -> None:
        pass
