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
from fl4health_b200.common.typing import Config, NDArrays, Scalar, to_tensor
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
        home = getattr(model, "_device", torch.device("cpu"))
        components, values = (to_tensor(array, home).float() for array in parameters[:2])
        model.set_principal_components(components, values)


class FedPCAClient:
    """Not a ``BasicClient``: there is no iterative training.  ``fit`` = one SVD of the local data matrix; ``evaluate`` =
    how well the merged subspace explains the local validation matrix."""

    def __init__(self, data_path: Path, device: torch.device, model_save_dir: Path, client_name: str | None = None,
                 metrics: Sequence[Metric] | None = None) -> None:
        self.data_path, self.model_save_dir = data_path, model_save_dir
        self.device = torch.device(device)
        self.client_name = client_name if client_name is not None else self.generate_hash()
        self.parameter_exchanger: ParameterExchanger = _PcaExchanger()
        self.initialized = False
        self.model: PcaModule
        self.train_data_tensor: Tensor
        self.val_data_tensor: Tensor
        self.num_train_samples: int
        self.num_val_samples: int

    def generate_hash(self, length: int = 8) -> str:
        """Unique id used as the client name when none is given (parity: ``fed_pca_client.py:44-55``)."""
        return generate_hash(length)

    # ------------------------------------------------------------------------------------------ set-up
    def get_data_loaders(self, config: Config) -> tuple[DataLoader, DataLoader]:
        raise NotImplementedError

    def get_model(self, config: Config) -> PcaModule:
        options = {key: narrow_dict_type(config, key, kind) for key, kind in
                   (("low_rank", bool), ("full_svd", bool), ("rank_estimation", int))}
        return PcaModule(**options)

    def get_data_tensor(self, data_loader: DataLoader) -> Tensor:
        """Whole dataset as one matrix (default: concatenate the loader's batches)."""
        return torch.cat([features for features, *_ in data_loader], dim=0)

    def setup_client(self, config: Config) -> None:
        self.model = self.get_model(config).to(self.device)
        self.model._device = self.device  # type: ignore[assignment]
        loaders = dict(zip(("train", "val"), self.get_data_loaders(config)))
        for split, loader in loaders.items():
            setattr(self, f"{split}_data_tensor", self.get_data_tensor(loader).to(self.device))
            setattr(self, f"num_{split}_samples", len(loader.dataset))  # type: ignore[arg-type]
        self.initialized = True

    def _decompose_local_data(self, center_data: bool) -> None:
        self.model.set_principal_components(*self.model(self.train_data_tensor, center_data))

    # ------------------------------------------------------------------------------------------ protocol
    def get_parameters(self, config: Config) -> NDArrays:
        if not self.initialized:
            log(INFO, "Setting up client and providing full model parameters to the server for initialization")
            self.setup_client(config)
            self._decompose_local_data(bool(config.get("center_data", True)))
        return self.parameter_exchanger.push_parameters(self.model, config=config)

    def set_parameters(self, parameters: NDArrays, config: Config) -> None:
        self.parameter_exchanger.pull_parameters(parameters, self.model, config)
        self.save_model()

    def fit(self, parameters: NDArrays, config: Config) -> tuple[NDArrays, int, dict[str, Scalar]]:
        if not self.initialized:
            self.setup_client(config)
        self._decompose_local_data(narrow_dict_type(config, "center_data", bool))
        summary: dict[str, Scalar] = {
            "cumulative_explained_variance": self.model.compute_cumulative_explained_variance(),
            "top_explained_variance_ratio": self.model.compute_explained_variance_ratios()[0].item(),
        }
        return self.get_parameters(config), self.num_train_samples, summary

    def evaluate(self, parameters: NDArrays, config: Config) -> tuple[float, int, dict[str, Scalar]]:
        if not self.initialized:
            self.setup_client(config)
        model = self.model
        if not hasattr(model, "data_mean"):  # evaluate-before-fit: centre with the local training mean
            model.set_data_mean(model.maybe_reshape(self.train_data_tensor))
        self.set_parameters(parameters, config)
        kept = narrow_dict_type(config, "num_components_eval", int) if "num_components_eval" in config else None
        held_out = model.center_data(model.maybe_reshape(self.val_data_tensor)).to(self.device)
        error = model.compute_reconstruction_error(held_out, kept)
        return error, self.num_val_samples, {"projection_variance": model.compute_projection_variance(held_out, kept)}

    def get_properties(self, config: Config) -> dict[str, Scalar]:
        if not self.initialized:
            self.setup_client(config)
        return {"num_train_samples": self.num_train_samples, "num_val_samples": self.num_val_samples}

    def save_model(self) -> None:
        destination = Path(self.model_save_dir) / f"client_{self.client_name}_pca.pt"
        torch.save(self.model, destination)
        log(INFO, f"Model parameters saved to {destination}.")

    def shutdown(self) -> None:
        pass
