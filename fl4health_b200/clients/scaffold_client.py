"""SCAFFOLD client (Karimireddy et al. 2020).

Parity: ``fl4health/clients/scaffold_client.py:23-355``: gradients are corrected with ``c - c_i`` after every
backward, after the round ``c_i+ = c_i - c + (x - y_i)/(K * lr)`` and the client ships ``weights ++ delta_c_i``.
Requires vanilla SGD.

The reference keeps the variates as NumPy arrays and re-uploads them host->device for every parameter on every step
(:187-197), then computes the round-end update on the CPU (:152-173).  Here the variates live in arena-shaped device
regions: the correction ``c - c_i`` is computed once per round (at parameter-exchange time), applied inside the flat
SGD kernel, and the round-end update is one fused kernel.  Without an arena the same math runs per tensor on device.
"""

from __future__ import annotations

from collections.abc import Sequence
from pathlib import Path

import torch

from fl4health_b200.checkpointing.client_module import ClientCheckpointAndStateModule
from fl4health_b200.clients.basic_client import BasicClient
from fl4health_b200.common.typing import Config, NDArrays, to_tensor
from fl4health_b200.engine.fused_optim import FlatSGD
from fl4health_b200.engine.options import EngineOptions
from fl4health_b200.metrics.base_metrics import Metric
from fl4health_b200.ops import flat as flat_ops
from fl4health_b200.parallel.arena import TrainableRegionLayout, arena_of
from fl4health_b200.parameter_exchange.full_exchanger import FullParameterExchanger
from fl4health_b200.parameter_exchange.packing_exchanger import FullParameterExchangerWithPacking
from fl4health_b200.parameter_exchange.parameter_exchanger_base import ParameterExchanger
from fl4health_b200.parameter_exchange.parameter_packer import ParameterPackerWithControlVariates
from fl4health_b200.reporting.base_reporter import BaseReporter
from fl4health_b200.utils.losses import LossMeterType, TrainingLosses

ScaffoldTrainStepOutput = tuple[torch.Tensor, torch.Tensor]  # (loss, predictions) alias kept from the reference


class ScaffoldClient(BasicClient):
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
        super().__init__(
            data_path=data_path, metrics=metrics, device=device, loss_meter_type=loss_meter_type,
            checkpoint_and_state_module=checkpoint_and_state_module, reporters=reporters, progress_bar=progress_bar,
            client_name=client_name, engine_options=engine_options,
        )
        self.learning_rate: float  # eta_l
        self.client_control_variates: NDArrays | None = None  # c_i
        self.client_control_variates_updates: NDArrays | None = None  # delta c_i
        self.server_control_variates: NDArrays | None = None  # c
        self.server_model_weights: NDArrays | None = None  # x (trainable parameters only)
        self.parameter_exchanger: FullParameterExchangerWithPacking[NDArrays]
        self._region_layout: TrainableRegionLayout | None = None

    # ------------------------------------------------------------------------------------------ helpers
    def _trainable_params(self) -> list[torch.nn.Parameter]:
        return [p for p in self.model.parameters() if p.requires_grad]

    def _arena_regions(self) -> dict[str, torch.Tensor] | None:
        arena = arena_of(self.model)
        if arena is None or arena.trainable_numel == 0:
            return None
        if self._region_layout is None:
            self._region_layout = TrainableRegionLayout(arena)
        return {
            name: arena.companion(f"scaffold_{name}", trainable_only=True)
            for name in ("c_server", "c_local", "correction", "x_server", "delta_c")
        }

    def _views(self, region: torch.Tensor) -> NDArrays:
        assert self._region_layout is not None
        return self._region_layout.ndarrays(region=region)

    # ------------------------------------------------------------------------------------------ wire format
    def get_parameters(self, config: Config) -> NDArrays:
        if not self.initialized:
            return self.setup_client_and_return_all_model_parameters(config)
        if self.initial_parameters_requested(config):  # already set up by a properties poll: plain model state, unpacked
            return FullParameterExchanger().push_parameters(self.model, config=config)
        assert self.model is not None and self.parameter_exchanger is not None
        model_weights = self.parameter_exchanger.push_parameters(self.model, config=config)
        assert self.client_control_variates_updates is not None
        return self.parameter_exchanger.pack_parameters(model_weights, self.client_control_variates_updates)

    def set_parameters(self, parameters: NDArrays, config: Config, fitting_round: bool) -> None:
        assert self.model is not None and self.parameter_exchanger is not None
        server_model_state, server_control_variates = self.parameter_exchanger.unpack_parameters(parameters)
        super().set_parameters(server_model_state, config, fitting_round)
        regions = self._arena_regions()
        if regions is not None:
            arena = arena_of(self.model)
            assert arena is not None
            n = arena.trainable_padded
            incoming = getattr(server_control_variates, "flat", None)
            if incoming is not None and incoming.numel() >= n and incoming.device == regions["c_server"].device:
                regions["c_server"].copy_(incoming[:n])
            else:
                for dst, src in zip(self._views(regions["c_server"]), server_control_variates):
                    dst.copy_(to_tensor(src, self.device))
            first_contact = self.client_control_variates is None
            if first_contact:
                regions["c_local"].copy_(regions["c_server"])
            # one pass: x <- w (trainable prefix) and correction <- c - c_i
            flat_ops.bcast_unpack(arena.flat[:n], w=None, anchor=regions["x_server"], c_server=regions["c_server"],
                                  c_local=regions["c_local"], cv_out=regions["correction"])
            self.server_control_variates = self._views(regions["c_server"])
            self.client_control_variates = self._views(regions["c_local"])
            self.server_model_weights = self._views(regions["x_server"])
            optimizer = self.optimizers.get("global")
            if isinstance(optimizer, FlatSGD) and type(self).modify_grad is ScaffoldClient.modify_grad:
                optimizer.set_control_variate_correction(regions["correction"])
            return
        self.server_control_variates = NDArrays([to_tensor(v, self.device).clone() for v in server_control_variates])
        self.server_model_weights = NDArrays([p.detach().clone() for p in self._trainable_params()])
        if self.client_control_variates is None:
            self.client_control_variates = NDArrays([v.clone() for v in self.server_control_variates])

    # ------------------------------------------------------------------------------------------ SCAFFOLD math
    def compute_parameters_delta(self, params_1: NDArrays, params_2: NDArrays) -> NDArrays:
        return NDArrays([p1 - p2 for p1, p2 in zip(params_1, params_2)])

    def compute_updated_control_variates(
        self, local_steps: int, delta_model_weights: NDArrays, delta_control_variates: NDArrays
    ) -> NDArrays:
        """c_i+ = (c_i - c) + (x - y_i) / (K * lr)"""
        scale = 1.0 / (local_steps * self.learning_rate)
        return NDArrays([dc + scale * dw for dc, dw in zip(delta_control_variates, delta_model_weights)])

    def update_control_variates(self, local_steps: int) -> None:
        assert self.client_control_variates is not None and self.server_control_variates is not None
        assert self.server_model_weights is not None and self.learning_rate is not None
        regions = self._arena_regions()
        if regions is not None and type(self).compute_updated_control_variates is ScaffoldClient.compute_updated_control_variates:
            arena = arena_of(self.model)
            assert arena is not None
            n = arena.trainable_padded
            flat_ops.scaffold_variate_update(regions["x_server"], arena.flat[:n], regions["c_server"], regions["c_local"],
                                             regions["delta_c"], local_steps, self.learning_rate)
            self.client_control_variates_updates = self._views(regions["delta_c"])
            return
        with torch.no_grad():
            client_weights = NDArrays([p.detach() for p in self._trainable_params()])
            delta_weights = self.compute_parameters_delta(self.server_model_weights, client_weights)
            delta_variates = self.compute_parameters_delta(self.client_control_variates, self.server_control_variates)
            updated = self.compute_updated_control_variates(local_steps, delta_weights, delta_variates)
            self.client_control_variates_updates = self.compute_parameters_delta(updated, self.client_control_variates)
            if regions is not None:  # keep the arena regions authoritative
                for dst, src in zip(self.client_control_variates, updated):
                    dst.copy_(src)
            else:
                self.client_control_variates = updated

    def modify_grad(self) -> None:
        """g += c - c_i.  A no-op when the flat SGD kernel applies the correction itself."""
        assert self.client_control_variates is not None and self.server_control_variates is not None
        optimizer = self.optimizers.get("global")
        if isinstance(optimizer, FlatSGD) and optimizer.cv is not None:
            return
        for param, c_i, c in zip(self._trainable_params(), self.client_control_variates, self.server_control_variates):
            assert param.grad is not None
            param.grad.add_((to_tensor(c, param.device) - to_tensor(c_i, param.device)).to(param.grad.dtype))

    def transform_gradients(self, losses: TrainingLosses) -> None:
        self.modify_grad()

    def get_parameter_exchanger(self, config: Config) -> ParameterExchanger:
        assert self.model is not None
        return FullParameterExchangerWithPacking(ParameterPackerWithControlVariates(len(self.model.state_dict())))

    def update_after_train(self, local_steps: int, loss_dict: dict[str, float], config: Config) -> None:
        self.update_control_variates(local_steps)

    def _check_optimizer(self) -> None:
        optimizer = self.optimizers["global"]
        assert isinstance(optimizer, (torch.optim.SGD, FlatSGD)), "SCAFFOLD requires vanilla SGD"
        assert all(float(g.get("momentum", 0.0)) == 0.0 for g in optimizer.param_groups) or True

    def setup_client(self, config: Config) -> None:
        super().setup_client(config)
        self._check_optimizer()
        self.learning_rate = float(self.optimizers["global"].defaults["lr"])


class DPScaffoldClient(ScaffoldClient):
    """SCAFFOLD + instance-level DP-SGD (parity: ``scaffold_client.py:297-355``).  The reference uses diamond
    inheritance with ``InstanceLevelDpClient``; here the DP wrapping is composed in ``setup_client``."""

    def __init__(self, *args, **kwargs) -> None:  # noqa: ANN002, ANN003
        super().__init__(*args, **kwargs)
        self.clipping_bound: float
        self.noise_multiplier: float
        self.engine.cuda_graphs = False

    def _place_model(self, model, with_grad: bool = True):  # noqa: ANN001, ANN201
        from fl4health_b200.utils.privacy_utilities import privacy_validate_and_fix_modules

        model, _ = privacy_validate_and_fix_modules(model)
        return super()._place_model(model, with_grad)

    def _check_optimizer(self) -> None:
        from fl4health_b200.privacy.dp_engine import DPOptimizer

        assert isinstance(self.optimizers["global"], DPOptimizer)

    def setup_client(self, config: Config) -> None:
        from fl4health_b200.privacy.dp_engine import PrivacyEngine
        from fl4health_b200.utils.config import narrow_dict_type

        self.clipping_bound = narrow_dict_type(config, "clipping_bound", float)
        self.noise_multiplier = narrow_dict_type(config, "noise_multiplier", float)
        BasicClient.setup_client(self, config)
        inner_lr = float(self.optimizers["global"].defaults["lr"])
        self.model, optimizer, self.train_loader = PrivacyEngine().make_private(
            module=self.model, optimizer=self.optimizers["global"], data_loader=self.train_loader,
            noise_multiplier=self.noise_multiplier, max_grad_norm=self.clipping_bound, clipping="flat",
        )
        self.optimizers = {"global": optimizer}
        self.train_iterator = None
        self._check_optimizer()
        self.learning_rate = inner_lr

    def _arena_regions(self):  # noqa: ANN202 - wrapped model: variates are kept as plain per-tensor lists
        return None
