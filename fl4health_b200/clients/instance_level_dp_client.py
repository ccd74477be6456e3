"""Instance-level DP client (parity: ``fl4health/clients/instance_level_dp_client.py:17-114``): DP-SGD with flat
clipping + Gaussian noise through the in-house engine (``privacy.dp_engine``), Poisson batch sampling.
Config keys: ``clipping_bound`` (float), ``noise_multiplier`` (float)."""

from __future__ import annotations

from typing import Any

from fl4health_b200.clients.basic_client import BasicClient
from fl4health_b200.common.typing import Config
from fl4health_b200.privacy.dp_engine import PrivacyEngine
from fl4health_b200.utils.config import narrow_dict_type
from fl4health_b200.utils.privacy_utilities import privacy_validate_and_fix_modules


class InstanceLevelDpClient(BasicClient):
    def __init__(self, *args: Any, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        self.clipping_bound: float
        self.noise_multiplier: float
        # The DP step is capturable: the engine book-keeps per-sample norms instead of materialising per-sample
        # gradients and draws its noise from a device-resident stream (privacy/dp_engine.py).  Poisson sampling makes
        # the batch size vary, so one graph is kept per size seen (engine/graph_runner.py).  A flat gradient region is
        # kept (no pointer-table gradients): clipped sums are written into it and noised by ONE kernel.
        from dataclasses import replace

        self.engine = replace(self.engine, table_grads=False)

    def _place_model(self, model, with_grad: bool = True):  # noqa: ANN001, ANN201
        # fix DP-incompatible layers BEFORE the arena is laid out so optimizer/arena see the final parameters
        model, _ = privacy_validate_and_fix_modules(model)
        return super()._place_model(model, with_grad)

    def setup_client(self, config: Config) -> None:
        self.clipping_bound = narrow_dict_type(config, "clipping_bound", float)
        self.noise_multiplier = narrow_dict_type(config, "noise_multiplier", float)
        super().setup_client(config)
        self.setup_opacus_objects(config)

    def setup_opacus_objects(self, config: Config) -> None:
        privacy_engine = PrivacyEngine()
        self.model, optimizer, self.train_loader = privacy_engine.make_private(
            module=self.model, optimizer=self.optimizers["global"], data_loader=self.train_loader,
            noise_multiplier=self.noise_multiplier, max_grad_norm=self.clipping_bound, clipping="flat",
        )
        self.optimizers = {"global": optimizer}
        self.train_iterator = None
