"""nnU-Net client / server protocol with an injected toy segmentation backend (the nnunetv2 backend is an optional
dependency): plans negotiation, deep-supervision dict plumbing, ignore-label masking, poly LR, utilities."""

import pickle
from pathlib import Path

import pytest
import torch
from torch import nn

from fl4health_b200.checkpointing.checkpointer import LatestTorchModuleCheckpointer
from fl4health_b200.checkpointing.server_module import NnUnetServerCheckpointAndStateModule
from fl4health_b200.clients.nnunet_client import LabelInfo, NnunetClient, PreparedExperiment
from fl4health_b200.metrics.efficient_metrics import MultiClassDice
from fl4health_b200.metrics.metric_aggregation import evaluate_metrics_aggregation_fn, fit_metrics_aggregation_fn
from fl4health_b200.servers.client_manager import SimpleClientManager
from fl4health_b200.servers.nnunet_server import NnunetServer
from fl4health_b200.simulation import run_simulation
from fl4health_b200.strategies.basic_fedavg import BasicFedAvg
from fl4health_b200.utils.nnunet_utils import (
    LocalPolyLRScheduler,
    NnunetConfig,
    NnUNetDataLoaderWrapper,
    PolyLRSchedulerWrapper,
    collapse_one_hot_tensor,
    convert_deep_supervision_dict_to_list,
    convert_deep_supervision_list_to_dict,
    get_dataset_n_voxels,
    get_segs_from_probs,
    prepare_loss_arg,
)
from fl4health_b200.utils.random import set_all_random_seeds


class ToySegNet(nn.Module):
    def __init__(self, in_ch: int, heads: int, deep_supervision: bool) -> None:
        super().__init__()
        self.body = nn.Conv2d(in_ch, 8, 3, padding=1)
        self.head = nn.Conv2d(8, heads, 1)
        self.low = nn.Conv2d(8, heads, 1)
        self.deep_supervision = deep_supervision

    def forward(self, x):
        h = torch.relu(self.body(x))
        full = self.head(h)
        if self.deep_supervision and self.training:
            return [full, self.low(nn.functional.avg_pool2d(h, 2))]
        return full


class DeepSupervisionLoss(nn.Module):
    def forward(self, preds, targets):
        if isinstance(preds, list):
            return sum(w * nn.functional.cross_entropy(p, t[:, 0].long()) for w, p, t in zip((1.0, 0.5), preds, targets))
        return nn.functional.cross_entropy(preds, targets[:, 0].long())


class ToyAugmenter:
    """Infinite generator of ``{"data", "target"}`` batches like nnU-Net's augmenters (targets: list when deep supervision)."""

    def __init__(self, seed: int, deep_supervision: bool, n: int = 4, batch_size: int = 4) -> None:
        self.gen = torch.Generator().manual_seed(seed)
        self.deep_supervision, self.n, self.batch_size = deep_supervision, n, batch_size

    def __iter__(self):
        while True:
            x = torch.randn(self.batch_size, 2, 16, 16, generator=self.gen)
            y = (x[:, :1] + x[:, 1:] > 0).long() + (x[:, :1] > 1).long()  # classes 0,1,2
            yield {"data": x, "target": [y, y[:, :, ::2, ::2]] if self.deep_supervision else y}


class ToyBackend:
    dataset_name = "Dataset999_Toy"

    def __init__(self, seed: int, deep_supervision: bool = True) -> None:
        self.seed, self.deep_supervision = seed, deep_supervision
        self.planned = False

    def plan(self):
        self.planned = True
        return {"plans_name": "toy_plans", "configurations": {"2d": {"median_image_size_in_voxels": [16, 16], "batch_size": 4}}}

    def prepare(self, plans, config, fold, batch_size, device):
        assert config is NnunetConfig._2D and plans["plans_name"] == "toy_plans"
        torch.manual_seed(0)
        return PreparedExperiment(
            network=ToySegNet(2, 3, self.deep_supervision), loss=DeepSupervisionLoss(),
            train_loader=NnUNetDataLoaderWrapper(ToyAugmenter(self.seed, self.deep_supervision), config, set_len=4),
            val_loader=NnUNetDataLoaderWrapper(ToyAugmenter(100 + self.seed, False), config, set_len=2),
            labels=LabelInfo(ignore_label=None, has_regions=False, num_segmentation_heads=3), num_input_channels=2,
            enable_deep_supervision=self.deep_supervision, initial_lr=0.05,
        )


def _cfg(r: int):
    return {"current_server_round": r, "local_epochs": 1, "batch_size": 4, "nnunet_config": "2d", "n_server_rounds": 3}


def test_nnunet_federation_negotiates_plans_and_trains(tmp_path: Path) -> None:
    set_all_random_seeds(8)
    dice = MultiClassDice(batch_dim=None, label_dim=1, threshold=1)
    clients = [NnunetClient(torch.device("cpu"), 999, fold=0, metrics=[dice], backend=ToyBackend(i), client_name=f"n{i}", verbose=False)
               for i in range(2)]
    strategy = BasicFedAvg(min_fit_clients=2, min_evaluate_clients=2, min_available_clients=2, on_fit_config_fn=_cfg,
                           on_evaluate_config_fn=_cfg, fit_metrics_aggregation_fn=fit_metrics_aggregation_fn,
                           evaluate_metrics_aggregation_fn=evaluate_metrics_aggregation_fn)
    module = NnUnetServerCheckpointAndStateModule(model_checkpointers=LatestTorchModuleCheckpointer(str(tmp_path), "nnunet.pt"))
    built = {}

    def builder(plans, config, in_ch, heads, deep_supervision):
        built.update(plans=plans, in_ch=in_ch, heads=heads)
        return ToySegNet(in_ch, heads, deep_supervision)

    server = NnunetServer(SimpleClientManager(), {"n_server_rounds": 3, "nnunet_config": "2d"}, _cfg, strategy,
                          checkpoint_and_state_module=module, model_builder=builder)
    history = run_simulation(server, clients, 3)
    assert built["in_ch"] == 2 and built["heads"] == 3 and built["plans"]["plans_name"] == "toy_plans"
    assert sum(c.backend.planned for c in clients) == 1  # exactly one client planned; the other received the plans
    assert pickle.loads(server.nnunet_plans_bytes)["plans_name"] == "toy_plans"
    losses = [v for _, v in history.losses_distributed]
    assert len(losses) == 3 and losses[-1] < losses[0]
    assert "val - prediction - MultiClassDice" in history.metrics_distributed
    assert (tmp_path / "nnunet.pt").exists()
    lrs = clients[0].optimizers["global"].param_groups[0]["lr"]
    assert lrs < 0.05  # poly decay across rounds


def test_ignore_label_masking_and_metric_targets() -> None:
    client = NnunetClient(torch.device("cpu"), 999, fold=0, backend=ToyBackend(0), verbose=False)
    client.experiment = PreparedExperiment(nn.Identity(), nn.Identity(), None, None, LabelInfo(2, False, 2), 1, False)
    pred = torch.ones(1, 2, 2, 2)
    target = torch.tensor([[[[0, 1], [2, 2]]]])
    masked, new_target = client.mask_data(pred, target.clone())
    assert masked[0, :, 1].sum() == 0 and masked[0, :, 0].sum() == 4 and new_target.max() == 1
    seen = {}

    class Spy:
        def update(self, preds, tgt):
            seen["pred"], seen["target"] = preds["prediction"], tgt

    client.update_metric_manager({"0-2x2": pred, "1-1x1": pred[:, :, :1, :1]}, {"0-2x2": target, "1-1x1": target[:, :, :1, :1]}, Spy())
    assert seen["target"].shape == (1, 2, 2, 2) and seen["target"].dtype == torch.bool
    assert seen["pred"][0, :, 1].sum() == 0  # ignored voxels removed from the prediction too


def test_nnunet_utils() -> None:
    tensors = [torch.zeros(2, 3, 8, 8), torch.zeros(2, 3, 4, 4)]
    keyed = convert_deep_supervision_list_to_dict(tensors, 2)
    assert list(keyed) == ["0-8x8", "1-4x4"] and convert_deep_supervision_dict_to_list(dict(reversed(keyed.items())))[0].shape[-1] == 8
    assert isinstance(prepare_loss_arg(keyed), list) and prepare_loss_arg({"0-8x8": tensors[0]}) is tensors[0]
    with pytest.raises(ValueError):
        prepare_loss_arg([1, 2])  # type: ignore[arg-type]
    probs = torch.tensor([[[0.2, 0.7]], [[0.8, 0.3]]]).permute(1, 0, 2).reshape(1, 2, 1, 2)
    assert get_segs_from_probs(probs)[0, :, 0, 0].tolist() == [0, 1]
    regions = get_segs_from_probs(torch.tensor([[[[0.9]], [[0.9]], [[0.2]]]]), has_regions=True)
    assert regions[0, :, 0, 0].tolist() == [False, False, False]  # background voxel: every region masked out
    assert collapse_one_hot_tensor(torch.tensor([[0, 1], [1, 0]]), dim=0).tolist() == [1, 0]
    plans = {"configurations": {"2d": {"median_image_size_in_voxels": [10, 10]}, "3d_fullres": {"median_image_size_in_voxels": [4, 10, 10]}}}
    assert get_dataset_n_voxels(plans, 5) == 2000.0
    opt = torch.optim.SGD([nn.Parameter(torch.zeros(1))], lr=0.1)
    sched = PolyLRSchedulerWrapper(opt, initial_lr=0.1, max_steps=100, steps_per_lr=25)
    seen = []
    for _ in range(100):
        opt.step()
        sched.step()
        seen.append(opt.param_groups[0]["lr"])
    assert seen[0] == pytest.approx(0.1) and seen[30] == pytest.approx(0.1 * (1 - 1 / 4) ** 0.9) and seen[-1] == pytest.approx(0.0)
    local = LocalPolyLRScheduler(opt, 0.1, 10)
    local.step()
    local.step()
    assert opt.param_groups[0]["lr"] == pytest.approx(0.1 * (1 - 2 / 10) ** 0.9)  # the base class steps once at construction
    wrapper = NnUNetDataLoaderWrapper(ToyAugmenter(0, True), "2d", set_len=3)
    batches = list(wrapper)
    assert len(batches) == 3 and set(batches[0][1]) == {"0-16x16", "1-8x8"} and len(list(wrapper)) == 3


def test_flexible_nnunet_client_trains_like_the_plain_one() -> None:
    from fl4health_b200.clients.flexible.nnunet import FlexibleNnunetClient

    def run(cls):
        set_all_random_seeds(8)
        clients = [cls(torch.device("cpu"), 999, fold=0, backend=ToyBackend(i), client_name=f"n{i}", verbose=False) for i in range(2)]
        strategy = BasicFedAvg(min_fit_clients=2, min_evaluate_clients=2, min_available_clients=2, on_fit_config_fn=_cfg,
                               on_evaluate_config_fn=_cfg, fit_metrics_aggregation_fn=fit_metrics_aggregation_fn,
                               evaluate_metrics_aggregation_fn=evaluate_metrics_aggregation_fn)
        server = NnunetServer(SimpleClientManager(), {"n_server_rounds": 2, "nnunet_config": "2d"}, _cfg, strategy)
        return [v for _, v in run_simulation(server, clients, 2).losses_distributed]

    plain, flexible = run(NnunetClient), run(FlexibleNnunetClient)
    assert all(a == pytest.approx(b, rel=1e-5) for a, b in zip(plain, flexible))


def test_initial_parameter_request_after_a_properties_poll_is_unpacked() -> None:
    """Plan negotiation sets the polled client up; if the server then asks THE SAME client for the initial parameters
    (round 0), a personalised client must answer with the plain model state, not with its packed (weights + loss) payload
    — otherwise the strategy appends the drift weight a second time and every client fails to unpack."""
    from fl4health_b200.clients.flexible.nnunet import FlexibleNnunetClient
    from fl4health_b200.mixins.personalized import PersonalizedMode, make_it_personal

    cls = make_it_personal(FlexibleNnunetClient, PersonalizedMode.DITTO)
    client = cls(torch.device("cpu"), 999, fold=0, backend=ToyBackend(0), verbose=False)
    config = {**_cfg(0), "nnunet_plans": pickle.dumps(ToyBackend(0).plan())}
    client.get_properties(dict(config))
    assert client.initialized
    n_state = len(client.model.state_dict())
    assert len(client.get_parameters(dict(config))) == n_state  # round 0 = initial-parameter request
    assert len(client.get_parameters({**config, "current_server_round": 1})) == n_state + 1  # regular payload: + loss
