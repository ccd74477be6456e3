"""Worker for the multi-GPU-client test (torch.distributed.run, gloo on CPU): W ranks, groups of G ranks per client."""

import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
os.environ.setdefault("FL4H_LOG_LEVEL", "WARNING")

import torch  # noqa: E402

from fl4health_b200.clients.basic_client import BasicClient  # noqa: E402
from fl4health_b200.engine.data import BatchedTensorLoader  # noqa: E402
from fl4health_b200.metrics import Accuracy  # noqa: E402
from fl4health_b200.metrics.metric_aggregation import evaluate_metrics_aggregation_fn, fit_metrics_aggregation_fn  # noqa: E402
from fl4health_b200.parallel.client_group import ClientGroup, ReplicatedClientMixin, Zero1ClientMixin, shard_dataset  # noqa: E402
from fl4health_b200.parallel.spmd import SpmdContext, build_spmd_federation  # noqa: E402
from fl4health_b200.servers.base_server import FlServer  # noqa: E402
from fl4health_b200.servers.client_manager import SimpleClientManager  # noqa: E402
from fl4health_b200.strategies.basic_fedavg import BasicFedAvg  # noqa: E402
from tests.helpers import synthetic_cifar  # noqa: E402


class GroupNormNet(torch.nn.Module):
    """No batch statistics: G replicas on half batches then compute exactly the full-batch gradient."""

    def __init__(self) -> None:
        super().__init__()
        self.conv = torch.nn.Conv2d(3, 4, 3, padding=1)
        self.norm = torch.nn.GroupNorm(2, 4)
        self.fc = torch.nn.Linear(4 * 8 * 8, 10)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = torch.nn.functional.adaptive_avg_pool2d(torch.relu(self.norm(self.conv(x))), 8)
        return self.fc(torch.flatten(x, 1))


class _Hooks:
    data_seed = 0

    def get_model(self, config):  # noqa: ANN001, ANN201
        torch.manual_seed(99)
        return GroupNormNet()

    def get_data_loaders(self, config):  # noqa: ANN001, ANN201
        full_train, full_val = synthetic_cifar(128, self.data_seed), synthetic_cifar(32, 10_000 + self.data_seed)
        train, val = shard_dataset(full_train, self.client_group, seed=5), shard_dataset(full_val, self.client_group, seed=6)
        per_replica = int(config["batch_size"]) // self.client_group.group_size
        return (BatchedTensorLoader(train, per_replica, shuffle=True, generator=torch.Generator().manual_seed(3)),
                BatchedTensorLoader(val, per_replica))

    def get_criterion(self, config):  # noqa: ANN001, ANN201
        return torch.nn.CrossEntropyLoss()

    def get_optimizer(self, config):  # noqa: ANN001, ANN201
        return torch.optim.SGD(self.model.parameters(), lr=0.05, momentum=0.9)


class ShardedClient(ReplicatedClientMixin, _Hooks, BasicClient):
    pass


class Zero1Client(Zero1ClientMixin, _Hooks, BasicClient):
    """Same hooks, optimizer state sharded over the replicas."""


def _twin_optimizers(self, config):  # noqa: ANN001, ANN202
    return {"global": torch.optim.SGD(self.global_model.parameters(), lr=0.05, momentum=0.9),
            "local": torch.optim.SGD(self.model.parameters(), lr=0.05, momentum=0.9)}


def _client_class() -> type:
    if os.environ.get("FL4H_TEST_ZERO1") == "1":
        return Zero1Client
    variant = os.environ.get("FL4H_TEST_VARIANT", "basic")
    if variant == "ditto":  # own train_step, two models, two optimizers
        from fl4health_b200.clients.ditto_client import DittoClient

        return type("ShardedDitto", (ReplicatedClientMixin, _Hooks, DittoClient), {"get_optimizer": _twin_optimizers})
    if variant == "apfl":
        from fl4health_b200.clients.apfl_client import ApflClient
        from fl4health_b200.model_bases.apfl_base import ApflModule

        def apfl_model(self, config):  # noqa: ANN001, ANN202
            torch.manual_seed(99)
            return ApflModule(GroupNormNet())

        def apfl_optimizers(self, config):  # noqa: ANN001, ANN202
            return {"global": torch.optim.SGD(self.model.global_model.parameters(), lr=0.05),
                    "local": torch.optim.SGD(self.model.local_model.parameters(), lr=0.05)}

        return type("ShardedApfl", (ReplicatedClientMixin, _Hooks, ApflClient), {"get_model": apfl_model, "get_optimizer": apfl_optimizers})
    return ShardedClient


def _all_parameters(client) -> torch.Tensor:  # noqa: ANN001
    """Every trainable tensor the client owns: the exchanged model and any personal / twin model."""
    modules = [client.model] + ([client.global_model] if isinstance(getattr(client, "global_model", None), torch.nn.Module) else [])
    return torch.cat([p.detach().double().flatten() for m in modules for p in m.parameters()])


def main() -> None:
    out_path, group_size = sys.argv[1], int(sys.argv[2])
    client_cls = _client_class()
    ctx = SpmdContext()
    group = ClientGroup.from_world(ctx.rank, ctx.world_size, group_size)

    def fn(server_round: int) -> dict:
        return {"current_server_round": server_round, "local_steps": 4, "batch_size": 16}

    n = ctx.world_size
    strategy_cls, extra = BasicFedAvg, {}
    if os.environ.get("FL4H_TEST_VARIANT") == "ditto":
        from fl4health_b200.strategies.fedavg_with_adaptive_constraint import FedAvgWithAdaptiveConstraint

        strategy_cls, extra = FedAvgWithAdaptiveConstraint, {"initial_parameters": None, "initial_loss_weight": 0.5}
    strategy = strategy_cls(**extra, min_fit_clients=n, min_evaluate_clients=n, min_available_clients=n, on_fit_config_fn=fn, on_evaluate_config_fn=fn,
                           fit_metrics_aggregation_fn=fit_metrics_aggregation_fn, evaluate_metrics_aggregation_fn=evaluate_metrics_aggregation_fn)
    server = FlServer(SimpleClientManager(), {"n_server_rounds": 2}, strategy, on_init_parameters_config_fn=fn)
    client = client_cls(Path("."), [Accuracy()], ctx.device, client_name=f"client{group.client_index}.replica{group.group_rank}")
    client.client_group, client.data_seed = group, group.client_index
    build_spmd_federation(ctx, server, client)

    # checkpoint of the replicas' agreement: parameters just before the aggregate of round 1
    seen = {}
    original_fit = client.fit

    def spying_fit(parameters, config):  # noqa: ANN001, ANN202
        result = original_fit(parameters, config)
        seen[int(config["current_server_round"])] = _all_parameters(client).clone()
        return result

    client.fit = spying_fit
    history, _ = server.fit(num_rounds=2)
    # unit-level oracle: averaged half-batch gradients == gradient of the mean loss over the union batch
    from fl4health_b200.parallel.client_group import average_gradients

    torch.manual_seed(7)
    probe, x, y = GroupNormNet().to(ctx.device), torch.randn(16, 3, 32, 32).to(ctx.device), torch.randint(0, 10, (16,)).to(ctx.device)
    reference = torch.autograd.grad(torch.nn.functional.cross_entropy(probe(x), y), list(probe.parameters()))
    mine = slice(group.group_rank, None, group.group_size)
    torch.nn.functional.cross_entropy(probe(x[mine]), y[mine]).backward()
    average_gradients(probe, group)
    grad_error = max(float((p.grad - r).abs().max()) for p, r in zip(probe.parameters(), reference))
    payload = {"rank": ctx.rank, "grad_error": grad_error, "client": group.client_index, "losses": history.losses_distributed,
               "pre_aggregate": {str(k): [float(v.sum()), float(v.abs().sum())] for k, v in seen.items()},
               "train_samples": client.num_train_samples, "optimizer_class": type(client.optimizers["global"]).__name__,
               "trainable_elements": sum(p.numel() for p in client.model.parameters() if p.requires_grad),
               "optimizer_state_elements": sum(v.numel() for state in client.optimizers["global"].state.values() for v in state.values()
                                               if isinstance(v, torch.Tensor)) if hasattr(client.optimizers["global"], "state") else -1,
               "final": [float(v) for v in torch.cat([p.detach().double().flatten() for p in client.model.parameters()])[:50]]}
    Path(f"{out_path}.rank{ctx.rank}").write_text(json.dumps(payload))
    ctx.barrier()
    ctx.shutdown()


if __name__ == "__main__":
    main()
