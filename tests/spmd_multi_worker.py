"""Worker for the multi-client-per-rank SPMD test (torch.distributed.run, gloo on CPU)."""

import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
os.environ.setdefault("FL4H_LOG_LEVEL", "WARNING")

import torch  # noqa: E402

from fl4health_b200.metrics.metric_aggregation import evaluate_metrics_aggregation_fn, fit_metrics_aggregation_fn  # noqa: E402
from fl4health_b200.parallel.spmd import SpmdContext  # noqa: E402
from fl4health_b200.parallel.spmd_multi import build_spmd_federation_multi  # noqa: E402
from fl4health_b200.servers.base_server import FlServer  # noqa: E402
from fl4health_b200.servers.client_manager import SimpleClientManager  # noqa: E402
from fl4health_b200.strategies.basic_fedavg import BasicFedAvg  # noqa: E402
from fl4health_b200.strategies.fedopt import FedAdam  # noqa: E402
from fl4health_b200.utils.random import set_all_random_seeds  # noqa: E402
from tests.helpers import fit_config_fn, make_clients  # noqa: E402


def main() -> None:
    out_path, strategy_name, split = sys.argv[1], sys.argv[2], [int(v) for v in sys.argv[3].split(",")]
    ctx = SpmdContext()
    set_all_random_seeds(42)
    total = sum(split)
    fraction = float(os.environ.get("FL4H_TEST_FRACTION", "1.0"))
    sampled = max(1, int(total * fraction))
    common = dict(fraction_fit=fraction, fraction_evaluate=fraction, min_fit_clients=sampled, min_evaluate_clients=sampled,
                  min_available_clients=total,
                  on_fit_config_fn=fit_config_fn(), on_evaluate_config_fn=fit_config_fn(),
                  fit_metrics_aggregation_fn=fit_metrics_aggregation_fn,
                  evaluate_metrics_aggregation_fn=evaluate_metrics_aggregation_fn)
    start = sum(split[: ctx.rank])
    clients = make_clients(total, device=str(ctx.device))[start : start + split[ctx.rank]]
    if strategy_name == "fedadam":
        from fl4health_b200.common.typing import ndarrays_to_parameters
        from fl4health_b200.models import Net
        from fl4health_b200.parallel.arena import attach_arena

        torch.manual_seed(1234)
        template = Net().to(ctx.device)
        arena = attach_arena(template, ctx.device, with_grad=False)
        strategy = FedAdam(initial_parameters=ndarrays_to_parameters(arena.ndarrays()), eta=0.05, **common)
    else:
        strategy = BasicFedAvg(**common)
    rounds = int(os.environ.get("FL4H_TEST_ROUNDS", "2"))
    server = FlServer(SimpleClientManager(), {"n_server_rounds": rounds}, strategy, on_init_parameters_config_fn=fit_config_fn())
    proxies = build_spmd_federation_multi(ctx, server, clients)
    history, _ = server.fit(num_rounds=rounds)
    if ctx.rank == 0:
        trained = next((c for c in clients if getattr(c, "initialized", False)), None)
        state = {} if trained is None else {k: v.detach().cpu().double().sum().item() for k, v in trained.model.state_dict().items()}
        Path(out_path).write_text(json.dumps({"losses": history.losses_distributed, "state": state, "clients": len(proxies)}))
    ctx.barrier()
    ctx.shutdown()


if __name__ == "__main__":
    main()
