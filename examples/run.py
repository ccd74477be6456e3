"""``python -m examples.run <scenario>``: run one example scenario in-process (or one client per GPU with ``--spmd``
under ``torch.distributed.run``)."""

from __future__ import annotations

import argparse
import json
import sys

import torch

from examples.common import describe, load_example_config
from examples.scenarios import SCENARIOS
import examples.extra_scenarios  # noqa: F401  (registers the remaining scenarios)
from fl4health_b200.simulation import run_simulation
from fl4health_b200.utils.random import set_all_random_seeds


def _replicated(clients: list, ctx, ranks_per_client: int):  # noqa: ANN001, ANN202
    """This rank's replica of client ``rank // G``: same client index (hence the same data), a 1/G shard of it, and
    gradient averaging inside the group (``fl4health_b200.parallel.client_group``)."""
    from fl4health_b200.engine.data import BatchedTensorLoader
    from fl4health_b200.parallel.client_group import ClientGroup, ReplicatedClientMixin, shard_dataset

    group = ClientGroup.from_world(ctx.rank, ctx.world_size, ranks_per_client)
    client = clients[ctx.rank]
    client.__class__ = type(f"Replicated{type(client).__name__}", (ReplicatedClientMixin, type(client)), {})
    client.client_group, client.client_index = group, group.client_index
    client.client_name = f"{client.client_name}.replica{group.group_rank}"
    make_loaders = client.get_data_loaders

    def sharded_loaders(config):  # noqa: ANN001, ANN202
        train, val = make_loaders(config)
        shard = lambda loader, shuffle: BatchedTensorLoader(  # noqa: E731
            shard_dataset(loader.dataset, group, seed=11), max(loader.batch_size // group.group_size, 1), shuffle=shuffle,
            placement="device" if ctx.device.type == "cuda" else "host", device=ctx.device)
        return shard(train, True), shard(val, False)

    client.get_data_loaders = sharded_loaders
    return client


def _attach_json_reporters(server, clients: list, folder: str) -> None:  # noqa: ANN001
    """The reporting path of the reference's smoke tests: per-round server / client payloads as JSON files."""
    from fl4health_b200.reporting.json_reporter import JsonReporter

    server.reports_manager.reporters.append(JsonReporter(run_id="server", output_folder=folder))
    for index, client in enumerate(clients):
        client.reports_manager.reporters.append(JsonReporter(run_id=f"client_{index}", output_folder=folder))


def main(argv: list[str] | None = None) -> dict:
    parser = argparse.ArgumentParser(description=__doc__)
    parser.add_argument("scenario", choices=sorted(SCENARIOS), nargs="?")
    parser.add_argument("--list", action="store_true", help="list the available scenarios")
    parser.add_argument("--config", default=None, help="YAML config (defaults to examples/configs/<scenario>.yaml)")
    parser.add_argument("--rounds", type=int, default=None)
    parser.add_argument("--clients", type=int, default=None)
    parser.add_argument("--device", default="cuda" if torch.cuda.is_available() else "cpu")
    parser.add_argument("--spmd", action="store_true", help="one client per rank (launch with torch.distributed.run)")
    parser.add_argument("--clients-per-rank", type=int, default=1, help="with --spmd: host this many clients on every rank")
    parser.add_argument("--ranks-per-client", type=int, default=1,
                        help="with --spmd: every client spans this many GPUs (replicas on disjoint data shards, gradients averaged per step)")
    parser.add_argument("--metrics-dir", default=None,
                        help="attach a JsonReporter to the server and every client; writes server.json / client_<i>.json there")
    args = parser.parse_args(argv)
    if args.list or args.scenario is None:
        print("\n".join(sorted(SCENARIOS)))
        return {}
    config = load_example_config(args.scenario, args.config, {"n_server_rounds": args.rounds, "n_clients": args.clients})
    set_all_random_seeds(config["seed"])
    describe(args.scenario, config)
    if args.spmd:
        from fl4health_b200.parallel.spmd import SpmdContext, build_spmd_federation

        ctx = SpmdContext()
        per_rank = max(1, args.clients_per_rank)
        config["n_clients"] = ctx.world_size * per_rank
        server, clients = SCENARIOS[args.scenario](config, ctx.device)
        if args.ranks_per_client > 1:
            assert per_rank == 1, "--ranks-per-client and --clients-per-rank are mutually exclusive"
            from fl4health_b200.parallel.client_group import require_replica_safe_strategy

            require_replica_safe_strategy(server.strategy, args.ranks_per_client)
            build_spmd_federation(ctx, server, _replicated(clients, ctx, args.ranks_per_client))
        elif per_rank == 1:
            build_spmd_federation(ctx, server, clients[ctx.rank])
        else:  # more clients than GPUs: every rank hosts a contiguous block of them
            from fl4health_b200.parallel.spmd_multi import build_spmd_federation_multi

            build_spmd_federation_multi(ctx, server, clients[ctx.rank * per_rank : (ctx.rank + 1) * per_rank])
        history, _ = server.fit(num_rounds=config["n_server_rounds"])
        is_reporting_rank = ctx.rank == 0  # every rank holds the same history: one summary line is enough
        ctx.shutdown()
    else:
        server, clients = SCENARIOS[args.scenario](config, torch.device(args.device))
        if args.metrics_dir is not None:
            _attach_json_reporters(server, clients, args.metrics_dir)
        history = run_simulation(server, clients, config["n_server_rounds"])
        if args.metrics_dir is not None:
            for host in (server, *clients):
                host.reports_manager.shutdown()
    summary = {"scenario": args.scenario, "rounds": config["n_server_rounds"], "losses": history.losses_distributed,
               "metrics": {k: v[-1][1] for k, v in history.metrics_distributed.items()}}
    if not args.spmd or is_reporting_rank:
        print(json.dumps(summary, default=float))
    return summary


if __name__ == "__main__":
    main(sys.argv[1:])
