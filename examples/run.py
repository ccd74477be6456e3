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
        if per_rank == 1:
            build_spmd_federation(ctx, server, clients[ctx.rank])
        else:  # more clients than GPUs: every rank hosts a contiguous block of them
            from fl4health_b200.parallel.spmd_multi import build_spmd_federation_multi

            build_spmd_federation_multi(ctx, server, clients[ctx.rank * per_rank : (ctx.rank + 1) * per_rank])
        history, _ = server.fit(num_rounds=config["n_server_rounds"])
        is_reporting_rank = ctx.rank == 0  # every rank holds the same history: one summary line is enough
        ctx.shutdown()
    else:
        server, clients = SCENARIOS[args.scenario](config, torch.device(args.device))
        history = run_simulation(server, clients, config["n_server_rounds"])
    summary = {"scenario": args.scenario, "rounds": config["n_server_rounds"], "losses": history.losses_distributed,
               "metrics": {k: v[-1][1] for k, v in history.metrics_distributed.items()}}
    if not args.spmd or is_reporting_rank:
        print(json.dumps(summary, default=float))
    return summary


if __name__ == "__main__":
    main(sys.argv[1:])
