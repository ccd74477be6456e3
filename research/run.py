"""Command line for the research harness.

    python -m research.run train --task cifar10 --method ditto --lr 0.01 --lam 1.0 --runs 3
    python -m research.run sweep --task cifar10 --method mr_mtl --grid lr=0.001,0.01,0.1 --grid lam=0.1,1.0
    python -m research.run best  --dir research_out/cifar10/mr_mtl
    python -m research.run test  --dir research_out/cifar10/mr_mtl/lr_0.01_lam_1.0
    torchrun --nproc-per-node 5 --master-addr 127.0.0.1 -m research.run train --task cifar10 --method fedavg --spmd
"""

from __future__ import annotations

import argparse
import json
import os
from dataclasses import fields
from typing import Any

import torch

from research.harness.experiment import ExperimentSpec, run_experiment, sweep
from research.harness.selection import evaluate_on_test, find_best_hp


def _coerce(text: str) -> Any:
    for cast in (int, float):
        try:
            return cast(text)
        except ValueError:
            pass
    return {"true": True, "false": False, "none": None}.get(text.lower(), text)


def _spec_from_args(args: argparse.Namespace) -> ExperimentSpec:
    names = {f.name for f in fields(ExperimentSpec)}
    values = {k: v for k, v in vars(args).items() if k in names and v is not None}
    if args.task_kwargs:
        values["task_kwargs"] = {k: _coerce(v) for k, v in (item.split("=", 1) for item in args.task_kwargs)}
    return ExperimentSpec(**values)


def main(argv: list[str] | None = None) -> Any:
    parser = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    sub = parser.add_subparsers(dest="command", required=True)
    for name in ("train", "sweep"):
        p = sub.add_parser(name)
        for f in fields(ExperimentSpec):
            if f.name == "task_kwargs":
                continue
            if f.type in ("bool", bool):
                p.add_argument(f"--{f.name.replace('_', '-')}", dest=f.name, action=argparse.BooleanOptionalAction, default=None)
            else:
                p.add_argument(f"--{f.name.replace('_', '-')}", dest=f.name, type=_coerce, default=None)
        p.add_argument("--task-kwargs", nargs="*", default=None, help="key=value arguments of the task builder, e.g. n_clients=8")
        p.add_argument("--device", default=None)
        p.add_argument("--spmd", action="store_true", help="run under torchrun: one slice of the clients per rank")
        if name == "sweep":
            p.add_argument("--grid", action="append", default=[], help="name=v1,v2,... (repeatable)")
    for name in ("best", "test"):
        p = sub.add_parser(name)
        p.add_argument("--dir", required=True)
        if name == "test":
            p.add_argument("--which", default="best", choices=("best", "last"))
    args = parser.parse_args(argv)

    if args.command == "best":
        folder, loss = find_best_hp(args.dir)
        print(json.dumps({"best_folder": str(folder), "best_loss": loss}))
        return folder, loss
    if args.command == "test":
        report = evaluate_on_test(args.dir, which=args.which)
        print(json.dumps(report))
        return report
    spec = _spec_from_args(args)
    device = torch.device(args.device) if args.device else None
    if args.command == "train":
        results = run_experiment(spec, device, args.spmd)
        if int(os.environ.get("RANK", "0")) == 0:  # under torchrun every rank holds the same results
            print(json.dumps([{k: r[k] for k in ("best_aggregated_loss", "seconds", "rounds_per_s")} for r in results]))
        return results
    grid = {k: [_coerce(v) for v in vs.split(",")] for k, vs in (item.split("=", 1) for item in args.grid)}
    results = sweep(spec, grid, device, args.spmd)
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps({key: [r["best_aggregated_loss"] for r in runs] for key, runs in results.items()}))
    return results


if __name__ == "__main__":
    main()
