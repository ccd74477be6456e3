"""Export a FLamby dataset's per-centre splits to the tensor files the harness tasks read.

FLamby is an optional, separately licensed dependency: this script imports it lazily and is the only place that does.
``python -m research.flamby.export --dataset fed_heart_disease --out research_data`` writes
``<out>/fed_heart_disease/client_<i>_{train,val,test}.pt`` (validation = a seeded 20 % of each centre's training
pool, as the reference's FLamby clients do: ``research/flamby/flamby_data_utils.py``)."""

from __future__ import annotations

import argparse
import importlib
from pathlib import Path

import torch

_DATASETS = {"fed_heart_disease": ("flamby.datasets.fed_heart_disease", "FedHeartDisease", 4),
             "fed_isic2019": ("flamby.datasets.fed_isic2019", "FedIsic2019", 6), "fed_ixi": ("flamby.datasets.fed_ixi", "FedIXITiny", 3)}


def _stack(dataset: object) -> tuple[torch.Tensor, torch.Tensor]:
    xs, ys = zip(*[dataset[i] for i in range(len(dataset))])  # type: ignore[index,arg-type]
    return torch.stack([torch.as_tensor(x).float() for x in xs]), torch.stack([torch.as_tensor(y) for y in ys]).long().flatten()


def main(argv: list[str] | None = None) -> None:
    parser = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    parser.add_argument("--dataset", choices=sorted(_DATASETS), required=True)
    parser.add_argument("--out", type=Path, default=Path("research_data"))
    parser.add_argument("--seed", type=int, default=2021)
    args = parser.parse_args(argv)
    module_name, class_name, centres = _DATASETS[args.dataset]
    try:
        cls = getattr(importlib.import_module(module_name), class_name)
    except ImportError as exc:
        raise SystemExit(f"FLamby is not installed ({exc}); the tasks fall back to synthetic records without it") from exc
    out = args.out / args.dataset
    out.mkdir(parents=True, exist_ok=True)
    for centre in range(centres):
        pool_x, pool_y = _stack(cls(center=centre, train=True, pooled=False))
        test_x, test_y = _stack(cls(center=centre, train=False, pooled=False))
        order = torch.randperm(len(pool_x), generator=torch.Generator().manual_seed(args.seed + centre))
        n_val = max(int(0.2 * len(order)), 1)
        for name, x, y in (("train", pool_x[order[n_val:]], pool_y[order[n_val:]]), ("val", pool_x[order[:n_val]], pool_y[order[:n_val]]),
                           ("test", test_x, test_y)):
            torch.save({"data": x, "targets": y}, out / f"client_{centre}_{name}.pt")


if __name__ == "__main__":
    main()
