"""Per-site RxRx1 tensors for the ``rxrx1`` task.

Input is the directory produced by ``fl4health_b200.datasets.rxrx1.preprocess`` (``clients/meta_data_<k>.csv`` plus the
pickled image tensors); each client's training images are split train / val with ``create_splits`` and its test images
are kept apart (same protocol as ``research/rxrx1/data/data_utils.py`` in the reference).  Output::

    <out>/rxrx1/client_<i>_{train,val,test}.pt

    python -m research.rxrx1.preprocess --dataset-dir datasets/rxrx1 --out research_data --clients 4
"""

from __future__ import annotations

import argparse
from pathlib import Path

import pandas as pd
import torch

from fl4health_b200.datasets.rxrx1.load_data import construct_rxrx1_tensor_dataset, create_splits


def main(argv: list[str] | None = None) -> None:
    parser = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    parser.add_argument("--dataset-dir", type=Path, required=True)
    parser.add_argument("--out", type=Path, default=Path("research_data"))
    parser.add_argument("--clients", type=int, default=4)
    parser.add_argument("--seed", type=int, default=2021)
    args = parser.parse_args(argv)
    out = args.out / "rxrx1"
    out.mkdir(parents=True, exist_ok=True)
    for client in range(args.clients):
        metadata = pd.read_csv(args.dataset_dir / "clients" / f"meta_data_{client + 1}.csv")
        train_full, _ = construct_rxrx1_tensor_dataset(metadata, args.dataset_dir, client, "train")
        test, _ = construct_rxrx1_tensor_dataset(metadata, args.dataset_dir, client, "test")
        train_idx, val_idx = create_splits(train_full, seed=args.seed)
        for name, data, targets in (("train", train_full.data[train_idx], train_full.targets[train_idx]),
                                    ("val", train_full.data[val_idx], train_full.targets[val_idx]), ("test", test.data, test.targets)):
            torch.save({"data": data.float(), "targets": targets.long()}, out / f"client_{client}_{name}.pt")


if __name__ == "__main__":
    main()
