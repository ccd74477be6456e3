"""Dirichlet(β) label-skew partition of CIFAR-10 into per-client train / val / test tensors.

The training split is partitioned with a freshly drawn Dirichlet allocation; the validation and test splits are then
partitioned with the *training* allocation as a fixed prior so that every client's three splits share one label
distribution (same protocol as ``research/cifar10/preprocess.py:85-121``).  Output (read by the ``cifar10`` task)::

    <out>/beta_<β>/seed_<s>/client_<i>_{train,val,test}.pt     # {"data": float32 [N,3,32,32] in [-1,1], "targets": int64 [N]}

    python -m research.cifar10.preprocess --dataset-dir datasets/cifar10 --out research_data --clients 5 --beta 0.5
"""

from __future__ import annotations

import argparse
from logging import INFO
from pathlib import Path

import torch

from fl4health_b200.common.logger import log
from fl4health_b200.utils.dataset import TensorDataset
from fl4health_b200.utils.partitioners import DirichletLabelBasedAllocation


def to_model_input(images_hwc_uint8: torch.Tensor) -> torch.Tensor:
    """uint8 ``[N,32,32,3]`` -> float32 ``[N,3,32,32]`` in ``[-1, 1]`` (ToTensor + Normalize(0.5, 0.5))."""
    return images_hwc_uint8.permute(0, 3, 1, 2).float().div_(255.0).sub_(0.5).div_(0.5).contiguous()


def partition(train: TensorDataset, val: TensorDataset, test: TensorDataset, n_clients: int, beta: float, seed: int,
              class_num: int = 10) -> list[tuple[TensorDataset, TensorDataset, TensorDataset]]:
    torch.manual_seed(seed)
    import numpy as np

    np.random.seed(seed)
    labels = list(range(class_num))
    fresh = DirichletLabelBasedAllocation(number_of_partitions=n_clients, unique_labels=labels, beta=beta, min_label_examples=1)
    train_parts, train_distribution = fresh.partition_dataset(train, max_retries=None)
    with_prior = DirichletLabelBasedAllocation(number_of_partitions=n_clients, unique_labels=labels, prior_distribution=train_distribution)
    val_parts, _ = with_prior.partition_dataset(val, max_retries=None)
    test_parts, _ = with_prior.partition_dataset(test, max_retries=None)
    return list(zip(train_parts, val_parts, test_parts))


def save_partitions(parts: list[tuple[TensorDataset, TensorDataset, TensorDataset]], out_dir: Path) -> None:
    out_dir.mkdir(parents=True, exist_ok=True)
    for index, triple in enumerate(parts):
        for name, dataset in zip(("train", "val", "test"), triple):
            torch.save({"data": dataset.data, "targets": dataset.targets}, out_dir / f"client_{index}_{name}.pt")
            log(INFO, f"client {index} {name}: {len(dataset.data)} samples, label histogram "
                      f"{torch.bincount(dataset.targets, minlength=10).tolist()}")


def main(argv: list[str] | None = None) -> None:
    from fl4health_b200.utils.load_data import get_cifar10_data_and_target_tensors

    parser = argparse.ArgumentParser(description="Dirichlet partition of CIFAR-10 for the pFL benchmark")
    parser.add_argument("--dataset-dir", type=Path, required=True, help="directory holding cifar-10-batches-py")
    parser.add_argument("--out", type=Path, default=Path("research_data"))
    parser.add_argument("--clients", type=int, default=5)
    parser.add_argument("--beta", type=float, default=0.5)
    parser.add_argument("--seed", type=int, default=2021)
    parser.add_argument("--val-fraction", type=float, default=0.2)
    args = parser.parse_args(argv)

    images, targets = get_cifar10_data_and_target_tensors(args.dataset_dir, train=True)
    test_images, test_targets = get_cifar10_data_and_target_tensors(args.dataset_dir, train=False)
    order = torch.randperm(len(images), generator=torch.Generator().manual_seed(args.seed))
    n_val = int(len(order) * args.val_fraction)
    data = to_model_input(images)
    train = TensorDataset(data[order[n_val:]], targets[order[n_val:]])
    val = TensorDataset(data[order[:n_val]], targets[order[:n_val]])
    test = TensorDataset(to_model_input(test_images), test_targets)
    save_partitions(partition(train, val, test, args.clients, args.beta, args.seed), args.out / f"beta_{args.beta}" / f"seed_{args.seed}")


if __name__ == "__main__":
    main()
