"""Samplers, Dirichlet partitioner, synthetic FedProx data, dataset converters, raw MNIST/CIFAR readers."""

import gzip
import pickle
import struct
from pathlib import Path

import numpy as np
import pytest
import torch
from torch import nn

from fl4health_b200.utils.data_generation import SyntheticIidFedProxDataset, SyntheticNonIidFedProxDataset
from fl4health_b200.utils.dataset import TensorDataset
from fl4health_b200.utils.dataset_converter import AutoEncoderDatasetConverter
from fl4health_b200.utils.load_data import load_cifar10_data, load_mnist_data, load_mnist_test_data, split_data_and_targets
from fl4health_b200.utils.msd_dataset_sources import MsdDataset, get_msd_dataset_enum, msd_num_labels, msd_urls
from fl4health_b200.utils.parameter_extraction import check_shape_match, get_all_model_parameters
from fl4health_b200.utils.partitioners import DirichletLabelBasedAllocation
from fl4health_b200.utils.peft_parameter_extraction import get_all_peft_parameters_from_model
from fl4health_b200.utils.sampler import DirichletLabelBasedSampler, MinorityLabelBasedSampler


def _dataset(n_per_class: int = 100, classes: int = 5) -> TensorDataset:
    targets = torch.arange(classes).repeat_interleave(n_per_class)
    return TensorDataset(torch.randn(len(targets), 3), targets)


def test_minority_sampler() -> None:
    ds = MinorityLabelBasedSampler(list(range(5)), 0.1, {1, 3}).subsample(_dataset())
    counts = torch.bincount(ds.targets, minlength=5).tolist()
    assert counts == [100, 10, 100, 10, 100] and len(ds) == 320


def test_dirichlet_sampler_is_seeded_and_sized() -> None:
    a = DirichletLabelBasedSampler(list(range(5)), hash_key=7, sample_percentage=0.5, beta=0.5)
    b = DirichletLabelBasedSampler(list(range(5)), hash_key=7, sample_percentage=0.5, beta=0.5)
    assert np.allclose(a.probabilities, b.probabilities) and a.probabilities.sum() == pytest.approx(1.0)
    da, db = a.subsample(_dataset()), b.subsample(_dataset())
    assert len(da) == 250 and torch.equal(da.targets, db.targets)
    freq = torch.bincount(da.targets, minlength=5).float() / 250
    assert np.allclose(freq.numpy(), a.probabilities, atol=0.02)


def test_dirichlet_partitioner() -> None:
    torch.manual_seed(0)
    np.random.seed(0)
    ds = _dataset(200, 4)
    parts, probs = DirichletLabelBasedAllocation(3, list(range(4)), min_label_examples=2, beta=5.0).partition_dataset(ds, max_retries=50)
    assert len(parts) == 3 and set(probs) == {0, 1, 2, 3}
    total = sum(len(p) for p in parts)
    assert 800 - 4 * 3 <= total <= 800  # only rounding remainders are dropped
    rows = torch.cat([p.data for p in parts])
    assert len(torch.unique(rows, dim=0)) == total  # disjoint shards
    prior = {label: np.array([0.5, 0.25, 0.25]) for label in range(4)}
    parts, _ = DirichletLabelBasedAllocation(3, list(range(4)), prior_distribution=prior).partition_dataset(ds)
    assert [len(p) for p in parts] == [400, 200, 200]
    with pytest.raises(ValueError):
        DirichletLabelBasedAllocation(3, list(range(4)), min_label_examples=150, beta=1.0).partition_dataset(ds, max_retries=3)
    with pytest.raises(AssertionError):
        DirichletLabelBasedAllocation(3, [0], beta=1.0, prior_distribution=prior)


def test_synthetic_fedprox_data() -> None:
    torch.manual_seed(1)
    non_iid = SyntheticNonIidFedProxDataset(num_clients=3, alpha=0.5, beta=0.5, samples_per_client=200).generate()
    assert len(non_iid) == 3 and non_iid[0].data.shape == (200, 60) and non_iid[0].targets.shape == (200, 10)
    assert torch.all(non_iid[0].targets.sum(dim=1) == 1)
    var = non_iid[0].data.var(dim=0)
    assert var[0] > var[30] > var[59] * 0.5  # Sigma_jj = j^-1.2 decays
    two_layer = SyntheticNonIidFedProxDataset(2, 1.0, 1.0, hidden_dim=16, samples_per_client=50).generate()
    assert two_layer[1].targets.shape == (50, 10)
    iid = SyntheticIidFedProxDataset(num_clients=2, samples_per_client=100)
    a, b = iid.generate()
    assert a.data.shape == (100, 60) and not torch.equal(a.data, b.data)
    # the labelling function is shared across IID clients
    assert torch.equal(iid.one_layer_map_inputs_to_outputs(a.data, iid.w, iid.b), a.targets)


def test_autoencoder_converters() -> None:
    ds = TensorDataset(torch.randn(10, 1, 4, 4), torch.arange(10) % 3)
    plain = AutoEncoderDatasetConverter().convert_dataset(ds)
    x, y = plain[2]
    assert torch.equal(x, y) and x.shape == (1, 4, 4) and len(plain) == 10
    labelled = AutoEncoderDatasetConverter(condition="label", do_one_hot_encoding=True)
    labelled.convert_dataset(ds)
    x, y = labelled[4]
    assert x.shape == (16 + 3,) and y.shape == (1, 4, 4) and labelled.get_condition_vector_size() == 3
    batch_x, batch_y = labelled.get_batch(torch.tensor([0, 1, 2]))
    data, cond = labelled.get_unpacking_function()(batch_x)
    assert data.shape == (3, 1, 4, 4) and torch.equal(data, batch_y) and torch.equal(cond.argmax(1), ds.targets[:3])
    fixed = AutoEncoderDatasetConverter(condition=torch.tensor([1.0, 0.0]))
    fixed.convert_dataset(ds)
    assert fixed[0][0].shape == (18,) and fixed.get_condition_vector_size() == 2


def _write_idx(path: Path, array: np.ndarray, gz: bool) -> None:
    header = struct.pack(">BBBB", 0, 0, 0x08, array.ndim) + struct.pack(">" + "I" * array.ndim, *array.shape)
    opener = gzip.open if gz else open
    with opener(path, "wb") as handle:
        handle.write(header + array.astype(np.uint8).tobytes())


def test_mnist_and_cifar_raw_readers(tmp_path: Path) -> None:
    raw = tmp_path / "MNIST" / "raw"
    raw.mkdir(parents=True)
    rng = np.random.default_rng(0)
    images, labels = rng.integers(0, 256, (50, 28, 28)), rng.integers(0, 10, (50,))
    _write_idx(raw / "train-images-idx3-ubyte.gz", images, True)
    _write_idx(raw / "train-labels-idx1-ubyte.gz", labels, True)
    _write_idx(raw / "t10k-images-idx3-ubyte", images[:10], False)
    _write_idx(raw / "t10k-labels-idx1-ubyte", labels[:10], False)
    train, val, counts = load_mnist_data(tmp_path, batch_size=8, hash_key=3)
    assert counts == {"train_set": 40, "validation_set": 10}
    x, y = next(iter(train))
    assert x.shape == (8, 1, 28, 28) and x.min() >= -1 and x.max() <= 1 and y.dtype == torch.int64
    test, test_counts = load_mnist_test_data(tmp_path, batch_size=4)
    assert test_counts == {"eval_set": 10} and next(iter(test))[0].shape == (4, 1, 28, 28)
    again, _, _ = load_mnist_data(tmp_path, batch_size=8, hash_key=3)
    assert torch.equal(again.dataset.targets, train.dataset.targets)  # split is deterministic per hash key
    cifar = tmp_path / "cifar-10-batches-py"
    cifar.mkdir()
    for i in range(1, 6):
        with open(cifar / f"data_batch_{i}", "wb") as handle:
            pickle.dump({"data": rng.integers(0, 256, (20, 3072)).astype(np.uint8), "labels": rng.integers(0, 10, 20).tolist()}, handle)
    train, val, counts = load_cifar10_data(tmp_path, batch_size=16)
    assert counts == {"train_set": 80, "validation_set": 20} and next(iter(train))[0].shape == (16, 3, 32, 32)
    with pytest.raises(FileNotFoundError):
        load_mnist_data(tmp_path / "nowhere", 4)


def test_small_helpers() -> None:
    a, b, c, d = split_data_and_targets(torch.arange(20).float().reshape(10, 2), torch.arange(10), 0.3, hash_key=1)
    assert len(a) == 7 and len(c) == 3 and set(b.tolist()) | set(d.tolist()) == set(range(10))
    assert get_msd_dataset_enum("Task04_Hippocampus") is MsdDataset.TASK04_HIPPOCAMPUS
    assert msd_num_labels[MsdDataset.TASK01_BRAINTUMOUR] == 4 and msd_urls[MsdDataset.TASK09_SPLEEN].endswith("Task09_Spleen.tar")
    model = nn.Linear(3, 2)
    assert len(get_all_model_parameters(model).tensors) == 2
    check_shape_match(model.parameters(), nn.Linear(3, 2).parameters(), "mismatch")
    with pytest.raises(AssertionError):
        check_shape_match(model.parameters(), nn.Linear(4, 2).parameters(), "mismatch")

    class WithLora(nn.Module):
        def __init__(self) -> None:
            super().__init__()
            self.base = nn.Linear(4, 4)
            self.lora_A = nn.Linear(4, 2, bias=False)
            self.lora_B = nn.Linear(2, 4, bias=False)

    assert len(get_all_peft_parameters_from_model(WithLora()).tensors) == 2
