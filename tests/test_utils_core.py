"""Unit tests of the small utility modules (mirrors the reference's tests/utils/{config,random,functions,client}_test.py)."""

import random

import numpy as np
import pytest
import torch
from torch import nn

from fl4health_b200.utils import client as client_utils
from fl4health_b200.utils import config as config_utils
from fl4health_b200.utils import functions
from fl4health_b200.utils.random import generate_hash, restore_random_state, save_random_state, set_all_random_seeds, unset_all_random_seeds
from fl4health_b200.utils.logging import LoggingMode


# ---- config ------------------------------------------------------------------------------------------------------
def test_check_and_load_config(tmp_path) -> None:
    good = tmp_path / "good.yaml"
    good.write_text("n_server_rounds: 3\nbatch_size: 16\nlocal_epochs: 2\n")
    assert config_utils.load_config(str(good)) == {"n_server_rounds": 3, "batch_size": 16, "local_epochs": 2}
    for bad, message in (({"batch_size": 4}, "n_server_rounds must be specified"),
                         ({"n_server_rounds": 2.5, "batch_size": 4}, "n_server_rounds must be of type"),
                         ({"n_server_rounds": 0, "batch_size": 4}, "greater than 0"),
                         ({"n_server_rounds": 2, "batch_size": True}, "batch_size must be of type")):
        with pytest.raises(config_utils.InvalidConfigError, match=message):
            config_utils.check_config(bad)


def test_narrow_dict_type_and_helpers() -> None:
    config = {"a": 3, "b": "text"}
    assert config_utils.narrow_dict_type(config, "a", int) == 3
    with pytest.raises(ValueError, match="not present"):
        config_utils.narrow_dict_type(config, "missing", int)
    with pytest.raises(ValueError, match="correct type"):
        config_utils.narrow_dict_type(config, "b", int)

    class Holder:
        pass

    holder = Holder()
    config_utils.narrow_dict_type_and_set_attribute(holder, config, "a", "doubled", int, func=lambda v: 2 * v)
    assert holder.doubled == 6
    assert config_utils.make_dict_with_epochs_or_steps(local_epochs=2) == {"local_epochs": 2}
    assert config_utils.make_dict_with_epochs_or_steps(local_steps=7) == {"local_steps": 7}
    assert config_utils.make_dict_with_epochs_or_steps() == {}


# ---- random ------------------------------------------------------------------------------------------------------
def test_seeding_and_state_round_trip() -> None:
    set_all_random_seeds(123)
    first = (random.random(), float(np.random.rand()), float(torch.rand(1)))
    set_all_random_seeds(123)
    assert (random.random(), float(np.random.rand()), float(torch.rand(1))) == first
    state = save_random_state()
    after_save = (random.random(), float(np.random.rand()), float(torch.rand(1)))
    restore_random_state(*state)
    assert (random.random(), float(np.random.rand()), float(torch.rand(1))) == after_save
    unset_all_random_seeds()
    assert len(generate_hash()) == 8 and len(generate_hash(12)) == 12 and generate_hash() != generate_hash()


# ---- functions ---------------------------------------------------------------------------------------------------
def test_bernoulli_sample_gradient_and_sigmoid_inverse() -> None:
    probs = torch.full((1000,), 0.3, requires_grad=True)
    sample = functions.bernoulli_sample(probs)
    assert set(sample.detach().unique().tolist()) <= {0.0, 1.0} and 0.2 < float(sample.detach().mean()) < 0.4
    sample.sum().backward()
    assert torch.equal(probs.grad, probs.detach())  # the probabilities themselves stand in for the gradient
    x = torch.tensor([0.1, 0.5, 0.9])
    assert torch.allclose(torch.sigmoid(functions.sigmoid_inverse(x)), x, atol=1e-6)
    assert torch.isinf(functions.sigmoid_inverse(torch.tensor([0.0, 1.0]))).all()


def test_pseudo_sort_helpers() -> None:
    assert functions.select_zeroeth_element(np.array([[4.0, 1.0], [2.0, 3.0]])) == 4.0
    assert functions.select_zeroeth_element(torch.zeros(0)) == 0.0
    arrays = [np.array([1.0, 9.0]), torch.tensor([[2.0]]), np.array(["layer_name"])]
    assert functions.pseudo_sort_scoring_function((None, arrays, 10)) == 13.0


# ---- client helpers ----------------------------------------------------------------------------------------------
def test_client_helper_functions() -> None:
    metrics: dict = {}
    client_utils.fold_loss_dict_into_metrics(metrics, {"checkpoint": 0.5}, LoggingMode.VALIDATION)
    client_utils.fold_loss_dict_into_metrics(metrics, {"checkpoint": 0.7}, LoggingMode.TEST)
    assert metrics == {"val - checkpoint": 0.5, "test - checkpoint": 0.7}
    assert client_utils.set_pack_losses_with_val_metrics({"pack_losses_with_val_metrics": True})
    assert not client_utils.set_pack_losses_with_val_metrics({})

    moved = client_utils.move_data_to_device({"a": torch.ones(2), "b": torch.zeros(2)}, torch.device("cpu"))
    assert set(moved) == {"a", "b"}
    with pytest.raises(TypeError):
        client_utils.move_data_to_device([torch.ones(2)], torch.device("cpu"))  # type: ignore[type-var]

    assert client_utils.check_if_batch_is_empty_and_verify_input(torch.zeros(0, 3))
    assert not client_utils.check_if_batch_is_empty_and_verify_input({"x": torch.zeros(2, 3), "y": torch.zeros(2)})
    with pytest.raises(ValueError, match="same size"):
        client_utils.check_if_batch_is_empty_and_verify_input({"x": torch.zeros(2, 3), "y": torch.zeros(3)})

    model = nn.Sequential(nn.Linear(2, 2), nn.BatchNorm1d(2))
    model.train()
    frozen = client_utils.clone_and_freeze_model(model)
    assert not frozen.training and all(not p.requires_grad for p in frozen.parameters())
    assert all(p.requires_grad for p in model.parameters()) and frozen[0].weight is not model[0].weight

    assert list(client_utils.maybe_progress_bar(range(3), False)) == [0, 1, 2]
    assert list(client_utils.maybe_progress_bar(range(3), True)) == [0, 1, 2]

    class Loader:
        batch_size = 4

        def __len__(self) -> int:
            return 5

    assert client_utils.process_and_check_validation_steps({}, Loader()) is None
    assert client_utils.process_and_check_validation_steps({"num_validation_steps": 3}, Loader()) == 3
    with pytest.raises(AssertionError):
        client_utils.process_and_check_validation_steps({"num_validation_steps": 0}, Loader())
