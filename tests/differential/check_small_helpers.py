"""Small helpers vs the reference, including how they FAIL: config loading / checking / narrowing, client utilities,
random-state save / restore, parameter extraction, data splitting."""
import random
import tempfile
from pathlib import Path

import numpy as np
import torch
from torch import nn
from torch.utils.data import DataLoader

import fl4health.utils.client as ref_client
import fl4health.utils.config as ref_config
import fl4health.utils.dataset as ref_ds
import fl4health.utils.parameter_extraction as ref_extract
import fl4health.utils.random as ref_random
import fl4health_b200.utils.client as my_client
import fl4health_b200.utils.config as my_config
import fl4health_b200.utils.dataset as my_ds
import fl4health_b200.utils.parameter_extraction as my_extract
import fl4health_b200.utils.random as my_random

import flwr.common as fc
import fl4health_b200.common.typing as mt

agreed = 0


def outcome(fn, *args, **kwargs):
    """('ok', value) or ('raises', exception type name): both sides must agree on which."""
    try:
        return "ok", fn(*args, **kwargs)
    except Exception as error:  # noqa: BLE001
        return "raises", type(error).__name__


def same_outcome(ref_fn, my_fn, *args, compare=lambda a, b: a == b, **kwargs) -> None:
    global agreed
    a, b = outcome(ref_fn, *args, **kwargs), outcome(my_fn, *args, **kwargs)
    assert a[0] == b[0], (ref_fn.__name__, args, a, b)
    if a[0] == "ok":
        assert compare(a[1], b[1]), (ref_fn.__name__, args, a[1], b[1])
    else:
        assert a[1] == b[1], (ref_fn.__name__, args, a, b)
    agreed += 1


# -- config ----------------------------------------------------------------------------------------------------------
scratch = Path(tempfile.mkdtemp(prefix="fl4h_cfg_"))
documents = {
    "good": "n_server_rounds: 5\nn_clients: 3\nbatch_size: 32\nlocal_steps: 10\n",
    "good_epochs": "n_server_rounds: 5\nn_clients: 3\nbatch_size: 32\nlocal_epochs: 2\nextra: [1, 2]\n",
    "missing_rounds": "n_clients: 3\nbatch_size: 32\nlocal_steps: 10\n",
    "missing_batch": "n_server_rounds: 5\nlocal_steps: 10\n",
    "wrong_type": "n_server_rounds: five\nbatch_size: 32\nlocal_steps: 10\n",
    "negative": "n_server_rounds: -1\nbatch_size: 32\nlocal_steps: 10\n",
    "float_rounds": "n_server_rounds: 2.5\nbatch_size: 32\nlocal_steps: 10\n",
}
for name, text in documents.items():
    path = scratch / f"{name}.yaml"
    path.write_text(text)
    same_outcome(ref_config.load_config, my_config.load_config, str(path))
same_outcome(ref_config.load_config, my_config.load_config, str(scratch / "does_not_exist.yaml"))
bag = {"rounds": 3, "rate": 0.5, "name": "x", "flag": True, "none": None}
for key, kind in (("rounds", int), ("rounds", float), ("rate", float), ("rate", int), ("name", str), ("flag", bool), ("flag", int), ("absent", int), ("none", int)):
    same_outcome(ref_config.narrow_dict_type, my_config.narrow_dict_type, bag, key, kind)
for epochs, steps in ((2, None), (None, 7), (2, 7), (None, None)):
    same_outcome(ref_config.make_dict_with_epochs_or_steps, my_config.make_dict_with_epochs_or_steps, epochs, steps)


class Holder:
    pass


for key, attribute, kind, fn in (("rounds", "n", int, None), ("rate", "r", float, lambda v: v * 2), ("absent", "a", int, None), ("name", "s", int, None)):
    a_holder, b_holder = Holder(), Holder()
    a = outcome(ref_config.narrow_dict_type_and_set_attribute, a_holder, bag, key, attribute, kind, fn)
    b = outcome(my_config.narrow_dict_type_and_set_attribute, b_holder, bag, key, attribute, kind, fn)
    assert a[0] == b[0] and vars(a_holder) == vars(b_holder), (key, a, b, vars(a_holder), vars(b_holder))
    agreed += 1

# -- client utilities --------------------------------------------------------------------------------------------------
metrics_a, metrics_b = {"acc": 0.5}, {"acc": 0.5}
ref_client.fold_loss_dict_into_metrics(metrics_a, {"checkpoint": 1.5, "extra": 0.25}, ref_client.LoggingMode.VALIDATION)
my_client.fold_loss_dict_into_metrics(metrics_b, {"checkpoint": 1.5, "extra": 0.25}, my_client.LoggingMode.VALIDATION)
assert metrics_a == metrics_b, (metrics_a, metrics_b); agreed += 1
for config in ({}, {"pack_losses_with_val_metrics": True}, {"pack_losses_with_val_metrics": False}, {"pack_losses_with_val_metrics": "yes"}):
    same_outcome(ref_client.set_pack_losses_with_val_metrics, my_client.set_pack_losses_with_val_metrics, dict(config))
cpu = torch.device("cpu")
for data in (torch.ones(2, 3), {"a": torch.ones(2), "b": torch.zeros(3)}, [torch.ones(2)], "text"):
    same_outcome(ref_client.move_data_to_device, my_client.move_data_to_device, data, cpu,
                 compare=lambda a, b: type(a) is type(b) and (torch.equal(a, b) if isinstance(a, torch.Tensor) else all(torch.equal(a[k], b[k]) for k in a)))
for batch in (torch.ones(2, 3), torch.ones(0, 3), {"a": torch.ones(2, 3), "b": torch.ones(2)}, {"a": torch.ones(0, 3), "b": torch.ones(0)},
              {"a": torch.ones(2, 3), "b": torch.ones(3)}, [torch.ones(2)]):
    same_outcome(ref_client.check_if_batch_is_empty_and_verify_input, my_client.check_if_batch_is_empty_and_verify_input, batch)
net = nn.Sequential(nn.Linear(3, 2), nn.BatchNorm1d(2))
frozen_ref, frozen_mine = ref_client.clone_and_freeze_model(net), my_client.clone_and_freeze_model(net)
for a, b in ((frozen_ref, frozen_mine),):
    assert a is not net and b is not net and not a.training and not b.training
    assert all(not p.requires_grad for p in a.parameters()) and all(not p.requires_grad for p in b.parameters())
    assert all(torch.equal(x, y) for x, y in zip(a.state_dict().values(), b.state_dict().values()))
agreed += 1
loader = DataLoader(ref_ds.TensorDataset(torch.randn(40, 3), torch.zeros(40).long()), batch_size=8)
for config in ({}, {"num_validation_steps": 2}, {"num_validation_steps": 5}, {"num_validation_steps": 9}, {"num_validation_steps": 0}, {"num_validation_steps": 1.5}):
    same_outcome(ref_client.process_and_check_validation_steps, my_client.process_and_check_validation_steps, dict(config), loader)
assert list(ref_client.maybe_progress_bar(range(4), False)) == list(my_client.maybe_progress_bar(range(4), False)); agreed += 1

# -- random state ------------------------------------------------------------------------------------------------------
for module in (ref_random, my_random):
    module.set_all_random_seeds(2024)
    first = (random.random(), float(np.random.rand()), float(torch.rand(())))
    state = module.save_random_state()
    second = (random.random(), float(np.random.rand()), float(torch.rand(())))
    module.restore_random_state(*state)
    assert (random.random(), float(np.random.rand()), float(torch.rand(()))) == second
    module.set_all_random_seeds(2024)
    assert (random.random(), float(np.random.rand()), float(torch.rand(()))) == first
    module.unset_all_random_seeds()
ref_random.set_all_random_seeds(7); a = (random.random(), float(np.random.rand()), float(torch.rand(())))
my_random.set_all_random_seeds(7); b = (random.random(), float(np.random.rand()), float(torch.rand(())))
assert a == b; agreed += 1  # the same seed drives the same streams
assert len(ref_random.generate_hash()) == len(my_random.generate_hash()) and len(ref_random.generate_hash(12)) == len(my_random.generate_hash(12)) == 12; agreed += 1

# -- parameter extraction ----------------------------------------------------------------------------------
a = fc.parameters_to_ndarrays(ref_extract.get_all_model_parameters(net))
b = [x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x) for x in mt.parameters_to_ndarrays(my_extract.get_all_model_parameters(net))]
assert len(a) == len(b) and all(x.shape == y.shape and np.array_equal(x, y) for x, y in zip(a, b)); agreed += 1
same_outcome(ref_extract.check_shape_match, my_extract.check_shape_match, [torch.ones(2, 3)], [torch.ones(2, 3)], "mismatch")
same_outcome(ref_extract.check_shape_match, my_extract.check_shape_match, [torch.ones(2, 3)], [torch.ones(3, 2)], "mismatch")
# (``utils.load_data`` of the reference imports monai at module load: not comparable in this image)

# -- client managers: how many clients a fraction selects, "all", the fixed sample's contract ------------------------------------
import fl4health.client_managers.fixed_sampling_client_manager as ref_fixed
import fl4health.client_managers.fixed_without_replacement_manager as ref_fraction
import fl4health.client_managers.poisson_sampling_manager as ref_poisson
import fl4health_b200.client_managers.fixed_sampling_client_manager as my_fixed
import fl4health_b200.client_managers.fixed_without_replacement_manager as my_fraction
import fl4health_b200.client_managers.poisson_sampling_manager as my_poisson


class Proxy:
    def __init__(self, cid: str) -> None:
        self.cid = cid


def populated(manager, n: int = 10):
    for index in range(n):
        manager.register(Proxy(f"c{index}"))
    return manager


for fraction in (0.1, 0.25, 0.34, 0.5, 0.75, 0.99, 1.0):
    theirs = populated(ref_fraction.FixedSamplingByFractionClientManager())
    ours = populated(my_fraction.FixedSamplingByFractionClientManager())
    picked_ref, picked_mine = theirs.sample_fraction(fraction, 1), ours.sample_fraction(fraction, 1)
    assert len(picked_ref) == len(picked_mine), (fraction, len(picked_ref), len(picked_mine))
    assert len({p.cid for p in picked_mine}) == len(picked_mine)  # without replacement
    agreed += 1
for module_ref, module_mine, name in ((ref_fraction, my_fraction, "FixedSamplingByFractionClientManager"), (ref_poisson, my_poisson, "PoissonSamplingClientManager")):
    theirs, ours = populated(getattr(module_ref, name)()), populated(getattr(module_mine, name)())
    assert sorted(p.cid for p in theirs.sample_all(10)) == sorted(p.cid for p in ours.sample_all(10)); agreed += 1
sizes_ref = [len(populated(ref_poisson.PoissonSamplingClientManager(), 200).sample_fraction(0.3, 1)) for _ in range(30)]
sizes_mine = [len(populated(my_poisson.PoissonSamplingClientManager(), 200).sample_fraction(0.3, 1)) for _ in range(30)]
assert abs(np.mean(sizes_ref) - 60) < 8 and abs(np.mean(sizes_mine) - 60) < 8 and np.std(sizes_mine) > 2, (np.mean(sizes_ref), np.mean(sizes_mine))  # Binomial(200, 0.3)
agreed += 1
theirs, ours = populated(ref_fixed.FixedSamplingClientManager()), populated(my_fixed.FixedSamplingClientManager())
for manager in (theirs, ours):
    first = [p.cid for p in manager.sample(4, 4)]
    assert [p.cid for p in manager.sample(4, 4)] == first  # the same cohort until reset (fit and evaluate hit the same clients)
    manager.reset_sample()
    assert len(manager.sample(4, 4)) == 4
agreed += 1
print("configs agree:", agreed)
