"""Wire layouts of the parameter exchangers and packers vs the reference: what is pushed for the same model (number,
order, shapes, dtypes, values of the arrays) and what a model looks like after pulling the same payload."""
import copy

import numpy as np
import torch
from torch import nn

import fl4health.parameter_exchange.full_exchanger as ref_full
import fl4health.parameter_exchange.layer_exchanger as ref_layer
import fl4health.parameter_exchange.parameter_packer as ref_pack
import fl4health.parameter_exchange.parameter_selection_criteria as ref_sel
import fl4health.parameter_exchange.sparse_coo_parameter_exchanger as ref_coo
import fl4health_b200.parameter_exchange.full_exchanger as my_full
import fl4health_b200.parameter_exchange.layer_exchanger as my_layer
import fl4health_b200.parameter_exchange.parameter_packer as my_pack
import fl4health_b200.parameter_exchange.parameter_selection_criteria as my_sel
import fl4health_b200.parameter_exchange.sparse_coo_parameter_exchanger as my_coo

torch.manual_seed(21)
agreed = 0


class Net(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.conv = nn.Conv2d(1, 4, 3)
        self.norm = nn.BatchNorm2d(4)
        self.body = nn.Linear(16, 8)
        self.head = nn.Linear(8, 3)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.head(torch.relu(self.body(torch.relu(self.norm(self.conv(x))).flatten(1)[:, :16])))


def as_numpy(arrays) -> list[np.ndarray]:
    return [a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a) for a in arrays]


def same_arrays(a, b) -> None:
    a, b = as_numpy(a), as_numpy(b)
    assert len(a) == len(b), (len(a), len(b))
    for x, y in zip(a, b):
        assert x.shape == y.shape and x.dtype.kind == y.dtype.kind, (x.shape, y.shape, x.dtype, y.dtype)
        if x.dtype.kind in "US":
            assert (x == y).all()
        else:
            assert np.allclose(x, y, atol=1e-6), np.abs(x - y).max()


def same_models(a: nn.Module, b: nn.Module) -> None:
    for (name, x), (_, y) in zip(a.state_dict().items(), b.state_dict().items()):
        assert torch.equal(x, y), name


def drifted(model: nn.Module) -> nn.Module:
    other = copy.deepcopy(model)
    with torch.no_grad():
        for index, p in enumerate(other.parameters()):
            p.add_(0.05 * (index + 1) * torch.randn_like(p))
    return other


initial = Net()
model = drifted(initial)
model(torch.randn(5, 1, 6, 6))  # moves the BatchNorm buffers

# -- exchangers -----------------------------------------------------------------------------------------------
pairs = [
    (ref_full.FullParameterExchanger(), my_full.FullParameterExchanger()),
    (ref_layer.FixedLayerExchanger(["body.weight", "head.bias", "norm.running_mean"]),
     my_layer.FixedLayerExchanger(["body.weight", "head.bias", "norm.running_mean"])),
    (ref_layer.LayerExchangerWithExclusions(model, {nn.BatchNorm2d}), my_layer.LayerExchangerWithExclusions(model, {nn.BatchNorm2d})),
    (ref_layer.LayerExchangerWithExclusions(model, {nn.Linear, nn.BatchNorm2d}), my_layer.LayerExchangerWithExclusions(model, {nn.Linear, nn.BatchNorm2d})),
]
for theirs, ours in pairs:
    pushed_ref, pushed_mine = theirs.push_parameters(model, initial), ours.push_parameters(model, initial)
    same_arrays(pushed_ref, pushed_mine)
    target_ref, target_mine = Net(), None
    target_mine = copy.deepcopy(target_ref)
    theirs.pull_parameters(as_numpy(pushed_ref), target_ref)
    ours.pull_parameters(as_numpy(pushed_ref), target_mine)
    same_models(target_ref, target_mine)
    agreed += 1

for threshold, percentage, normalize, more in ((0.05, 0.5, True, True), (0.2, 0.25, False, True), (0.05, 0.75, True, False)):
    for how in ("select_by_threshold", "select_by_percentage"):
        fn_ref = getattr(ref_sel.LayerSelectionFunctionConstructor(threshold, percentage, normalize, more), how)()
        fn_mine = getattr(my_sel.LayerSelectionFunctionConstructor(threshold, percentage, normalize, more), how)()
        theirs, ours = ref_layer.DynamicLayerExchanger(fn_ref), my_layer.DynamicLayerExchanger(fn_mine)
        pushed_ref, pushed_mine = theirs.push_parameters(model, initial), ours.push_parameters(model, initial)
        same_arrays(pushed_ref, pushed_mine)
        target_ref = Net(); target_mine = copy.deepcopy(target_ref)
        theirs.pull_parameters(as_numpy(pushed_ref), target_ref)
        ours.pull_parameters(as_numpy(pushed_ref), target_mine)
        same_models(target_ref, target_mine)
        agreed += 1

for sparsity in (0.1, 0.5):
    for score in ("largest_final_magnitude_scores", "smallest_final_magnitude_scores", "largest_magnitude_change_scores",
                  "smallest_magnitude_change_scores", "largest_increase_in_magnitude_scores", "smallest_increase_in_magnitude_scores"):
        theirs = ref_coo.SparseCooParameterExchanger(sparsity, getattr(ref_sel, score))
        ours = my_coo.SparseCooParameterExchanger(sparsity, getattr(my_sel, score))
        pushed_ref, pushed_mine = theirs.push_parameters(model, initial), ours.push_parameters(model, initial)
        same_arrays(pushed_ref, pushed_mine)
        target_ref = Net(); target_mine = copy.deepcopy(target_ref)
        theirs.pull_parameters(as_numpy(pushed_ref), target_ref)
        ours.pull_parameters(as_numpy(pushed_ref), target_mine)
        same_models(target_ref, target_mine)
        agreed += 1

# -- packers: pack with one implementation, unpack with the other ------------------------------------------------
weights = [p.detach().numpy() for p in model.parameters()]
variates = [np.random.rand(*w.shape).astype(np.float32) for w in weights]
packers = [
    (ref_pack.ParameterPackerWithControlVariates(len(weights)), my_pack.ParameterPackerWithControlVariates(len(weights)), variates),
    (ref_pack.ParameterPackerWithClippingBit(), my_pack.ParameterPackerWithClippingBit(), 1.0),
    (ref_pack.ParameterPackerAdaptiveConstraint(), my_pack.ParameterPackerAdaptiveConstraint(), 0.37),
    (ref_pack.ParameterPackerWithLayerNames(), my_pack.ParameterPackerWithLayerNames(), [n for n, _ in model.named_parameters()]),
]
for theirs, ours, extra in packers:
    packed_ref, packed_mine = theirs.pack_parameters(weights, extra), ours.pack_parameters(weights, extra)
    same_arrays(packed_ref, packed_mine)
    (w_a, e_a), (w_b, e_b) = theirs.unpack_parameters(as_numpy(packed_mine)), ours.unpack_parameters(as_numpy(packed_ref))
    same_arrays(w_a, w_b)
    if isinstance(extra, list) and isinstance(extra[0], np.ndarray):
        same_arrays(e_a, e_b)
    else:
        assert (list(e_a) == list(e_b)) if isinstance(extra, list) else (float(e_a) == float(e_b)), (e_a, e_b)
    agreed += 1

dense = torch.randn(4, 5) * (torch.rand(4, 5) > 0.6)
for a, b in zip(ref_pack.SparseCooParameterPacker.extract_coo_info_from_dense(dense), my_pack.SparseCooParameterPacker.extract_coo_info_from_dense(dense)):
    same_arrays([a], [b])
agreed += 1
print("configs agree:", agreed)
