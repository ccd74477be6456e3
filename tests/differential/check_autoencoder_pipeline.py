"""Auto-encoder preprocessing vs the reference: dataset converters (targets replaced by the data; a fixed condition or the
one-hot label concatenated to the input; unpacking on the model side) and the processors that map samples to latent codes
with a saved AE / VAE / conditional VAE (random reparameterisation seeded identically)."""
import tempfile
from pathlib import Path

import torch
from torch import nn

import fl4health.model_bases.autoencoders_base as ref_ae
import fl4health.preprocessing.autoencoders.dim_reduction as ref_proc
import fl4health.utils.dataset as ref_ds
import fl4health.utils.dataset_converter as ref_conv
import fl4health_b200.model_bases.autoencoders_base as my_ae
import fl4health_b200.preprocessing.autoencoders.dim_reduction as my_proc
import fl4health_b200.utils.dataset as my_ds
import fl4health_b200.utils.dataset_converter as my_conv

agreed = 0
generator = torch.Generator().manual_seed(71)
data, labels = torch.randn(12, 1, 4, 4, generator=generator), torch.randint(0, 3, (12,), generator=generator)

# -- dataset converters ----------------------------------------------------------------------------------------------
fixed_condition = torch.tensor([1.0, 0.0, 2.0])
for kwargs in (dict(), dict(condition=fixed_condition), dict(condition="label", do_one_hot_encoding=True), dict(condition="label", do_one_hot_encoding=False)):
    theirs, ours = ref_conv.AutoEncoderDatasetConverter(**kwargs), my_conv.AutoEncoderDatasetConverter(**kwargs)
    # without one-hot encoding the label itself is concatenated: it has to be a vector already
    targets = labels if kwargs.get("do_one_hot_encoding", True) or "condition" not in kwargs or not isinstance(kwargs["condition"], str) \
        else torch.nn.functional.one_hot(labels, 3).float()
    converted_ref = theirs.convert_dataset(ref_ds.TensorDataset(data.clone(), targets.clone()))
    converted_mine = ours.convert_dataset(my_ds.TensorDataset(data.clone(), targets.clone()))
    assert len(converted_ref) == len(converted_mine)
    for index in range(len(converted_ref)):
        (x_ref, y_ref), (x_mine, y_mine) = converted_ref[index], converted_mine[index]
        assert x_ref.shape == x_mine.shape and torch.allclose(x_ref.float(), x_mine.float()), (kwargs, index)
        assert y_ref.shape == y_mine.shape and torch.allclose(y_ref.float(), y_mine.float()), (kwargs, index)
    if kwargs:  # (without a condition the reference has no ``condition_vector_size`` attribute to report)
        assert theirs.get_condition_vector_size() == ours.get_condition_vector_size(), kwargs
        batch = torch.stack([converted_ref[i][0] for i in range(4)])
        (a_data, a_condition), (b_data, b_condition) = theirs.get_unpacking_function()(batch), ours.get_unpacking_function()(batch)
        assert torch.equal(a_data, b_data) and torch.equal(a_condition, b_condition), kwargs
    agreed += 1


# -- processors over saved models ----------------------------------------------------------------------------------------
class GaussianEncoder(nn.Module):
    def __init__(self, inputs: int) -> None:
        super().__init__()
        self.mu, self.logvar = nn.Linear(inputs, 5), nn.Linear(inputs, 5)

    def forward(self, x, condition=None):
        x = x.flatten(1) if x.dim() > 1 else x  # a single (already flat) sample or a batch
        x = x if condition is None else torch.cat([x, condition], dim=-1)
        return self.mu(x), self.logvar(x)


class ConditionalDecoder(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.out = nn.Linear(5 + 3, 16)

    def forward(self, z, condition):
        return self.out(torch.cat([z, condition], dim=1))


def pinned(module: nn.Module) -> nn.Module:
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for parameter in module.parameters():
            parameter.copy_(torch.randn(parameter.shape, generator=g) * 0.3)
    return module


scratch = Path(tempfile.mkdtemp(prefix="fl4h_ae_"))
def unpack(packed: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:  # module level: the saved model pickles it by name
    return packed[:, :-3].reshape(-1, 1, 4, 4), packed[:, -3:]


builders = {
    "ae": lambda m: m.BasicAe(pinned(nn.Sequential(nn.Flatten(), nn.Linear(16, 5))), pinned(nn.Linear(5, 16))),
    "vae": lambda m: m.VariationalAe(pinned(GaussianEncoder(16)), pinned(nn.Linear(5, 16))),
    "cvae": lambda m: m.ConditionalVae(pinned(GaussianEncoder(19)), pinned(ConditionalDecoder()), unpack_input_condition=unpack),
}
for kind, build in builders.items():
    torch.save(build(ref_ae), scratch / f"{kind}_ref.pt")
    torch.save(build(my_ae), scratch / f"{kind}_mine.pt")
cpu = torch.device("cpu")
samples = data.flatten(1)
theirs, ours = ref_proc.AeProcessor(scratch / "ae_ref.pt", cpu), my_proc.AeProcessor(scratch / "ae_mine.pt", cpu)
assert torch.allclose(theirs(data), ours(data), atol=1e-6); agreed += 1
for mu_only in (True, False):
    theirs, ours = ref_proc.VaeProcessor(scratch / "vae_ref.pt", cpu, mu_only), my_proc.VaeProcessor(scratch / "vae_mine.pt", cpu, mu_only)
    for sample in (data,):
        torch.manual_seed(9); a = theirs(sample)
        torch.manual_seed(9); b = ours(sample)
        assert a.shape == b.shape and torch.allclose(a, b, atol=1e-6), (mu_only, a.shape, b.shape)
    agreed += 1
    condition = torch.tensor([0.0, 1.0, 0.0])
    theirs = ref_proc.CvaeFixedConditionProcessor(scratch / "cvae_ref.pt", condition, cpu, mu_only)
    ours = my_proc.CvaeFixedConditionProcessor(scratch / "cvae_mine.pt", condition, cpu, mu_only)
    for sample in (samples, samples[0]):
        torch.manual_seed(10); a = theirs(sample)
        torch.manual_seed(10); b = ours(sample)
        assert a.shape == b.shape and torch.allclose(a, b, atol=1e-6), (mu_only, a.shape, b.shape)
    agreed += 1
    theirs = ref_proc.CvaeVariableConditionProcessor(scratch / "cvae_ref.pt", cpu, mu_only)
    ours = my_proc.CvaeVariableConditionProcessor(scratch / "cvae_mine.pt", cpu, mu_only)
    conditions = torch.nn.functional.one_hot(labels, 3).float()
    for sample, cond in ((samples, conditions),):
        torch.manual_seed(11); a = theirs(sample, cond)
        torch.manual_seed(11); b = ours(sample, cond)
        assert a.shape == b.shape and torch.allclose(a, b, atol=1e-6), (mu_only, a.shape, b.shape)
    agreed += 1
print("configs agree:", agreed)
