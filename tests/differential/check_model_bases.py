"""Model bases vs the reference: state-dict keys (checkpoints are interchangeable), which layers are exchanged with the
server, and the forward outputs / stored features for the same weights and inputs."""
import copy

import torch
from torch import nn

import fl4health.model_bases.apfl_base as ref_apfl
import fl4health.model_bases.autoencoders_base as ref_ae
import fl4health.model_bases.ensemble_base as ref_ens
import fl4health.model_bases.fedsimclr_base as ref_simclr
import fl4health.model_bases.fenda_base as ref_fenda
import fl4health.model_bases.gpfl_base as ref_gpfl
import fl4health.model_bases.masked_layers.masked_conv as ref_mconv
import fl4health.model_bases.masked_layers.masked_linear as ref_mlin
import fl4health.model_bases.masked_layers.masked_normalization_layers as ref_mnorm
import fl4health.model_bases.moon_base as ref_moon
import fl4health.model_bases.parallel_split_models as ref_par
import fl4health.model_bases.pca as ref_pca
import fl4health.model_bases.perfcl_base as ref_perfcl
import fl4health.model_bases.sequential_split_models as ref_seq
import fl4health_b200.model_bases.apfl_base as my_apfl
import fl4health_b200.model_bases.autoencoders_base as my_ae
import fl4health_b200.model_bases.ensemble_base as my_ens
import fl4health_b200.model_bases.fedsimclr_base as my_simclr
import fl4health_b200.model_bases.fenda_base as my_fenda
import fl4health_b200.model_bases.gpfl_base as my_gpfl
import fl4health_b200.model_bases.masked_layers.masked_conv as my_mconv
import fl4health_b200.model_bases.masked_layers.masked_linear as my_mlin
import fl4health_b200.model_bases.masked_layers.masked_normalization_layers as my_mnorm
import fl4health_b200.model_bases.moon_base as my_moon
import fl4health_b200.model_bases.parallel_split_models as my_par
import fl4health_b200.model_bases.pca as my_pca
import fl4health_b200.model_bases.perfcl_base as my_perfcl
import fl4health_b200.model_bases.sequential_split_models as my_seq

torch.manual_seed(41)
agreed = 0


def same(a, b, tol=1e-6) -> None:
    if isinstance(a, dict):
        assert a.keys() == b.keys(), (a.keys(), b.keys())
        for key in a:
            same(a[key], b[key], tol)
    elif isinstance(a, (tuple, list)):
        assert len(a) == len(b)
        for x, y in zip(a, b):
            same(x, y, tol)
    else:
        a, b = a.float(), b.float()  # (the reference's vote returns integer one-hots, ours the predictions' dtype)
        assert a.shape == b.shape and torch.allclose(a, b, atol=tol, rtol=tol), (a.shape, b.shape, (a - b).abs().max())


def twin(theirs: nn.Module, ours: nn.Module) -> None:
    """Same parameter / buffer names, in the same order; then the same values."""
    assert list(theirs.state_dict()) == list(ours.state_dict()), (list(theirs.state_dict()), list(ours.state_dict()))
    ours.load_state_dict(theirs.state_dict())


def extractor() -> nn.Module:
    return nn.Sequential(nn.Conv2d(1, 3, 3), nn.ReLU(), nn.Flatten(), nn.Linear(3 * 6 * 6, 10))


def head(width: int = 10, classes: int = 4) -> nn.Module:
    return nn.Sequential(nn.ReLU(), nn.Linear(width, classes))


images = torch.randn(5, 1, 8, 8)

# sequential split (+ exchange-base variant), MOON
for flatten in (False, True):
    theirs, ours = ref_seq.SequentiallySplitModel(extractor(), head(), flatten), my_seq.SequentiallySplitModel(extractor(), head(), flatten)
    twin(theirs, ours); same(theirs(images), ours(images)); agreed += 1
theirs, ours = ref_seq.SequentiallySplitExchangeBaseModel(extractor(), head(), False), my_seq.SequentiallySplitExchangeBaseModel(extractor(), head(), False)
twin(theirs, ours); assert theirs.layers_to_exchange() == ours.layers_to_exchange(); agreed += 1
for projection in (None, nn.Linear(10, 6)):
    width = 10 if projection is None else 6  # the head consumes the projected features
    theirs = ref_moon.MoonModel(extractor(), head(width), copy.deepcopy(projection))
    ours = my_moon.MoonModel(extractor(), head(width), copy.deepcopy(projection))
    twin(theirs, ours); same(theirs(images), ours(images)); agreed += 1


# parallel split family: FENDA, FENDA with feature state, PerFCL
def join_heads(module):
    class Concat(module.ParallelSplitHeadModule):
        def __init__(self, mode) -> None:
            super().__init__(mode)
            self.classifier = nn.Linear(20 if mode == module.ParallelFeatureJoinMode.CONCATENATE else 10, 4)

        def parallel_output_join(self, local_tensor, global_tensor):
            return torch.cat([local_tensor, global_tensor], dim=1)

        def head_forward(self, input_tensor):
            return self.classifier(input_tensor)

    return Concat


for mode in ("CONCATENATE", "SUM"):
    h_ref = join_heads(ref_par)(getattr(ref_par.ParallelFeatureJoinMode, mode))
    h_mine = join_heads(my_par)(getattr(my_par.ParallelFeatureJoinMode, mode))
    for ref_cls, my_cls, kwargs in (
        (ref_par.ParallelSplitModel, my_par.ParallelSplitModel, {}), (ref_fenda.FendaModel, my_fenda.FendaModel, {}),
        (ref_fenda.FendaModelWithFeatureState, my_fenda.FendaModelWithFeatureState, {"flatten_features": True}),
        (ref_perfcl.PerFclModel, my_perfcl.PerFclModel, {}),
    ):
        theirs = ref_cls(extractor(), extractor(), copy.deepcopy(h_ref), **kwargs)
        ours = my_cls(extractor(), extractor(), copy.deepcopy(h_mine), **kwargs)
        twin(theirs, ours); same(theirs(images), ours(images))
        if hasattr(theirs, "layers_to_exchange"):
            assert theirs.layers_to_exchange() == ours.layers_to_exchange()
        agreed += 1

# APFL: personal / global mixture and the alpha update after one backward pass
for adaptive in (True, False):
    theirs, ours = ref_apfl.ApflModule(extractor(), adaptive, 0.3, 0.05), my_apfl.ApflModule(extractor(), adaptive, 0.3, 0.05)
    twin(theirs, ours)
    out_ref, out_mine = theirs(images), ours(images)
    same(out_ref, out_mine)
    assert theirs.layers_to_exchange() == ours.layers_to_exchange()
    for out in (out_ref, out_mine):
        (out["personal"].pow(2).mean() + out["local"].pow(2).mean() + out["global"].pow(2).mean()).backward()
    theirs.update_alpha(); ours.update_alpha()
    assert abs(float(theirs.alpha) - float(ours.alpha)) < 1e-6, (theirs.alpha, ours.alpha)
    agreed += 1

# GPFL: conditional-value modulation and the global conditional embedding
theirs, ours = ref_gpfl.GpflModel(extractor(), head(), 10, 4), my_gpfl.GpflModel(extractor(), head(), 10, 4)
twin(theirs, ours)
assert theirs.layers_to_exchange() == ours.layers_to_exchange()
conditional_global, conditional_personal = torch.randn(10), torch.randn(10)
for training in (True, False):
    theirs.train(training); ours.train(training)
    same(theirs(images, conditional_global, conditional_personal), ours(images, conditional_global, conditional_personal))
labels = torch.randint(0, 4, (5,))
features = torch.randn(5, 10)
same(theirs.gce(features, labels), ours.gce(features, labels)); same(theirs.gce.lookup(labels), ours.gce.lookup(labels))
agreed += 1

# ensembles
for mode in ("AVERAGE", "VOTE"):
    members = {f"m{i}": extractor() for i in range(3)}
    theirs = ref_ens.EnsembleModel(copy.deepcopy(members), getattr(ref_ens.EnsembleAggregationMode, mode))
    ours = my_ens.EnsembleModel(copy.deepcopy(members), getattr(my_ens.EnsembleAggregationMode, mode))
    twin(theirs, ours); same(theirs(images), ours(images)); agreed += 1

# SimCLR: pre-training returns projections, fine-tuning predictions
for pretrain in (True, False):
    kwargs = dict(projection_head=nn.Linear(10, 6), prediction_head=nn.Linear(10, 4), pretrain=pretrain)
    theirs = ref_simclr.FedSimClrModel(extractor(), **copy.deepcopy(kwargs)); ours = my_simclr.FedSimClrModel(extractor(), **copy.deepcopy(kwargs))
    twin(theirs, ours); same(theirs(images), ours(images)); agreed += 1

# auto-encoders (the random reparameterisation draw is seeded identically)
encoder = nn.Sequential(nn.Flatten(), nn.Linear(64, 12))
decoder = nn.Sequential(nn.Linear(12, 64))
theirs, ours = ref_ae.BasicAe(copy.deepcopy(encoder), copy.deepcopy(decoder)), my_ae.BasicAe(copy.deepcopy(encoder), copy.deepcopy(decoder))
twin(theirs, ours); same(theirs(images), ours(images)); agreed += 1


class GaussianEncoder(nn.Module):
    def __init__(self, inputs: int = 64) -> None:
        super().__init__()
        self.mu, self.logvar = nn.Linear(inputs, 6), nn.Linear(inputs, 6)

    def forward(self, x, condition=None):
        x = x.flatten(1) if condition is None else torch.cat([x.flatten(1), condition], dim=1)
        return self.mu(x), self.logvar(x)


class ConditionalDecoder(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.out = nn.Linear(6 + 3, 64)

    def forward(self, z, condition):
        return self.out(torch.cat([z, condition], dim=1))


theirs, ours = ref_ae.VariationalAe(GaussianEncoder(), nn.Linear(6, 64)), my_ae.VariationalAe(GaussianEncoder(), nn.Linear(6, 64))
twin(theirs, ours)
torch.manual_seed(1); a = theirs(images)
torch.manual_seed(1); b = ours(images)
same(a, b); agreed += 1
condition = torch.nn.functional.one_hot(torch.randint(0, 3, (5,)), 3).float()
unpack = lambda packed: (packed[:, :-3].reshape(-1, 1, 8, 8), packed[:, -3:])  # noqa: E731
theirs = ref_ae.ConditionalVae(GaussianEncoder(67), ConditionalDecoder(), unpack_input_condition=unpack)
ours = my_ae.ConditionalVae(GaussianEncoder(67), ConditionalDecoder(), unpack_input_condition=unpack)
twin(theirs, ours)
packed = torch.cat([images.flatten(1), condition], dim=1)
torch.manual_seed(2); a = theirs(packed)
torch.manual_seed(2); b = ours(packed)
same(a, b); agreed += 1

# masked layers (FedPM): scores -> Bernoulli masks, seeded identically
for ref_cls, my_cls, make, data in (
    (ref_mlin.MaskedLinear, my_mlin.MaskedLinear, lambda c: c(7, 5), torch.randn(4, 7)),
    (ref_mconv.MaskedConv2d, my_mconv.MaskedConv2d, lambda c: c(2, 3, 3, padding=1), torch.randn(2, 2, 6, 6)),
    (ref_mconv.MaskedConv1d, my_mconv.MaskedConv1d, lambda c: c(2, 3, 3), torch.randn(2, 2, 9)),
    (ref_mconv.MaskedConvTranspose2d, my_mconv.MaskedConvTranspose2d, lambda c: c(2, 3, 3), torch.randn(2, 2, 5, 5)),
    (ref_mnorm.MaskedLayerNorm, my_mnorm.MaskedLayerNorm, lambda c: c(6), torch.randn(4, 6)),
    (ref_mnorm.MaskedBatchNorm2d, my_mnorm.MaskedBatchNorm2d, lambda c: c(3), torch.randn(4, 3, 5, 5)),
):
    torch.manual_seed(3); theirs = make(ref_cls)
    torch.manual_seed(3); ours = make(my_cls)
    twin(theirs, ours)
    torch.manual_seed(4); a = theirs(data)
    torch.manual_seed(4); b = ours(data)
    same(a, b, 1e-5)
    assert [n for n, p in theirs.named_parameters() if p.requires_grad] == [n for n, p in ours.named_parameters() if p.requires_grad]
    agreed += 1
plain = nn.Linear(7, 5)
theirs, ours = ref_mlin.MaskedLinear.from_pretrained(copy.deepcopy(plain)), my_mlin.MaskedLinear.from_pretrained(copy.deepcopy(plain))
assert torch.equal(theirs.weight, ours.weight) and list(theirs.state_dict()) == list(ours.state_dict()); agreed += 1

# PCA module
data = torch.randn(40, 9) @ torch.randn(9, 9)
for low_rank, full_svd in ((False, False), (False, True), (True, False)):
    theirs, ours = ref_pca.PcaModule(low_rank, full_svd, 5), my_pca.PcaModule(low_rank, full_svd, 5)
    torch.manual_seed(6); (pc_ref, sv_ref) = theirs(data, center_data=True)
    torch.manual_seed(6); (pc_mine, sv_mine) = ours(data, center_data=True)
    k = min(sv_ref.numel(), 4)
    assert torch.allclose(sv_ref[:k], sv_mine[:k], rtol=1e-3, atol=1e-3), (sv_ref, sv_mine)
    theirs.set_principal_components(pc_ref, sv_ref); ours.set_principal_components(pc_ref, sv_ref)
    same(theirs.project_lower_dim(data, 3, center_data=True), ours.project_lower_dim(data, 3, center_data=True), 1e-4)
    same(theirs.project_back(theirs.project_lower_dim(data, 3, center_data=True), add_mean=True), ours.project_back(ours.project_lower_dim(data, 3, center_data=True), add_mean=True), 1e-4)
    assert abs(theirs.compute_reconstruction_error(data, 3, center_data=True) - ours.compute_reconstruction_error(data, 3, center_data=True)) < 1e-3
    assert abs(theirs.compute_projection_variance(data, 3, center_data=True) - ours.compute_projection_variance(data, 3, center_data=True)) < 1e-2
    assert abs(float(theirs.compute_cumulative_explained_variance()) - float(ours.compute_cumulative_explained_variance())) < 1e-2
    same(theirs.compute_explained_variance_ratios(), ours.compute_explained_variance_ratios(), 1e-4)
    agreed += 1
print("configs agree:", agreed)
