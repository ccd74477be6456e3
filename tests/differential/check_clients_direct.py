"""Clients driven directly (no server), with the global generator seeded identically before every call on both sides where
the local step draws random numbers: FedPM (Bernoulli masks sampled from learnt scores in every forward; binary masks on
the wire), Ditto / MR-MTL with a Deep-MMD term (deep kernel trained on shuffled samples), the client-level-DP clipping
client (clipped update + clipping bit), the FedPCA client (local components, evaluation of merged components)."""
import sys
from pathlib import Path

import numpy as np
import torch
from torch import nn

sys.path.insert(0, str(Path(__file__).resolve().parent))
from check_federations import resolver, user_hooks  # noqa: E402

import flwr.common as fc
import fl4health_b200.common.typing as mt

agreed = 0


def pin_everything(module: nn.Module, seed: int = 7) -> nn.Module:
    generator = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for _, parameter in sorted(module.named_parameters()):
            parameter.copy_(torch.randn(parameter.shape, generator=generator) * 0.3)
    return module


def to_numpy(arrays) -> list[np.ndarray]:
    """Copies: what a client returns may alias its live parameters."""
    return [(a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)).copy() for a in arrays]


def same_payload(a, b, what, tol: float = 1e-5) -> None:
    a, b = to_numpy(a), to_numpy(b)
    assert len(a) == len(b), (what, len(a), len(b))
    for x, y in zip(a, b):
        assert x.shape == y.shape, (what, x.shape, y.shape)
        if x.dtype.kind in "US":
            assert (x == y).all(), what
        else:
            assert np.allclose(x.astype(np.float64), y.astype(np.float64), atol=tol), (what, np.abs(x.astype(np.float64) - y.astype(np.float64)).max())


def fedpm_clients(already_masked: bool):
    built = []
    for prefix in ("fl4health", "fl4health_b200"):
        side = resolver(prefix)
        masked = side("model_bases.masked_layers.masked_linear").MaskedLinear

        def model_factory(masked=masked):
            if already_masked:
                return nn.Sequential(masked(10, 16), nn.ReLU(), masked(16, 2))
            return nn.Sequential(nn.Linear(10, 16), nn.ReLU(), nn.Linear(16, 2))

        hooks = user_hooks(side, 0, lr=0.1)
        hooks["get_model"] = lambda self, config, factory=model_factory: pin_everything(factory()).to(self.device)
        if not already_masked:  # conversion re-initialises the scores: pin them right after set-up
            base_setup = side("clients.fedpm_client").FedPmClient.setup_client

            def setup_client(self, config, base_setup=base_setup):
                base_setup(self, config)
                pin_everything(self.model)

            hooks["setup_client"] = setup_client
        cls = type("PmClient", (side("clients.fedpm_client").FedPmClient,), hooks)
        accuracy = side("metrics").Accuracy
        built.append(cls(data_path=Path("."), metrics=[accuracy()], device=torch.device("cpu"), client_name="client_0"))
    return built


# (only models that are masked from the start: when the reference converts a plain model inside ``setup_client`` its
# optimizer already holds the unconverted model's parameters, so the new scores are never trained there)
for already_masked in (True,):
    theirs, ours = fedpm_clients(already_masked)
    config = {"current_server_round": 1, "local_steps": 4, "batch_size": 16, "is_masked_model": already_masked}
    torch.manual_seed(100); initial_ref = theirs.get_parameters(dict(config, current_server_round=0))
    torch.manual_seed(100); initial_mine = ours.get_parameters(dict(config, current_server_round=0))
    same_payload(initial_ref, initial_mine, "initial parameters")
    payload = to_numpy(initial_ref)
    for server_round in (1, 2, 3):
        config["current_server_round"] = server_round
        torch.manual_seed(200 + server_round); out_ref, n_ref, metrics_ref = theirs.fit([p.copy() for p in payload], dict(config))
        torch.manual_seed(200 + server_round); out_mine, n_mine, metrics_mine = ours.fit([p.copy() for p in payload], dict(config))
        # the scores learnt in this round (every forward sampled the same masks on both sides)
        for (name, a), (_, b) in zip(theirs.model.state_dict().items(), ours.model.state_dict().items()):
            assert torch.allclose(a, b, atol=1e-6), (server_round, name, (a - b).abs().max())
        # what travels: binary masks + the names of the score tensors.  The masks themselves are independent draws (the
        # reference samples them with scipy / NumPy, we sample on the device), so structure and rates are compared
        ref_arrays, my_arrays = to_numpy(out_ref), to_numpy(out_mine)
        assert len(ref_arrays) == len(my_arrays) and [a.shape for a in ref_arrays] == [a.shape for a in my_arrays]
        assert (ref_arrays[-1] == my_arrays[-1]).all()  # layer names
        for mask_ref, mask_mine in zip(ref_arrays[:-1], my_arrays[:-1]):
            assert set(np.unique(mask_ref)) <= {0, 1} and set(np.unique(mask_mine)) <= {0, 1}
        rate_ref = np.concatenate([m.reshape(-1) for m in ref_arrays[:-1]]).mean()
        rate_mine = np.concatenate([m.reshape(-1).astype(np.float64) for m in my_arrays[:-1]]).mean()
        assert abs(rate_ref - rate_mine) < 0.15, (rate_ref, rate_mine)
        assert n_ref == n_mine and metrics_ref.keys() == metrics_mine.keys()
        assert all(abs(float(metrics_ref[k]) - float(metrics_mine[k])) < 1e-5 for k in metrics_ref), (metrics_ref, metrics_mine)
        # the server would answer with per-parameter probabilities: feed both the same "aggregate" (the masks themselves)
        torch.manual_seed(300 + server_round); loss_ref, _, eval_ref = theirs.evaluate(to_numpy(out_ref), dict(config))
        torch.manual_seed(300 + server_round); loss_mine, _, eval_mine = ours.evaluate(to_numpy(out_ref), dict(config))
        assert abs(loss_ref - loss_mine) < 1e-5 and all(abs(float(eval_ref[k]) - float(eval_mine[k])) < 1e-5 for k in eval_ref)
        payload = to_numpy(out_ref)
    agreed += 1


# -- Ditto / MR-MTL with a Deep-MMD feature-alignment term: the deep kernel is trained every few steps on shuffled samples ------
def mmd_clients(module_path: str, class_name: str, optimizers: bool):
    built = []
    for prefix in ("fl4health", "fl4health_b200"):
        side = resolver(prefix)
        hooks = user_hooks(side, 1)
        if optimizers:
            hooks["get_optimizer"] = lambda self, config: {"global": torch.optim.SGD(self.global_model.parameters(), lr=0.05, momentum=0.9),
                                                           "local": torch.optim.SGD(self.model.parameters(), lr=0.05, momentum=0.9)}
        cls = type("MmdClient", (getattr(side(module_path), class_name),), hooks)
        torch.manual_seed(55); np.random.seed(55)
        client = cls(data_path=Path("."), metrics=[side("metrics").Accuracy()], device=torch.device("cpu"), client_name="client_0",
                     deep_mmd_loss_weight=2.0, feature_extraction_layers_with_size={"body": 16}, mmd_kernel_train_interval=2, num_accumulating_batches=2)
        built.append(client)
    theirs, ours = built
    for layer, loss in theirs.deep_mmd_losses.items():  # same deep kernel on both sides
        ours.deep_mmd_losses[layer].featurizer.load_state_dict(loss.featurizer.state_dict())
        for name in ("epsilon_opt", "sigma_q_opt", "sigma_phi_opt"):
            getattr(ours.deep_mmd_losses[layer], name).data.copy_(getattr(loss, name).data)
    return theirs, ours


for module_path, class_name, two_optimizers, packs_weight in (
    ("clients.deep_mmd_clients.ditto_deep_mmd_client", "DittoDeepMmdClient", True, True),
    ("clients.deep_mmd_clients.mr_mtl_deep_mmd_client", "MrMtlDeepMmdClient", False, True),
):
    theirs, ours = mmd_clients(module_path, class_name, two_optimizers)
    config = {"current_server_round": 1, "local_steps": 6, "batch_size": 16}
    torch.manual_seed(100); initial_ref = to_numpy(theirs.get_parameters(dict(config, current_server_round=0)))
    torch.manual_seed(100); initial_mine = to_numpy(ours.get_parameters(dict(config, current_server_round=0)))
    same_payload(initial_ref, initial_mine, "initial parameters")
    payload = initial_ref + [np.array(0.5)]  # the server appends the drift-penalty weight
    for server_round in (1, 2):
        config["current_server_round"] = server_round
        torch.manual_seed(400 + server_round); np.random.seed(server_round); out_ref, _, metrics_ref = theirs.fit([p.copy() for p in payload], dict(config))
        torch.manual_seed(400 + server_round); np.random.seed(server_round); out_mine, _, metrics_mine = ours.fit([p.copy() for p in payload], dict(config))
        same_payload(to_numpy(out_ref)[:-1], to_numpy(out_mine)[:-1], f"{class_name} round {server_round} weights", tol=2e-4)  # the deep kernel trains in fp64 there, fp32 here
        assert abs(float(to_numpy(out_ref)[-1]) - float(to_numpy(out_mine)[-1])) < 1e-4  # the packed training loss
        assert metrics_ref.keys() == metrics_mine.keys() and all(abs(float(metrics_ref[k]) - float(metrics_mine[k])) < 1e-4 for k in metrics_ref), (metrics_ref, metrics_mine)
        payload = to_numpy(out_ref)[:-1] + [np.array(0.5)]
    agreed += 1


# -- client-level DP: the clipping client (update = new - old weights, clipped to the bound the server sent; clipping bit) ---
for adaptive, bound in ((True, 0.05), (True, 50.0), (False, 0.05)):
    built = []
    for prefix in ("fl4health", "fl4health_b200"):
        side = resolver(prefix)
        cls = type("ClipClient", (side("clients.clipping_client").NumpyClippingClient,), user_hooks(side, 2))
        built.append(cls(data_path=Path("."), metrics=[side("metrics").Accuracy()], device=torch.device("cpu"), client_name="client_0"))
    theirs, ours = built
    config = {"current_server_round": 1, "local_steps": 4, "batch_size": 16, "adaptive_clipping": adaptive}
    initial = to_numpy(theirs.get_parameters(dict(config, current_server_round=0)))
    same_payload(initial, to_numpy(ours.get_parameters(dict(config, current_server_round=0))), "initial parameters")
    payload = initial + [np.array(bound)]
    for server_round in (1, 2, 3):
        config["current_server_round"] = server_round
        out_ref, n_ref, _ = theirs.fit([p.copy() for p in payload], dict(config))
        out_mine, n_mine, _ = ours.fit([p.copy() for p in payload], dict(config))
        ref_arrays, my_arrays = to_numpy(out_ref), to_numpy(out_mine)
        if server_round == 1:
            # On a CPU device the reference's round-1 "initial weights" are NumPy views of the live parameters
            # (``push_parameters`` -> ``.cpu().numpy()`` shares storage), so its first update is identically zero; on a GPU
            # the same line copies.  Ours is the true update; the comparison starts with round 2.
            assert all(not a.any() for a in ref_arrays[:-1]) and any(a.any() for a in my_arrays[:-1])
            payload = [w + u for w, u in zip(payload[:-1], my_arrays[:-1])] + [np.array(bound)]
            continue
        same_payload(ref_arrays[:-1], my_arrays[:-1], f"clipped update (bound {bound})", tol=1e-6)
        assert float(np.asarray(ref_arrays[-1]).reshape(-1)[0]) == float(np.asarray(my_arrays[-1]).reshape(-1)[0]) and n_ref == n_mine, (ref_arrays[-1], my_arrays[-1])  # the clipping bit
        norm = np.sqrt(sum(float((a.astype(np.float64) ** 2).sum()) for a in my_arrays[:-1]))
        assert norm <= bound * (1 + 1e-5)
        loss_ref, _, eval_ref = theirs.evaluate([p.copy() for p in payload], dict(config))
        loss_mine, _, eval_mine = ours.evaluate([p.copy() for p in payload], dict(config))
        assert abs(loss_ref - loss_mine) < 1e-5 and all(abs(float(eval_ref[k]) - float(eval_mine[k])) < 1e-6 for k in eval_ref)
        payload = [w + u for w, u in zip(payload[:-1], ref_arrays[:-1])] + [np.array(bound)]  # the server applies the update
    agreed += 1

# -- FedPCA client: local principal components, then evaluation of merged components ---------------------------------------------
import tempfile  # noqa: E402

from torch.utils.data import DataLoader  # noqa: E402

for low_rank, full_svd in ((False, True), (True, False)):
    built = []
    for prefix in ("fl4health", "fl4health_b200"):
        side = resolver(prefix)
        dataset_module = side("utils.dataset")

        def get_data_loaders(self, config, dataset_module=dataset_module):
            generator = torch.Generator().manual_seed(77)
            data = torch.randn(90, 12, generator=generator) @ torch.randn(12, 12, generator=generator)
            labels = torch.zeros(90).long()
            return (DataLoader(dataset_module.TensorDataset(data[:60], labels[:60]), batch_size=20), DataLoader(dataset_module.TensorDataset(data[60:], labels[60:]), batch_size=20))

        def get_data_tensor(self, data_loader):
            return torch.cat([batch for batch, _ in data_loader], dim=0)

        cls = type("PcaClient", (side("clients.fed_pca_client").FedPCAClient,), {"get_data_loaders": get_data_loaders, "get_data_tensor": get_data_tensor})
        built.append(cls(Path("."), torch.device("cpu"), Path(tempfile.mkdtemp()), client_name="client_0"))
    theirs, ours = built
    config = {"current_server_round": 1, "low_rank": low_rank, "full_svd": full_svd, "rank_estimation": 6, "center_data": True, "num_components_eval": 4}
    torch.manual_seed(5); out_ref, n_ref, metrics_ref = theirs.fit([], dict(config))
    torch.manual_seed(5); out_mine, n_mine, metrics_mine = ours.fit([], dict(config))
    (pc_ref, sv_ref), (pc_mine, sv_mine) = to_numpy(out_ref), to_numpy(out_mine)
    k = 4
    assert n_ref == n_mine and pc_ref.shape == pc_mine.shape and np.allclose(sv_ref[:k], sv_mine[:k], rtol=1e-3, atol=1e-3), (sv_ref, sv_mine)
    assert np.allclose(np.abs((pc_ref[:, :k] * pc_mine[:, :k]).sum(axis=0)), 1.0, atol=1e-3)  # same directions up to sign
    assert metrics_ref.keys() == metrics_mine.keys() and all(abs(float(metrics_ref[m]) - float(metrics_mine[m])) < 1e-3 for m in metrics_ref), (metrics_ref, metrics_mine)
    loss_ref, _, eval_ref = theirs.evaluate([pc_ref.copy(), sv_ref.copy()], dict(config))
    loss_mine, _, eval_mine = ours.evaluate([pc_ref.copy(), sv_ref.copy()], dict(config))
    assert abs(loss_ref - loss_mine) < 1e-3 * max(1.0, abs(loss_ref)), (loss_ref, loss_mine)
    assert all(abs(float(eval_ref[m]) - float(eval_mine[m])) < 1e-3 * max(1.0, abs(float(eval_ref[m]))) for m in eval_ref), (eval_ref, eval_mine)
    agreed += 1
print("configs agree:", agreed)
