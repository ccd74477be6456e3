"""Closed-form oracles for FedPM, FLASH, dynamic-layer and sparse-COO aggregation (values taken from the reference's
tests/strategies/test_fedpm.py, test_flash.py, test_fedavg_dynamic_layer.py)."""

import numpy as np
import torch

from fl4health_b200.common.typing import Code, FitRes, NDArrays, Status, ndarrays_to_parameters, parameters_to_ndarrays, to_numpy
from fl4health_b200.parameter_exchange.parameter_packer import SparseCooParameterPacker
from fl4health_b200.servers.client_proxy import InProcessClientProxy
from fl4health_b200.strategies.fedavg_dynamic_layer import FedAvgDynamicLayer
from fl4health_b200.strategies.fedavg_sparse_coo_tensor import FedAvgSparseCooTensor
from fl4health_b200.strategies.fedpm import FedPm
from fl4health_b200.strategies.flash import Flash

client0 = [np.identity(3), np.ones((4, 4))] + [np.array(["layer1", "layer2"])]
client1 = [np.ones((4, 4)), np.zeros((5, 5))] + [np.array(["layer2", "layer3"])]
client2 = [np.ones((3, 3)), np.identity(5)] + [np.array(["layer1", "layer3"])]
client3 = [np.zeros((4, 4)), np.ones((6, 6))] + [np.array(["layer2", "layer4"])]
FIT = [(NDArrays(c), n) for c, n in zip([client0, client1, client2, client3], [50, 50, 100, 200])]


def test_fedpm_bayesian_aggregation() -> None:
    strategy = FedPm()
    for _ in range(2):  # posterior mode is stable when the same evidence repeats
        out = strategy.aggregate_bayesian(FIT)
        l1 = np.full((3, 3), 0.5)
        np.fill_diagonal(l1, 1.0)
        l3 = np.zeros((5, 5))
        np.fill_diagonal(l3, 0.5)
        expected = {"layer1": l1, "layer2": np.full((4, 4), 2 / 3), "layer3": l3, "layer4": np.ones((6, 6))}
        assert out.keys() == expected.keys()
        for name, value in expected.items():
            assert np.allclose(to_numpy(out[name]), value, atol=1e-6), name
    strategy2 = FedPm()
    strategy2.aggregate_bayesian(FIT)
    alpha, beta = strategy2.beta_parameters["layer2"]
    assert (to_numpy(alpha) == 3).all() and (to_numpy(beta) == 2).all()
    alpha, beta = strategy2.beta_parameters["layer4"]
    assert (to_numpy(alpha) == 2).all() and (to_numpy(beta) == 1).all()
    strategy2.reset_beta_priors()
    assert all((to_numpy(a) == 1).all() and (to_numpy(b) == 1).all() for a, b in strategy2.beta_parameters.values())


def test_dynamic_layer_weighted_and_uniform() -> None:
    weighted = FedAvgDynamicLayer(weighted_aggregation=True).aggregate(FIT)
    # layer1: identity (n=50) and ones (n=100) -> (50 I + 100) / 150
    assert np.allclose(to_numpy(weighted["layer1"]), (50 * np.identity(3) + 100 * np.ones((3, 3))) / 150)
    # layer2: ones(50), ones(50), zeros(200) -> 100/300
    assert np.allclose(to_numpy(weighted["layer2"]), np.full((4, 4), 1 / 3))
    uniform = FedAvgDynamicLayer(weighted_aggregation=False).aggregate(FIT)
    assert np.allclose(to_numpy(uniform["layer2"]), np.full((4, 4), 2 / 3))
    assert np.allclose(to_numpy(uniform["layer4"]), np.ones((6, 6)))


def test_sparse_coo_aggregation_round_trip() -> None:
    packer = SparseCooParameterPacker()

    def payload(dense: dict[str, torch.Tensor]) -> NDArrays:
        vals, idx, shapes, names = NDArrays(), NDArrays(), NDArrays(), []
        for name, tensor in dense.items():
            v, i, s = packer.extract_coo_info_from_dense(tensor)
            vals.append(v), idx.append(i), shapes.append(s), names.append(name)
        return packer.pack_parameters(vals, (idx, shapes, names))

    a = {"w": torch.tensor([[1.0, 0.0], [0.0, 2.0]]), "b": torch.tensor([0.0, 4.0])}
    b = {"w": torch.tensor([[3.0, 0.0], [5.0, 0.0]])}
    out = FedAvgSparseCooTensor(weighted_aggregation=True).aggregate([(payload(a), 1), (payload(b), 3)])
    assert torch.allclose(out["w"], torch.tensor([[(1 + 9) / 4, 0.0], [15 / 4, 2 / 4]]))
    assert torch.allclose(out["b"], torch.tensor([0.0, 4.0]))
    out_u = FedAvgSparseCooTensor(weighted_aggregation=False).aggregate([(payload(a), 1), (payload(b), 3)])
    assert torch.allclose(out_u["w"], torch.tensor([[2.0, 0.0], [2.5, 1.0]]))


def _fit_results(layers_per_client):
    return [(InProcessClientProxy(f"c{i}", None), FitRes(Status(Code.OK), ndarrays_to_parameters(NDArrays(layers)), 1, {}))
            for i, layers in enumerate(layers_per_client)]


def test_flash_server_step() -> None:
    init = [np.zeros((3, 3)), np.zeros((4, 4))]
    flash = Flash(initial_parameters=ndarrays_to_parameters(NDArrays(init)), eta=0.1, eta_l=0.1, beta_1=0.9, beta_2=0.99,
                  tau=1e-9, min_fit_clients=2, min_available_clients=2)
    clients = [[np.ones((3, 3)) * 2.0, np.ones((4, 4)) * 2.0], [np.ones((3, 3)) * 2.5, np.ones((4, 4)) * 3.0]]
    params, _ = flash.aggregate_fit(1, _fit_results(clients), [])
    exp_m = [0.1 * np.ones((3, 3)) * 2.25, 0.1 * np.ones((4, 4)) * 2.5]
    exp_v = [0.01 * np.ones((3, 3)) * 2.25**2, 0.01 * np.ones((4, 4)) * 2.5**2]
    exp_d = [np.ones((3, 3)) * (2.25**2 - 0.050625), np.ones((4, 4)) * (2.5**2 - 0.0625)]
    for got, exp in zip(flash.m_t, exp_m):
        assert np.allclose(to_numpy(got), exp)
    for got, exp in zip(flash.v_t, exp_v):
        assert np.allclose(to_numpy(got), exp)
    for got, exp in zip(flash.d_t, exp_d):
        assert np.allclose(to_numpy(got), exp)
    new = [0.1 * m / (np.sqrt(v) - d + 1e-9) for m, v, d in zip(exp_m, exp_v, exp_d)]
    for got, exp in zip(parameters_to_ndarrays(params), new):
        assert np.allclose(to_numpy(got), exp, rtol=1e-5)


def test_fedpm_bit_packed_cross_rank_vote_matches_per_tensor_vote() -> None:
    """One client per rank: masks are packed to 1 bit / score, all-gathered as words and voted in one launch for every
    layer.  Same posterior as the per-tensor uint8 vote over materialised payloads, over two rounds (evidence carries)."""
    import numpy as np
    import torch

    from fl4health_b200.common.typing import Code, FitRes, NDArrays, Status
    from fl4health_b200.ops import flat as flat_ops
    from fl4health_b200.parallel.spmd import PayloadSpec, RemoteNDArrays, _TaggedParameters, _tag_local
    from fl4health_b200.parameter_exchange.parameter_packer import ParameterPackerWithLayerNames
    from fl4health_b200.servers.client_proxy import ClientProxy
    from fl4health_b200.strategies.fedpm import FedPm

    # pack / unpack round trip on an awkward length
    gen = torch.Generator().manual_seed(0)
    bits = (torch.rand(1000 + 13, generator=gen) < 0.4).to(torch.uint8)
    words = flat_ops.pack_mask_bits(bits)
    assert words.dtype == torch.int32 and words.numel() == 32 and torch.equal(flat_ops.unpack_mask_bits(words, bits.numel()), bits)
    assert torch.equal(flat_ops.pack_mask_bits(bits.float()), words) and torch.equal(flat_ops.pack_mask_bits(bits.bool()), words)

    names, shapes = ["conv.weight_scores", "conv.bias_scores", "fc.weight_scores"], [(4, 3, 3, 3), (4,), (10, 37)]
    packer = ParameterPackerWithLayerNames()

    class Proxy(ClientProxy):
        get_properties = get_parameters = fit = evaluate = reconnect = None  # type: ignore[assignment]

    def masks_of(round_index: int, client: int) -> NDArrays:
        g = torch.Generator().manual_seed(100 * round_index + client)
        return NDArrays([(torch.rand(shape, generator=g) < 0.3 + 0.2 * client).to(torch.uint8) for shape in shapes])

    class TwoRankWorld:  # what rank 0 of a two-rank federation sees
        world_size, rank, device = 2, 0, torch.device("cpu")

        def __init__(self) -> None:
            self.other: torch.Tensor | None = None
            self.gathers = 0

        def all_gather_rows(self, row: torch.Tensor) -> torch.Tensor:
            self.gathers += 1
            return torch.stack([row, self.other])

    world = TwoRankWorld()
    packed_strategy, plain_strategy = FedPm(), FedPm()
    for round_index in (1, 2):
        payloads = [packer.pack_parameters(masks_of(round_index, c), names) for c in range(2)]
        world.other = flat_ops.pack_mask_bits(torch.cat([m.reshape(-1) for m in masks_of(round_index, 1)]))
        mine = _tag_local(payloads[0], world)
        theirs = RemoteNDArrays(world, 1, PayloadSpec.of(payloads[1]))  # tensors never arrive: only the spec is known
        assert all(entry is None for entry in list(theirs)[:-1])
        spmd_results = [(Proxy(cid=f"rank{r:03d}"), FitRes(Status(Code.OK, ""), _TaggedParameters(p), 10, {})) for r, p in enumerate((mine, theirs))]
        plain_results = [(Proxy(cid=f"rank{r:03d}"), FitRes(Status(Code.OK, ""), _TaggedParameters(NDArrays(p)), 10, {})) for r, p in enumerate(payloads)]
        voted, _ = packed_strategy.aggregate_fit(round_index, spmd_results, [])
        expected, _ = plain_strategy.aggregate_fit(round_index, plain_results, [])
        assert packed_strategy.last_vote_path == "packed-bits" and plain_strategy.last_vote_path == "per-tensor"
        got_layers, got_names = packer.unpack_parameters(NDArrays(voted.tensors))
        want_layers, want_names = packer.unpack_parameters(NDArrays(expected.tensors))
        assert list(got_names) == list(want_names) == names
        for got, want in zip(got_layers, want_layers):
            assert got.shape == want.shape and torch.allclose(torch.as_tensor(got), torch.as_tensor(want), atol=1e-7)
    assert world.gathers == 2  # one collective per round, whatever the number of layers
    for name in names:
        for flat_prior, plain_prior in zip(packed_strategy.beta_parameters[name], plain_strategy.beta_parameters[name]):
            assert torch.equal(flat_prior, plain_prior)
    packed_strategy.reset_beta_priors()
    assert float(packed_strategy._flat_priors[1].min()) == float(packed_strategy._flat_priors[2].max()) == 1.0
    # the uniform-mean variant takes the same route
    mean_strategy = FedPm(bayesian_aggregation=False)
    voted, _ = mean_strategy.aggregate_fit(3, spmd_results, [])
    layers, _ = packer.unpack_parameters(NDArrays(voted.tensors))
    assert mean_strategy.last_vote_path == "packed-bits" and set(np.unique(torch.as_tensor(layers[2]).numpy())) <= {0.0, 0.5, 1.0}
