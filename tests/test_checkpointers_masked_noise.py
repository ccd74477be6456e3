"""Unit tests modelled on the reference's ``tests/checkpointing/test_function_checkpointer.py``,
``tests/checkpointing/test_opacus_checkpointers.py``, ``tests/models/test_masked_layers.py``,
``tests/models/test_feature_extractor_buffer.py``, ``tests/strategies/test_noisy_aggregation.py``,
``tests/servers/test_polling.py`` and ``tests/feature_alignment/test_string_columns_transformer.py``."""

from __future__ import annotations

import math

import pandas as pd
import pytest
import torch
from torch import nn

from fl4health_b200.checkpointing.checkpointer import (
    BestLossTorchModuleCheckpointer,
    BestMetricTorchModuleCheckpointer,
    FunctionTorchModuleCheckpointer,
    LatestTorchModuleCheckpointer,
)
from fl4health_b200.checkpointing.opacus_checkpointer import BestLossOpacusCheckpointer, LatestOpacusCheckpointer
from fl4health_b200.common.typing import Code, GetPropertiesIns, GetPropertiesRes, NDArrays, Status, to_tensor
from fl4health_b200.feature_alignment.string_columns_transformer import TextColumnTransformer, TextMulticolumnTransformer
from fl4health_b200.model_bases.feature_extractor_buffer import FeatureExtractorBuffer
from fl4health_b200.model_bases.masked_layers.masked_layers import (
    MaskedBatchNorm2d,
    MaskedConv2d,
    MaskedConvTranspose2d,
    MaskedLayerNorm,
    MaskedLinear,
)
from fl4health_b200.model_bases.masked_layers.masked_layers_utils import convert_to_masked_model, is_masked_module
from fl4health_b200.privacy.dp_engine import GradSampleModule
from fl4health_b200.servers.client_proxy import ClientProxy
from fl4health_b200.servers.polling import poll_client, poll_clients
from fl4health_b200.strategies.basic_fedavg import OpacusBasicFedAvg
from fl4health_b200.strategies.noisy_aggregate import (
    add_noise_to_array,
    gaussian_noisy_aggregate_clipping_bits,
    gaussian_noisy_unweighted_aggregate,
    gaussian_noisy_weighted_aggregate,
)


class TinyNet(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.conv = nn.Conv2d(1, 4, 3, padding=1)
        self.norm = nn.GroupNorm(2, 4)
        self.fc = nn.Linear(4 * 6 * 6, 3)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.fc(torch.relu(self.norm(self.conv(x))).flatten(1))


# ---------------------------------------------------------------------------------------------- module checkpointers
def test_function_checkpointer_minimise_and_maximise(tmp_path) -> None:
    def score(loss: float, metrics: dict) -> float:
        return loss + float(metrics["penalty"])

    model = TinyNet()
    minimiser = FunctionTorchModuleCheckpointer(str(tmp_path), "min.pt", score, "loss+penalty", maximize=False)
    minimiser.maybe_checkpoint(model, 1.0, {"penalty": 0.5})
    assert minimiser.best_score == pytest.approx(1.5)
    with torch.no_grad():
        model.fc.bias.fill_(7.0)
    minimiser.maybe_checkpoint(model, 1.2, {"penalty": 0.5})  # worse: file must keep the first model
    assert minimiser.best_score == pytest.approx(1.5)
    assert not torch.allclose(minimiser.load_checkpoint().fc.bias, torch.full((3,), 7.0))
    minimiser.maybe_checkpoint(model, 0.2, {"penalty": 0.1})
    assert minimiser.best_score == pytest.approx(0.3)
    assert torch.allclose(minimiser.load_checkpoint().fc.bias, torch.full((3,), 7.0))

    maximiser = FunctionTorchModuleCheckpointer(str(tmp_path), "max.pt", score, maximize=True)
    assert maximiser.checkpoint_score_name == "score"  # defaults to the function's __name__
    maximiser.maybe_checkpoint(model, 0.0, {"penalty": 0.0})  # a best score of 0.0 is a real score, not "unset"
    maximiser.maybe_checkpoint(model, -1.0, {"penalty": 0.0})
    assert maximiser.best_score == 0.0
    maximiser.maybe_checkpoint(model, 2.0, {"penalty": 0.0})
    assert maximiser.best_score == 2.0


def test_best_metric_and_latest_checkpointers(tmp_path) -> None:
    model = TinyNet()
    best = BestMetricTorchModuleCheckpointer(str(tmp_path), "best.pt", metric="accuracy", maximize=True)
    best.maybe_checkpoint(model, 9.0, {"val - prediction - accuracy": 0.4})
    best.maybe_checkpoint(model, 0.1, {"val - prediction - accuracy": 0.3})
    assert best.best_score == pytest.approx(0.4)
    with pytest.raises(KeyError):
        best.maybe_checkpoint(model, 0.1, {"accuracy": 0.9})  # prefix is part of the key

    latest = LatestTorchModuleCheckpointer(str(tmp_path), "latest.pt")
    for value in (1.0, 2.0):
        with torch.no_grad():
            model.fc.bias.fill_(value)
        latest.maybe_checkpoint(model, 100.0 * value, {})
    assert torch.allclose(latest.load_checkpoint().fc.bias, torch.full((3,), 2.0))

    loss_based = BestLossTorchModuleCheckpointer(str(tmp_path), "loss.pt")
    loss_based.maybe_checkpoint(model, 0.7, {})
    loss_based.maybe_checkpoint(model, 0.9, {})
    assert loss_based.best_score == pytest.approx(0.7)


def test_opacus_checkpointers_round_trip_both_key_conventions(tmp_path) -> None:
    wrapped = GradSampleModule(TinyNet())
    checkpointer = BestLossOpacusCheckpointer(str(tmp_path), "dp.pkl")
    checkpointer.maybe_checkpoint(wrapped, 0.5, {})
    with torch.no_grad():
        for p in wrapped.parameters():
            p.add_(1.0)
    checkpointer.maybe_checkpoint(wrapped, 0.9, {})  # worse loss: not saved
    with pytest.raises(NotImplementedError):
        checkpointer.load_checkpoint()

    plain = TinyNet()
    checkpointer.load_best_checkpoint_into_model(plain, target_is_grad_sample_module=False)
    rewrapped = GradSampleModule(TinyNet())
    checkpointer.load_best_checkpoint_into_model(rewrapped, target_is_grad_sample_module=True)
    for (_, a), (_, b), (_, c) in zip(plain.state_dict().items(), rewrapped.state_dict().items(), wrapped.state_dict().items()):
        assert torch.equal(a, b)
        assert torch.allclose(a + 1.0, c)  # the saved state predates the +1 edit

    latest = LatestOpacusCheckpointer(str(tmp_path), "dp_latest.pkl")
    latest.maybe_checkpoint(wrapped, 123.0, {})
    latest.load_best_checkpoint_into_model(plain)
    assert all(torch.equal(a, c) for a, c in zip(plain.state_dict().values(), wrapped.state_dict().values()))


def test_opacus_basic_fedavg_takes_wrapped_initial_parameters() -> None:
    wrapped = GradSampleModule(TinyNet())
    strategy = OpacusBasicFedAvg(model=wrapped, min_fit_clients=1, min_evaluate_clients=1, min_available_clients=1)
    initial = strategy.initialize_parameters(None)
    assert initial is not None and len(initial.tensors) == len(wrapped.state_dict())
    with pytest.raises(AssertionError):
        OpacusBasicFedAvg(model=TinyNet(), min_fit_clients=1, min_evaluate_clients=1, min_available_clients=1)


# ------------------------------------------------------------------------------------------------------ masked layers
@pytest.mark.parametrize(
    "stock, masked_type, shape",
    [
        (nn.Linear(6, 4), MaskedLinear, (5, 6)),
        (nn.Conv2d(2, 3, 3, padding=1), MaskedConv2d, (2, 2, 5, 5)),
        (nn.ConvTranspose2d(2, 3, 2, stride=2), MaskedConvTranspose2d, (2, 2, 4, 4)),
        (nn.LayerNorm(6), MaskedLayerNorm, (3, 6)),
        (nn.BatchNorm2d(2), MaskedBatchNorm2d, (4, 2, 3, 3)),
    ],
)
def test_masked_layer_from_pretrained(stock: nn.Module, masked_type: type, shape: tuple[int, ...]) -> None:
    torch.manual_seed(3)
    masked = masked_type.from_pretrained(stock)
    assert is_masked_module(masked)
    assert torch.equal(masked.weight, stock.weight) and not masked.weight.requires_grad and not masked.bias.requires_grad
    assert masked.weight_scores.requires_grad and masked.weight_scores.shape == stock.weight.shape
    assert {n for n, p in masked.named_parameters() if p.requires_grad} == {"weight_scores", "bias_scores"}
    x = torch.randn(shape)
    # scores -> +inf: every mask bit is 1 and the layer reproduces the stock layer
    with torch.no_grad():
        masked.weight_scores.fill_(50.0)
        masked.bias_scores.fill_(50.0)
    stock.train(), masked.train()
    assert torch.allclose(masked(x), stock(x), atol=1e-5)
    # scores -> -inf: everything is masked out
    with torch.no_grad():
        masked.weight_scores.fill_(-50.0)
        masked.bias_scores.fill_(-50.0)
    out = masked(x)
    assert torch.count_nonzero(out) == 0
    # straight-through estimator: scores receive gradients, frozen weights do not
    with torch.no_grad():
        masked.weight_scores.zero_()
        masked.bias_scores.zero_()
    masked(x).square().sum().backward()
    assert masked.weight_scores.grad is not None and masked.weight.grad is None


def test_convert_to_masked_model_is_recursive_and_non_destructive() -> None:
    model = nn.Sequential(nn.Conv2d(1, 2, 3), nn.BatchNorm2d(2), nn.ReLU(), nn.Flatten(), nn.Sequential(nn.Linear(2 * 4 * 4, 5), nn.LayerNorm(5)))
    masked = convert_to_masked_model(model)
    kinds = [type(m).__name__ for m in masked.modules() if is_masked_module(m)]
    assert kinds == ["MaskedConv2d", "MaskedBatchNorm2d", "MaskedLinear", "MaskedLayerNorm"]
    assert not any(is_masked_module(m) for m in model.modules())  # the original is left alone (deep copy)
    assert torch.equal(masked[0].weight, model[0].weight)
    assert masked(torch.randn(3, 1, 6, 6)).shape == (3, 5)
    assert isinstance(convert_to_masked_model(nn.Linear(3, 2)), MaskedLinear)  # a bare layer converts too
    again = convert_to_masked_model(masked)  # idempotent
    assert [type(m).__name__ for m in again.modules() if is_masked_module(m)] == kinds


# ------------------------------------------------------------------------------------------- feature extractor buffer
def test_feature_extractor_buffer_hooks_accumulate_and_flatten() -> None:
    model = TinyNet()
    buffer = FeatureExtractorBuffer(model, {"conv": True, "fc": False})
    buffer._maybe_register_hooks()
    buffer._maybe_register_hooks()  # second call is a no-op
    assert len(buffer.fhooks) == 2
    x = torch.randn(5, 1, 6, 6)
    model(x)
    model(x[:2])
    features = buffer.get_extracted_features()  # not accumulating: only the last batch is kept
    assert features["conv"].shape == (2, 4 * 6 * 6) and features["fc"].shape == (2, 3)
    buffer.enable_accumulating_features()
    buffer.clear_buffers()
    model(x)
    model(x[:2])
    features = buffer.get_extracted_features()
    assert features["conv"].shape == (7, 4 * 6 * 6) and features["fc"].shape == (7, 3)
    assert torch.allclose(features["fc"][:5], model(x))
    buffer.disable_accumulating_features()
    buffer.remove_hooks()
    assert not buffer.fhooks
    buffer.clear_buffers()
    model(x)
    assert buffer.extracted_features_buffers == {"conv": [], "fc": []}
    with pytest.raises(ValueError):
        FeatureExtractorBuffer(model, {"nope": True})._maybe_register_hooks()


def test_feature_extractor_buffer_prefix_picks_last_matching_module() -> None:
    model = nn.Sequential()
    model.add_module("block", nn.Sequential(nn.Linear(4, 8), nn.ReLU(), nn.Linear(8, 2)))
    buffer = FeatureExtractorBuffer(model, {"block": False})
    assert buffer.find_last_common_prefix("block", [n for n, _ in model.named_modules()]) == "block.2"
    buffer._maybe_register_hooks()
    out = model(torch.randn(3, 4))
    assert torch.equal(buffer.get_extracted_features()["block"], out)


# --------------------------------------------------------------------------------------------------- noisy aggregation
def test_noisy_aggregate_without_noise_is_the_plain_mean() -> None:
    updates = [(NDArrays([torch.full((2, 2), float(k)), torch.full((3,), 2.0 * k)]), 10 * k) for k in (1, 2, 3)]
    mean = gaussian_noisy_unweighted_aggregate(updates, noise_multiplier=0.0, clipping_bound=5.0)
    assert torch.allclose(to_tensor(mean[0]), torch.full((2, 2), 2.0)) and torch.allclose(to_tensor(mean[1]), torch.full((3,), 4.0))

    # weighted: w_k = min(n_k / cap, 1); sum_k w_k Delta_k / (q W) and then the (reference's) division by K
    weighted = gaussian_noisy_weighted_aggregate(updates, 0.0, 5.0, fraction_fit=0.5, per_client_example_cap=20.0, total_client_weight=2.5)
    expected = (0.5 * 1 + 1.0 * 2 + 1.0 * 3) / (0.5 * 2.5) / 3
    assert torch.allclose(to_tensor(weighted[0]), torch.full((2, 2), expected))


def test_noise_scale_matches_sigma_over_denominator() -> None:
    torch.manual_seed(0)
    noisy = add_noise_to_array(torch.zeros(200_000), noise_std_dev=3.0, denominator=4)
    assert abs(float(noisy.std()) - 0.75) < 0.01 and abs(float(noisy.mean())) < 0.01
    updates = [(NDArrays([torch.zeros(100_000)]), 1)] * 5
    out = to_tensor(gaussian_noisy_unweighted_aggregate(updates, noise_multiplier=2.0, clipping_bound=0.5)[0])
    assert abs(float(out.std()) - 1.0 / 5) < 0.005
    weighted = to_tensor(gaussian_noisy_weighted_aggregate(updates, 2.0, 0.5, 0.5, per_client_example_cap=2.0, total_client_weight=5.0)[0])
    assert abs(float(weighted.std()) - (2.0 * 0.5 * 0.5 / 0.5) / 5) < 0.005  # sigma = z * C * max(w_k) / q


def test_clipping_bit_aggregate() -> None:
    bits = NDArrays([torch.tensor(1.0), torch.tensor(0.0), torch.tensor(1.0), torch.tensor(1.0)])
    assert gaussian_noisy_aggregate_clipping_bits(bits, 0.0) == pytest.approx(0.75)
    torch.manual_seed(1)
    draws = [gaussian_noisy_aggregate_clipping_bits(bits, 0.4) for _ in range(2000)]
    mean = sum(draws) / len(draws)
    std = math.sqrt(sum((d - mean) ** 2 for d in draws) / len(draws))
    assert abs(mean - 0.75) < 0.02 and abs(std - 0.1) < 0.01


# ------------------------------------------------------------------------------------------------------------- polling
class _CannedProxy(ClientProxy):
    def __init__(self, cid: str, n: int, fail: bool = False) -> None:
        super().__init__(cid)
        self.n, self.fail = n, fail

    def get_properties(self, ins, timeout=None, group_id=None):  # type: ignore[no-untyped-def]
        if self.fail:
            raise RuntimeError("client unreachable")
        return GetPropertiesRes(Status(Code.OK), {"num_train_samples": self.n, "echo": ins.config.get("tag", "")})

    def get_parameters(self, ins, timeout=None, group_id=None):  # type: ignore[no-untyped-def]
        raise NotImplementedError

    def fit(self, ins, timeout=None, group_id=None):  # type: ignore[no-untyped-def]
        raise NotImplementedError

    def evaluate(self, ins, timeout=None, group_id=None):  # type: ignore[no-untyped-def]
        raise NotImplementedError

    def reconnect(self, ins, timeout=None, group_id=None):  # type: ignore[no-untyped-def]
        raise NotImplementedError


def test_poll_clients_collects_results_and_failures() -> None:
    ins = GetPropertiesIns({"tag": "x"})
    proxies = [_CannedProxy("a", 10), _CannedProxy("b", 20, fail=True), _CannedProxy("c", 30)]
    proxy, res = poll_client(proxies[0], ins)
    assert proxy is proxies[0] and res.properties["num_train_samples"] == 10
    results, failures = poll_clients([(p, ins) for p in proxies], max_workers=None, timeout=None)
    assert [(p.cid, r.properties["num_train_samples"], r.properties["echo"]) for p, r in results] == [("a", 10, "x"), ("c", 30, "x")]
    assert len(failures) == 1 and isinstance(failures[0], BaseException)


# ------------------------------------------------------------------------------------------ string column transformers
class _CountingVectoriser:
    """Bag-of-words stand-in with the fit/transform surface of an sklearn text vectoriser."""

    def fit(self, texts):  # type: ignore[no-untyped-def]
        self.vocabulary = sorted({word for text in texts for word in text.split()})
        return self

    def transform(self, texts):  # type: ignore[no-untyped-def]
        return [[text.split().count(word) for word in self.vocabulary] for text in texts]


def test_string_column_transformers() -> None:
    frame = pd.DataFrame({"a": ["red fish", "blue fish"], "b": ["one", "two fish"]})
    multi = TextMulticolumnTransformer(_CountingVectoriser()).fit(frame)
    assert multi.transformer.vocabulary == ["blue", "fish", "one", "red", "two"]
    assert multi.transform(frame) == [[0, 1, 1, 1, 0], [1, 2, 0, 0, 1]]
    single = TextColumnTransformer(_CountingVectoriser()).fit(frame[["a"]])
    assert single.transform(frame[["a"]]) == [[0, 1, 1], [1, 1, 0]]
    with pytest.raises(AssertionError):
        TextColumnTransformer(_CountingVectoriser()).fit(frame)  # exactly one column
