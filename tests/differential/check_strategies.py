"""Server strategies vs the reference: the same client results in, the same aggregate (parameters, packed extras, metrics,
evaluation loss) out -- over several rounds for the strategies that keep state between rounds."""
import numpy as np
import torch

import fl4health.parameter_exchange.parameter_packer as ref_pack
import fl4health.strategies.aggregate_utils as ref_utils
import fl4health.strategies.basic_fedavg as ref_basic
import fl4health.strategies.fedavg_dynamic_layer as ref_dynamic
import fl4health.strategies.fedavg_sparse_coo_tensor as ref_sparse
import fl4health.strategies.fedavg_with_adaptive_constraint as ref_adaptive
import fl4health.strategies.fedpca as ref_pca
import fl4health.strategies.fedpm as ref_pm
import fl4health.strategies.flash as ref_flash
import fl4health.strategies.model_merge_strategy as ref_merge
import fl4health.strategies.scaffold as ref_scaffold
import fl4health_b200.parameter_exchange.parameter_packer as my_pack
import fl4health_b200.strategies.aggregate_utils as my_utils
import fl4health_b200.strategies.basic_fedavg as my_basic
import fl4health_b200.strategies.fedavg_dynamic_layer as my_dynamic
import fl4health_b200.strategies.fedavg_sparse_coo_tensor as my_sparse
import fl4health_b200.strategies.fedavg_with_adaptive_constraint as my_adaptive
import fl4health_b200.strategies.fedpca as my_pca
import fl4health_b200.strategies.fedpm as my_pm
import fl4health_b200.strategies.flash as my_flash
import fl4health_b200.strategies.model_merge_strategy as my_merge
import fl4health_b200.strategies.scaffold as my_scaffold
from fl4health.metrics.metric_aggregation import evaluate_metrics_aggregation_fn as ref_eval_fn
from fl4health.metrics.metric_aggregation import fit_metrics_aggregation_fn as ref_fit_fn
from fl4health_b200.metrics.metric_aggregation import evaluate_metrics_aggregation_fn as my_eval_fn
from fl4health_b200.metrics.metric_aggregation import fit_metrics_aggregation_fn as my_fit_fn

import flwr.common as fc
import fl4health_b200.common.typing as mt

rng = np.random.default_rng(17)
agreed = 0
SHAPES = [(4, 3), (4,), (2, 4), (2,)]


class Proxy:
    def __init__(self, cid: str) -> None:
        self.cid = cid


def fit_results(payloads, counts, metrics=None):
    """The same payloads wrapped as the reference's (flwr) and as our result types."""
    metrics = metrics or [{"train - loss": float(i + 1), "acc": 0.5 + 0.1 * i} for i in range(len(payloads))]
    theirs = [(Proxy(f"c{i}"), fc.FitRes(fc.Status(fc.Code.OK, ""), fc.ndarrays_to_parameters(p), n, dict(m)))
              for i, (p, n, m) in enumerate(zip(payloads, counts, metrics))]
    ours = [(Proxy(f"c{i}"), mt.FitRes(mt.Status(mt.Code.OK, ""), mt.ndarrays_to_parameters(p), n, dict(m)))
            for i, (p, n, m) in enumerate(zip(payloads, counts, metrics))]
    return theirs, ours


def eval_results(losses, counts):
    metrics = [{"val - acc": 0.4 + 0.1 * i} for i in range(len(losses))]
    theirs = [(Proxy(f"c{i}"), fc.EvaluateRes(fc.Status(fc.Code.OK, ""), loss, n, dict(m))) for i, (loss, n, m) in enumerate(zip(losses, counts, metrics))]
    ours = [(Proxy(f"c{i}"), mt.EvaluateRes(mt.Status(mt.Code.OK, ""), loss, n, dict(m))) for i, (loss, n, m) in enumerate(zip(losses, counts, metrics))]
    return theirs, ours


def arrays(parameters, ours: bool):
    out = (mt if ours else fc).parameters_to_ndarrays(parameters)
    return [a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a) for a in out]


def same_arrays(a, b, tol=1e-5) -> None:
    assert len(a) == len(b), (len(a), len(b))
    for x, y in zip(a, b):
        x, y = np.asarray(x), np.asarray(y)
        assert x.shape == y.shape, (x.shape, y.shape)
        if x.dtype.kind in "US":
            assert (x == y).all()
        else:
            assert np.allclose(x.astype(np.float64), y.astype(np.float64), atol=tol, rtol=tol), np.abs(x - y).max()


def same_metrics(a, b) -> None:
    assert a.keys() == b.keys(), (a, b)
    assert all(abs(float(a[k]) - float(b[k])) < 1e-6 for k in a), (a, b)


def common(ours: bool) -> dict:
    return {"min_fit_clients": 3, "min_evaluate_clients": 3, "min_available_clients": 3,
            "fit_metrics_aggregation_fn": my_fit_fn if ours else ref_fit_fn,
            "evaluate_metrics_aggregation_fn": my_eval_fn if ours else ref_eval_fn}


def weights() -> list[np.ndarray]:
    return [rng.normal(size=s).astype(np.float32) for s in SHAPES]


def run(theirs, ours, payload_rounds, counts, tol=1e-5) -> None:
    """Feed the same rounds to both strategies and compare everything they return."""
    global agreed
    for server_round, payloads in enumerate(payload_rounds, start=1):
        res_ref, res_mine = fit_results(payloads, counts)
        (p_ref, m_ref), (p_mine, m_mine) = theirs.aggregate_fit(server_round, res_ref, []), ours.aggregate_fit(server_round, res_mine, [])
        same_arrays(arrays(p_ref, False), arrays(p_mine, True), tol)
        same_metrics(m_ref, m_mine)
        ev_ref, ev_mine = eval_results([1.0, 0.5, 0.25], counts)
        (l_ref, em_ref), (l_mine, em_mine) = theirs.aggregate_evaluate(server_round, ev_ref, []), ours.aggregate_evaluate(server_round, ev_mine, [])
        assert abs(l_ref - l_mine) < 1e-6
        same_metrics(em_ref, em_mine)
    agreed += 1


counts = [10, 30, 60]

# aggregate_utils
for weighted in (True, False):
    results = [(weights(), n) for n in counts]
    same_arrays(ref_utils.aggregate_results(results, weighted), my_utils.aggregate_results(results, weighted))
    losses = [(n, float(i)) for i, n in enumerate(counts)]
    assert abs(ref_utils.aggregate_losses(losses, weighted) - my_utils.aggregate_losses(losses, weighted)) < 1e-9
    agreed += 1

# BasicFedAvg
for weighted in (True, False):
    run(ref_basic.BasicFedAvg(**common(False), weighted_aggregation=weighted, weighted_eval_losses=weighted),
        my_basic.BasicFedAvg(**common(True), weighted_aggregation=weighted, weighted_eval_losses=weighted),
        [[weights() for _ in counts] for _ in range(2)], counts)

# adaptive constraint: payload = weights ++ [train loss]; the loss weight adapts over rounds (patience 2)
def with_loss(loss_per_client):
    return [ref_pack.ParameterPackerAdaptiveConstraint().pack_parameters(weights(), loss) for loss in loss_per_client]

trajectory = [[1.0, 1.1, 0.9], [1.2, 1.3, 1.1], [1.4, 1.5, 1.3], [0.5, 0.6, 0.4], [0.4, 0.5, 0.3], [0.3, 0.4, 0.2], [0.2, 0.3, 0.1]]
for weighted in (True, False):
    kwargs = dict(initial_loss_weight=0.3, adapt_loss_weight=True, loss_weight_delta=0.1, loss_weight_patience=2,
                  weighted_aggregation=weighted, weighted_train_losses=weighted)
    initial = weights()
    run(ref_adaptive.FedAvgWithAdaptiveConstraint(**common(False), initial_parameters=fc.ndarrays_to_parameters(initial), **kwargs),
        my_adaptive.FedAvgWithAdaptiveConstraint(**common(True), initial_parameters=mt.ndarrays_to_parameters(initial), **kwargs),
        [with_loss(losses) for losses in trajectory], counts)

# dynamic layers: every client sends a different subset of named layers
names = ["a.weight", "a.bias", "b.weight", "b.bias"]
def named_subset(keep):
    full = weights()
    return ref_pack.ParameterPackerWithLayerNames().pack_parameters([full[i] for i in keep], [names[i] for i in keep])
for weighted in (True, False):
    run(ref_dynamic.FedAvgDynamicLayer(**common(False), weighted_aggregation=weighted),
        my_dynamic.FedAvgDynamicLayer(**common(True), weighted_aggregation=weighted),
        [[named_subset([0, 1, 2]), named_subset([1, 2, 3]), named_subset([0, 3])], [named_subset([2]), named_subset([2, 3]), named_subset([0, 1, 2, 3])]], counts)

# sparse COO tensors: every client sends its own sparse selection of every tensor
def sparse_payload():
    values, coordinates, shapes, kept = [], [], [], []
    for name, shape in zip(names, SHAPES):
        keep = rng.random(shape) > 0.5
        keep.flat[0] = True  # the reference cannot sort a client whose selection of a tensor is empty
        dense = torch.from_numpy(rng.normal(size=shape).astype(np.float32)) * torch.from_numpy(keep)
        v, c, s = ref_pack.SparseCooParameterPacker.extract_coo_info_from_dense(dense)
        values.append(v); coordinates.append(c); shapes.append(s); kept.append(name)
    return ref_pack.SparseCooParameterPacker().pack_parameters(values, (coordinates, shapes, kept))
for weighted in (True, False):
    run(ref_sparse.FedAvgSparseCooTensor(**common(False), weighted_aggregation=weighted),
        my_sparse.FedAvgSparseCooTensor(**common(True), weighted_aggregation=weighted),
        [[sparse_payload() for _ in counts] for _ in range(2)], counts)

# Flash: server-side adaptive optimizer state over four rounds
initial = weights()
for weighted in (True, False):
    kwargs = dict(eta=0.1, eta_l=0.05, beta_1=0.9, beta_2=0.99, tau=1e-3, weighted_aggregation=weighted)
    run(ref_flash.Flash(**common(False), initial_parameters=fc.ndarrays_to_parameters(initial), **kwargs),
        my_flash.Flash(**common(True), initial_parameters=mt.ndarrays_to_parameters(initial), **kwargs),
        [[weights() for _ in counts] for _ in range(4)], counts, tol=1e-4)

# FedPM: binary masks with layer names; Bayesian aggregation keeps Beta posteriors across rounds
def masks():
    return ref_pack.ParameterPackerWithLayerNames().pack_parameters([(rng.random(s) > 0.5).astype(np.float32) for s in SHAPES], list(names))
for bayesian in (True, False):
    run(ref_pm.FedPm(**common(False), bayesian_aggregation=bayesian), my_pm.FedPm(**common(True), bayesian_aggregation=bayesian),
        [[masks() for _ in counts] for _ in range(3)], counts)

# SCAFFOLD: weights ++ control-variate updates, server learning rate
initial, variates = weights(), [np.zeros(s, dtype=np.float32) for s in SHAPES]
for lr in (1.0, 0.5):
    def scaffold_payload():
        return ref_pack.ParameterPackerWithControlVariates(len(SHAPES)).pack_parameters(weights(), weights())
    base = {k: v for k, v in common(False).items() if k not in ("min_fit_clients", "min_evaluate_clients")}
    mine = {k: v for k, v in common(True).items() if k not in ("min_fit_clients", "min_evaluate_clients")}
    packed_initial = ref_pack.ParameterPackerWithControlVariates(len(SHAPES)).pack_parameters(initial, variates)
    run(ref_scaffold.Scaffold(**base, initial_parameters=fc.ndarrays_to_parameters(initial), initial_control_variates=fc.ndarrays_to_parameters(variates), learning_rate=lr),
        my_scaffold.Scaffold(**mine, initial_parameters=mt.ndarrays_to_parameters(initial), initial_control_variates=mt.ndarrays_to_parameters(variates), learning_rate=lr),
        [[scaffold_payload() for _ in counts] for _ in range(3)], counts)

# FedPCA: principal components + singular values per client, merged by SVD or QR
def subspace():
    q, _ = np.linalg.qr(rng.normal(size=(12, 5)))
    return [q.astype(np.float32), np.sort(rng.random(5).astype(np.float32))[::-1].copy()]  # components are columns
for merging in (True, False):
    theirs, ours = ref_pca.FedPCA(**common(False), svd_merging=merging), my_pca.FedPCA(**common(True), svd_merging=merging)
    payloads = [subspace() for _ in counts]
    res_ref, res_mine = fit_results(payloads, counts)
    (p_ref, _), (p_mine, _) = theirs.aggregate_fit(1, res_ref, []), ours.aggregate_fit(1, res_mine, [])
    (v_ref, s_ref), (v_mine, s_mine) = arrays(p_ref, False), arrays(p_mine, True)
    assert np.allclose(s_ref, s_mine, atol=1e-4), (s_ref, s_mine)
    # components are defined up to sign: compare the projectors of the leading directions
    assert v_ref.shape == v_mine.shape
    assert np.allclose(np.abs(v_ref.T @ v_mine).diagonal()[:3], 1.0, atol=1e-3), np.abs(v_ref.T @ v_mine).diagonal()
    agreed += 1

# model merge: one-shot average of whole models
for weighted in (True, False):
    theirs = ref_merge.ModelMergeStrategy(**common(False), weighted_aggregation=weighted)
    ours = my_merge.ModelMergeStrategy(**common(True), weighted_aggregation=weighted)
    res_ref, res_mine = fit_results([weights() for _ in counts], counts)
    (p_ref, m_ref), (p_mine, m_mine) = theirs.aggregate_fit(1, res_ref, []), ours.aggregate_fit(1, res_mine, [])
    same_arrays(arrays(p_ref, False), arrays(p_mine, True)); same_metrics(m_ref, m_mine)
    agreed += 1
print("configs agree:", agreed)
