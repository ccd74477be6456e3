"""Data utilities vs the reference under identical seeds: samplers, the Dirichlet partitioner, the FedProx-paper synthetic
generators, result ordering helpers.  Where the two draw from the random generators in the same order the outputs are
compared exactly; otherwise the checked quantity is the one the utility promises (sizes, label proportions)."""
import random

import numpy as np
import torch

import fl4health.utils.data_generation as ref_gen
import fl4health.utils.dataset as ref_ds
import fl4health.utils.functions as ref_fn
import fl4health.utils.partitioners as ref_part
import fl4health.utils.sampler as ref_sampler
import fl4health_b200.utils.data_generation as my_gen
import fl4health_b200.utils.dataset as my_ds
import fl4health_b200.utils.functions as my_fn
import fl4health_b200.utils.partitioners as my_part
import fl4health_b200.utils.sampler as my_sampler

agreed = 0


def seed(value: int) -> None:
    random.seed(value); np.random.seed(value); torch.manual_seed(value)


def labelled(module, n: int = 600, classes: int = 5):
    generator = torch.Generator().manual_seed(99)
    return module.TensorDataset(torch.randn(n, 7, generator=generator), torch.randint(0, classes, (n,), generator=generator))


def label_histogram(dataset, classes: int = 5) -> np.ndarray:
    return np.bincount(np.asarray(dataset.targets), minlength=classes)


# -- samplers -------------------------------------------------------------------------------------------------------
for ratio, minority in ((0.1, {1, 3}), (0.5, {0})):
    seed(4); a = ref_sampler.MinorityLabelBasedSampler(list(range(5)), ratio, minority).subsample(labelled(ref_ds))
    seed(4); b = my_sampler.MinorityLabelBasedSampler(list(range(5)), ratio, minority).subsample(labelled(my_ds))
    assert (label_histogram(a) == label_histogram(b)).all(), (label_histogram(a), label_histogram(b))
    assert torch.equal(torch.as_tensor(a.targets).sort().values, torch.as_tensor(b.targets).sort().values)
    agreed += 1
for beta, fraction, hash_key in ((0.5, 0.5, 7), (100.0, 0.25, 11), (1.0, 0.75, 3)):
    # with a hash key the class proportions are a pure function of the key: the subsample sizes per label must agree
    seed(5); a = ref_sampler.DirichletLabelBasedSampler(list(range(5)), hash_key=hash_key, sample_percentage=fraction, beta=beta).subsample(labelled(ref_ds))
    seed(5); b = my_sampler.DirichletLabelBasedSampler(list(range(5)), hash_key=hash_key, sample_percentage=fraction, beta=beta).subsample(labelled(my_ds))
    assert len(a) == len(b) and (label_histogram(a) == label_histogram(b)).all(), (label_histogram(a), label_histogram(b))
    agreed += 1

# -- Dirichlet partitioner --------------------------------------------------------------------------------------------
for partitions, beta, minimum in ((4, 0.5, None), (3, 5.0, 2)):
    seed(6); a, pa = ref_part.DirichletLabelBasedAllocation(partitions, list(range(5)), min_label_examples=minimum, beta=beta).partition_dataset(labelled(ref_ds), max_retries=5)
    seed(6); b, pb = my_part.DirichletLabelBasedAllocation(partitions, list(range(5)), min_label_examples=minimum, beta=beta).partition_dataset(labelled(my_ds), max_retries=5)
    assert [len(x) for x in a] == [len(x) for x in b], ([len(x) for x in a], [len(x) for x in b])
    for x, y in zip(a, b):
        assert (label_histogram(x) == label_histogram(y)).all()
    assert pa.keys() == pb.keys() and all(np.allclose(pa[k], pb[k]) for k in pa)
    agreed += 1
prior = {label: np.full(3, 1 / 3) for label in range(5)}
seed(8); a, _ = ref_part.DirichletLabelBasedAllocation(3, list(range(5)), prior_distribution=prior).partition_dataset(labelled(ref_ds))
seed(8); b, _ = my_part.DirichletLabelBasedAllocation(3, list(range(5)), prior_distribution=prior).partition_dataset(labelled(my_ds))
assert [len(x) for x in a] == [len(x) for x in b]
agreed += 1

# -- FedProx synthetic data -------------------------------------------------------------------------------------------
for kwargs in (dict(num_clients=3, alpha=0.5, beta=0.5, samples_per_client=50), dict(num_clients=2, alpha=0.0, beta=0.0, temperature=2.0, input_dim=20, output_dim=4, samples_per_client=40),
               dict(num_clients=2, alpha=1.0, beta=1.0, hidden_dim=12, samples_per_client=30)):
    seed(9); a = ref_gen.SyntheticNonIidFedProxDataset(**kwargs).generate()
    seed(9); b = my_gen.SyntheticNonIidFedProxDataset(**kwargs).generate()
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert torch.allclose(torch.as_tensor(x.data), torch.as_tensor(y.data), atol=1e-5) and torch.equal(torch.as_tensor(x.targets), torch.as_tensor(y.targets))
    agreed += 1
seed(10); a = ref_gen.SyntheticIidFedProxDataset(num_clients=3, samples_per_client=40).generate()
seed(10); b = my_gen.SyntheticIidFedProxDataset(num_clients=3, samples_per_client=40).generate()
for x, y in zip(a, b):
    assert torch.allclose(torch.as_tensor(x.data), torch.as_tensor(y.data), atol=1e-5) and torch.equal(torch.as_tensor(x.targets), torch.as_tensor(y.targets))
agreed += 1

# -- ordering helpers and small functions ---------------------------------------------------------------------------
x = torch.rand(20) * 0.98 + 0.01
assert torch.allclose(ref_fn.sigmoid_inverse(x), my_fn.sigmoid_inverse(x)); agreed += 1
for array in (np.arange(24.0).reshape(2, 3, 4) + 5, np.array([3.5]), np.array(2.5)):
    assert ref_fn.select_zeroeth_element(array) == my_fn.select_zeroeth_element(array)
agreed += 1
probs = torch.rand(1000)
seed(12); a = ref_fn.BernoulliSample.apply(probs)
seed(12); b = my_fn.BernoulliSample.apply(probs)
assert torch.equal(a, b); agreed += 1
p = probs.clone().requires_grad_(True); ref_fn.BernoulliSample.apply(p).sum().backward()
q = probs.clone().requires_grad_(True); my_fn.BernoulliSample.apply(q).sum().backward()
assert torch.equal(p.grad, q.grad); agreed += 1  # straight-through gradient

import flwr.common as fc
import fl4health_b200.common.typing as mt


class Proxy:
    def __init__(self, cid: str) -> None:
        self.cid = cid


rng = np.random.default_rng(1)
payloads = [[rng.normal(size=(3, 2)).astype(np.float32), rng.integers(0, 5, size=(2,))] for _ in range(5)]
counts = [30, 10, 30, 20, 10]
ref_results = [(Proxy(str(i)), fc.FitRes(fc.Status(fc.Code.OK, ""), fc.ndarrays_to_parameters(p), n, {})) for i, (p, n) in enumerate(zip(payloads, counts))]
my_results = [(Proxy(str(i)), mt.FitRes(mt.Status(mt.Code.OK, ""), mt.ndarrays_to_parameters(p), n, {})) for i, (p, n) in enumerate(zip(payloads, counts))]
decoded_ref, decoded_mine = ref_fn.decode_and_pseudo_sort_results(ref_results), my_fn.decode_and_pseudo_sort_results(my_results)
by_cid_ref = {proxy.cid: ref_fn.pseudo_sort_scoring_function((proxy, arrays, n)) for proxy, arrays, n in decoded_ref}
by_cid_mine = {proxy.cid: my_fn.pseudo_sort_scoring_function((proxy, arrays, n)) for proxy, arrays, n in decoded_mine}
assert by_cid_ref.keys() == by_cid_mine.keys() and all(abs(by_cid_ref[c] - by_cid_mine[c]) < 1e-6 for c in by_cid_ref); agreed += 1
# deliberate difference: the reference orders results by that score (its client ids are random UUIDs); ours orders by the
# stable client id, which fixes the summation order without reading the payload
assert [proxy.cid for proxy, _, _ in decoded_mine] == sorted(by_cid_mine); agreed += 1
print("configs agree:", agreed)
