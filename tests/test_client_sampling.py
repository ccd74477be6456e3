"""Client managers: sampling semantics (reference ``fl4health/client_managers/*``) and the dedicated sampling streams
that keep replicated servers drawing the same cohort whatever the rest of the process does with the global RNGs."""

import random

import numpy as np

from fl4health_b200.client_managers.fixed_without_replacement_manager import FixedSamplingByFractionClientManager
from fl4health_b200.client_managers.poisson_sampling_manager import PoissonSamplingClientManager
from fl4health_b200.servers.client_manager import SimpleClientManager, sampling_streams
from fl4health_b200.servers.client_proxy import ClientProxy
from fl4health_b200.utils.random import set_all_random_seeds, unset_all_random_seeds


class _Proxy(ClientProxy):
    def get_properties(self, ins, timeout, server_round):  # noqa: ANN001, ANN201
        raise NotImplementedError

    get_parameters = fit = evaluate = reconnect = get_properties


def _registered(manager, n: int = 20):  # noqa: ANN001, ANN202
    for i in range(n):
        manager.register(_Proxy(cid=f"c{i:02d}"))
    return manager


def _cohorts(seed: int, pollute) -> list[list[str]]:  # noqa: ANN001
    """Three rounds of sampling by each manager type, with `pollute()` consuming the global generators in between --
    what one rank's data loading does and another's does not."""
    sampling_streams.seed(seed)
    drawn = []
    for manager, draw in ((_registered(SimpleClientManager()), lambda m: m.sample(5)),
                          (_registered(PoissonSamplingClientManager()), lambda m: m.sample_fraction(0.4)),
                          (_registered(FixedSamplingByFractionClientManager()), lambda m: m.sample_fraction(0.3))):
        for _ in range(3):
            pollute()
            drawn.append([proxy.cid for proxy in draw(manager)])
    return drawn


def test_cohorts_do_not_depend_on_the_global_generators() -> None:
    quiet = _cohorts(123, lambda: None)
    noisy = _cohorts(123, lambda: (random.random(), np.random.rand(7), random.shuffle(list(range(9)))))
    assert quiet == noisy
    assert _cohorts(124, lambda: None) != quiet
    assert all(len(cohort) == 5 for cohort in quiet[:3]) and all(len(cohort) == 6 for cohort in quiet[6:])


def test_global_seed_fixes_the_cohorts() -> None:
    def run() -> list[str]:
        set_all_random_seeds(77)
        return [proxy.cid for proxy in _registered(SimpleClientManager()).sample(6)]

    assert run() == run()
    unset_all_random_seeds()
    random.seed(5)  # unseeded streams start from one draw of the global generator
    first = [proxy.cid for proxy in _registered(SimpleClientManager()).sample(6)]
    sampling_streams.seed(None)
    random.seed(5)
    assert [proxy.cid for proxy in _registered(SimpleClientManager()).sample(6)] == first
    sampling_streams.seed(None)
