"""How the server reaches its clients for one round.

``LocalTransport`` (default) executes the sampled clients of this process — sequentially when they share a device,
which is the case for CPU simulation and single-GPU runs.  ``fl4health_b200.parallel.spmd.SpmdTransport`` implements the
same interface across ranks (one client per GPU) with collectives instead of threads.
"""

from __future__ import annotations

from typing import Any

from fl4health_b200.common.typing import Code, EvaluateIns, EvaluateRes, FitIns, FitRes, GetPropertiesIns, GetPropertiesRes
from fl4health_b200.servers.client_proxy import ClientProxy

FitResultsAndFailures = tuple[list[tuple[ClientProxy, FitRes]], list[Any]]
EvaluateResultsAndFailures = tuple[list[tuple[ClientProxy, EvaluateRes]], list[Any]]
PollResultsAndFailures = tuple[list[tuple[ClientProxy, GetPropertiesRes]], list[Any]]


class LocalTransport:
    def _run(self, pairs: list[tuple[ClientProxy, Any]], method: str, timeout: float | None, group_id: int | None) -> tuple[list, list]:
        results: list = []
        failures: list = []
        for proxy, ins in pairs:
            try:
                res = getattr(proxy, method)(ins, timeout=timeout, group_id=group_id)
            except Exception as exc:  # noqa: BLE001 - a failing client is a *failure*, the policy decides what next
                import traceback
                from logging import WARNING

                from fl4health_b200.common.logger import log

                log(WARNING, f"client {proxy.cid} raised in {method}: {exc!r}\n{traceback.format_exc()}")
                failures.append(exc)
                continue
            if res.status.code == Code.OK:
                results.append((proxy, res))
            else:
                failures.append((proxy, res))
        return results, failures

    def fit_clients(
        self, client_instructions: list[tuple[ClientProxy, FitIns]], max_workers: int | None, timeout: float | None,
        group_id: int | None = None,
    ) -> FitResultsAndFailures:
        return self._run(client_instructions, "fit", timeout, group_id)

    def evaluate_clients(
        self, client_instructions: list[tuple[ClientProxy, EvaluateIns]], max_workers: int | None, timeout: float | None,
        group_id: int | None = None,
    ) -> EvaluateResultsAndFailures:
        return self._run(client_instructions, "evaluate", timeout, group_id)

    def poll_clients(
        self, client_instructions: list[tuple[ClientProxy, GetPropertiesIns]], max_workers: int | None, timeout: float | None
    ) -> PollResultsAndFailures:
        return self._run(client_instructions, "get_properties", timeout, None)

    def is_coordinator(self) -> bool:
        """True on the process that owns server-side artifacts (checkpoints, reports)."""
        return True


_DEFAULT = LocalTransport()


def fit_clients(client_instructions, max_workers=None, timeout=None, group_id=None):  # noqa: ANN001, ANN201
    return _DEFAULT.fit_clients(client_instructions, max_workers, timeout, group_id)


def evaluate_clients(client_instructions, max_workers=None, timeout=None, group_id=None):  # noqa: ANN001, ANN201
    return _DEFAULT.evaluate_clients(client_instructions, max_workers, timeout, group_id)
