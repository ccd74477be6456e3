"""More clients than GPUs: K_r clients hosted by each rank of an SPMD job.

``parallel/spmd.py`` maps exactly one client to each rank — the layout of the headline benchmark, where the fused
peer-memory collectives read the client arenas in place.  Federations are usually larger than the machine (the
reference runs one OS process per client regardless of hardware), so this module adds the general case on top of the
same context and payload types:

* every rank builds the (replicated, deterministic) server and its OWN list of clients; client ids are
  ``rank{r}.{i}`` and sort by (rank, index), so client sampling picks the same set on every rank;
* ``fit`` / ``evaluate`` of the selected clients run sequentially on the rank that hosts them (one CUDA stream per
  rank — per-client CUDA graphs, arenas and loaders work unchanged);
* result metadata travels as one object all-gather per phase; payloads stay where they were produced: a weighted
  aggregate first combines a rank's local payloads with one streaming kernel and then reduces the per-rank partials with
  ONE all-reduce (``strategies/aggregate_utils._spmd_weighted_combine_multi``), so communication does not grow with the
  number of clients; strategies that need whole payloads fall back to per-payload broadcasts (``spmd.materialize``).

Nothing here is used by the 1:1 path.
"""

from __future__ import annotations

from logging import INFO, WARNING
from typing import Any

import numpy as np

from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import (
    Code,
    EvaluateIns,
    EvaluateRes,
    FitIns,
    FitRes,
    GetParametersIns,
    GetParametersRes,
    GetPropertiesIns,
    GetPropertiesRes,
    NDArrays,
    Status,
    ndarrays_to_parameters,
    parameters_to_ndarrays,
)
from fl4health_b200.parallel.spmd import (
    PayloadSpec,
    RemoteNDArrays,
    SpmdClientProxy,
    SpmdContext,
    _TaggedParameters,
    _tag_local,
)
from fl4health_b200.servers.client_proxy import ClientProxy


class SpmdHostedClientProxy(SpmdClientProxy):
    """Proxy of client ``index`` hosted by ``rank`` (``client`` is None on every other rank)."""

    def __init__(self, ctx: SpmdContext, rank: int, index: int, client: Any | None) -> None:
        super().__init__(ctx, rank, client)
        self.index = index
        self.cid = f"rank{rank:03d}.{index:03d}"
        if self.local is not None:
            self.local.cid = self.cid


class SpmdMultiClientTransport:
    """``fit_clients`` / ``evaluate_clients`` / ``poll_clients`` when ranks host several clients each."""

    def __init__(self, ctx: SpmdContext) -> None:
        self.ctx = ctx

    def is_coordinator(self) -> bool:
        return self.ctx.rank == 0

    def _run_local(self, pairs: list[tuple[ClientProxy, Any]], method: str, timeout: float | None, group_id: int | None) -> dict[str, tuple[Any, str | None]]:
        """{cid: (result, error)} for every selected client hosted by this rank, in selection order."""
        out: dict[str, tuple[Any, str | None]] = {}
        for proxy, ins in pairs:
            if isinstance(proxy, SpmdClientProxy) and proxy.rank == self.ctx.rank:
                try:
                    out[proxy.cid] = (getattr(proxy, method)(ins, timeout=timeout, group_id=group_id), None)
                except Exception as exc:  # noqa: BLE001
                    import traceback

                    log(WARNING, f"client {proxy.cid} failed in {method}: {exc!r}\n{traceback.format_exc()}")
                    out[proxy.cid] = (None, repr(exc))
        return out

    def _gather(self, local_meta: dict[str, Any]) -> dict[str, Any]:
        merged: dict[str, Any] = {}
        for part in self.ctx.all_gather_object(local_meta):
            merged.update(part)
        return merged

    def fit_clients(self, client_instructions: list[tuple[ClientProxy, FitIns]], max_workers: int | None,
                    timeout: float | None, group_id: int | None = None) -> tuple[list, list]:
        local = self._run_local(client_instructions, "fit", timeout, group_id)
        payloads: dict[str, NDArrays] = {}
        meta: dict[str, Any] = {}
        for cid, (res, err) in local.items():
            if res is None:
                meta[cid] = {"error": err}
                continue
            payloads[cid] = _tag_local(parameters_to_ndarrays(res.parameters), self.ctx)
            meta[cid] = {"n": res.num_examples, "metrics": res.metrics, "spec": payloads[cid].spec, "code": res.status.code}  # type: ignore[attr-defined]
        all_meta = self._gather(meta)
        results: list = []
        failures: list = []
        for proxy, _ in client_instructions:
            assert isinstance(proxy, SpmdClientProxy)
            m = all_meta.get(proxy.cid)
            if m is None or "error" in m:
                failures.append(RuntimeError(f"client {proxy.cid} failed: {m['error'] if m else 'no result'}"))
                continue
            arrays: NDArrays = payloads[proxy.cid] if proxy.rank == self.ctx.rank else RemoteNDArrays(self.ctx, proxy.rank, m["spec"])
            fit_res = FitRes(Status(m["code"]), _TaggedParameters(arrays), m["n"], m["metrics"])
            (results if m["code"] == Code.OK else failures).append((proxy, fit_res))
        return results, failures

    def evaluate_clients(self, client_instructions: list[tuple[ClientProxy, EvaluateIns]], max_workers: int | None,
                         timeout: float | None, group_id: int | None = None) -> tuple[list, list]:
        local = self._run_local(client_instructions, "evaluate", timeout, group_id)
        meta = {cid: ({"loss": res.loss, "n": res.num_examples, "metrics": res.metrics, "code": res.status.code} if res is not None
                      else {"error": err}) for cid, (res, err) in local.items()}
        all_meta = self._gather(meta)
        results: list = []
        failures: list = []
        for proxy, _ in client_instructions:
            m = all_meta.get(proxy.cid)
            if m is None or "error" in m:
                failures.append(RuntimeError(f"client {proxy.cid} failed: {m['error'] if m else 'no result'}"))
                continue
            eval_res = EvaluateRes(Status(m["code"]), m["loss"], m["n"], m["metrics"])
            (results if m["code"] == Code.OK else failures).append((proxy, eval_res))
        return results, failures

    def poll_clients(self, client_instructions: list[tuple[ClientProxy, GetPropertiesIns]], max_workers: int | None,
                     timeout: float | None) -> tuple[list, list]:
        local = self._run_local(client_instructions, "get_properties", timeout, None)
        meta = {cid: ({"properties": res.properties, "code": res.status.code} if res is not None else {"error": err})
                for cid, (res, err) in local.items()}
        all_meta = self._gather(meta)
        results: list = []
        failures: list = []
        for proxy, _ in client_instructions:
            m = all_meta.get(proxy.cid)
            if m is None or "error" in m:
                failures.append(RuntimeError(f"client {proxy.cid} failed to report properties"))
            else:
                results.append((proxy, GetPropertiesRes(Status(m["code"]), m["properties"])))
        return results, failures

    def get_parameters(self, proxy: ClientProxy, ins: GetParametersIns, timeout: float | None, server_round: int) -> GetParametersRes:
        """Initial-parameter request: the hosting rank's client answers, everyone receives a copy."""
        assert isinstance(proxy, SpmdClientProxy)
        arrays: NDArrays = NDArrays()
        spec = None
        if proxy.rank == self.ctx.rank:
            res = proxy.get_parameters(ins, timeout, server_round)
            arrays = NDArrays(res.parameters.tensors)
            spec = PayloadSpec.of(arrays)
        spec = self.ctx.broadcast_object(spec, src=proxy.rank)
        full = self.ctx.materialize(arrays, proxy.rank, spec)
        return GetParametersRes(Status(Code.OK), ndarrays_to_parameters(full))


def build_spmd_federation_multi(ctx: SpmdContext, server: Any, local_clients: list[Any]) -> list[SpmdHostedClientProxy]:
    """Register ``sum_r K_r`` proxies with the server's client manager: this rank's ``local_clients`` plus placeholders
    for the clients hosted elsewhere.  Ranks may host different numbers of clients."""
    import random

    from fl4health_b200.parallel.spmd import decorrelate_client_randomness
    from fl4health_b200.servers.client_manager import sampling_streams

    counts = ctx.all_gather_object(len(local_clients))
    seed = ctx.broadcast_object(sampling_streams.base_seed, src=0)  # replicated server logic: identical client sampling
    sampling_streams.seed(seed)  # server-side streams: rank-local use of the global RNGs cannot desynchronise them
    decorrelate_client_randomness(ctx.rank)
    proxies = []
    for rank, count in enumerate(counts):
        for index in range(count):
            client = local_clients[index] if rank == ctx.rank else None
            proxy = SpmdHostedClientProxy(ctx, rank, index, client)
            server.client_manager().register(proxy)
            proxies.append(proxy)
    server.transport = SpmdMultiClientTransport(ctx)
    log(INFO, f"SPMD federation: {sum(counts)} clients over {ctx.world_size} ranks ({counts})")
    return proxies
