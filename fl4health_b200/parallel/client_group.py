"""A client that spans several GPUs (intra-client data parallelism).

The reference only ever reaches this through one example — ``examples/fedllm_example`` launches each *client* with
``torchrun --nproc_per_node=2`` and lets DeepSpeed shard it (``client.py:104-106``, ``run_client_zero_3.slrm:72``) while
the library itself is unaware.  Here the notion is part of the SPMD runtime:

* the world of ``W`` ranks is cut into consecutive groups of ``G`` ranks; every group is ONE federated client
  (``ClientGroup``), every rank of the group holds a full replica of the client's model and a disjoint shard of its
  data (``shard_dataset``);
* ``ReplicatedClientMixin`` averages the gradients over the group in a pre-hook of every ``optimizer.step()`` (one
  collective on the flat arena gradient — the arena makes the "bucket" the whole model), so all replicas take the
  same step on the union batch, whichever ``train_step`` the client algorithm brings;
* the federation layer needs no change: each replica reports its shard's sample count, and because the replicas of a
  client are identical, the sample-weighted aggregate over all ``W`` ranks equals the client-weighted aggregate over
  the ``W / G`` clients (``Σ_r n_r w_r = Σ_k (Σ_{r∈k} n_r) w_k``).  Metrics aggregate the same way.

Buffers that are not parameters (batch-norm running statistics) are per-replica during a round — each replica sees
its own shard — and are merged by the round-end aggregate like any other exchanged tensor.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Any

import torch
import torch.distributed as dist

from fl4health_b200.parallel.arena import arena_of
from fl4health_b200.utils.dataset import TensorDataset


@dataclass
class ClientGroup:
    """This rank's place in the client it belongs to."""

    client_index: int
    group_rank: int
    group_size: int
    process_group: Any  # torch.distributed.ProcessGroup | None (None when group_size == 1 or the world is one rank)

    @classmethod
    def from_world(cls, rank: int, world_size: int, group_size: int) -> ClientGroup:
        """Consecutive ranks form a client (ranks ``[k*G, (k+1)*G)`` -> client ``k``): neighbours share an NVSwitch
        domain anyway, and the layout keeps ``rank // G`` as the client index.  Collective over the WORLD: every rank
        must call it (``new_group`` requires all ranks to create all groups in the same order)."""
        if world_size % group_size != 0:
            raise ValueError(f"world size {world_size} is not a multiple of the client group size {group_size}")
        mine = None
        if group_size > 1 and world_size > 1:
            for start in range(0, world_size, group_size):
                group = dist.new_group(list(range(start, start + group_size)))
                if start <= rank < start + group_size:
                    mine = group
        return cls(rank // group_size, rank % group_size, group_size, mine)

    def all_reduce_mean(self, tensor: torch.Tensor) -> torch.Tensor:
        if self.process_group is not None:
            dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=self.process_group)
            tensor.div_(self.group_size)
        return tensor


def require_replica_safe_strategy(strategy: Any, group_size: int) -> None:
    """Every replica registers as a federated client, which is exact only for aggregates that are (sample-weighted or
    uniform) MEANS of what the replicas report: identical replicas then count as one client with the summed weight.
    Strategies whose math depends on the number or identity of the participants are wrong under that bookkeeping:
    client-level DP clips every registered client separately (a real client would contribute ``G`` clipped updates,
    ``G`` times the sensitivity the accountant assumes), SCAFFOLD / FedDG-GA keep per-participant state and divide by
    the participant count, FedPM's Bayesian vote and FedPCA's subspace merge count participants.  Refuse those."""
    if group_size <= 1:
        return
    from fl4health_b200.strategies.client_dp_fedavgm import ClientLevelDPFedAvgM
    from fl4health_b200.strategies.feddg_ga import FedDgGa
    from fl4health_b200.strategies.fedpca import FedPCA
    from fl4health_b200.strategies.fedpm import FedPm
    from fl4health_b200.strategies.scaffold import Scaffold

    for family, why in ((ClientLevelDPFedAvgM, "per-client clipping and noise calibrated to the number of clients"),
                        (Scaffold, "per-client control variates averaged over the client count"),
                        (FedDgGa, "per-client generalisation-gap weights"), (FedPm, "a per-client Bayesian mask vote"),
                        (FedPCA, "a merge of per-client subspaces")):
        if isinstance(strategy, family):
            raise ValueError(f"{type(strategy).__name__} uses {why}; with {group_size} ranks per client every replica "
                             "would be counted as a client of its own. Run this strategy with one rank per client.")


def shard_dataset(dataset: TensorDataset, group: ClientGroup, seed: int = 0) -> TensorDataset:
    """Replica ``r`` of ``G`` keeps samples ``perm[r::G]`` of a seeded permutation: disjoint, equal-sized up to one."""
    assert dataset.targets is not None
    if group.group_size == 1:
        return dataset
    order = torch.randperm(len(dataset.data), generator=torch.Generator().manual_seed(seed))
    usable = len(order) - len(order) % group.group_size  # equal shards: every replica takes the same number of steps
    mine = order[:usable][group.group_rank::group.group_size]
    return TensorDataset(dataset.data[mine], dataset.targets[mine], dataset.transform, dataset.target_transform, dataset.batch_transform)


class ReplicatedClientMixin:
    """Combine as ``class C(ReplicatedClientMixin, SomeClient)`` and set ``client.client_group``.

    The group all-reduce hangs on the one point every training step reaches whatever ``train_step`` looks like:
    ``optimizer.step()``.  Each optimizer of the client gets a step pre-hook (``torch.optim.Optimizer.
    register_step_pre_hook``; the one-launch flat optimizers are ``Optimizer`` subclasses too) that averages the
    gradients of the parameters that optimizer owns over the group.  Clients with several optimizers and their own
    ``train_step`` (Ditto: global twin + personal model, APFL, FedRep's head / representation phases, ensembles) are
    covered without knowing about replication, and gradient corrections applied in ``transform_gradients`` (SCAFFOLD's
    control variates, DP clipping) still act on the local gradient first.
    """

    client_group: ClientGroup | None = None

    def setup_client(self, config: Any) -> None:
        group = self.client_group
        engine = getattr(self, "engine", None)
        if group is not None and group.process_group is not None and engine is not None and getattr(engine, "cuda_graphs", False):
            # the gradient collective sits inside train_step, i.e. inside the captured region; NCCL calls can be
            # captured, but that path has not been validated on hardware yet, so replicated clients run eagerly
            from dataclasses import replace
            from logging import WARNING

            from fl4health_b200.common.logger import log

            log(WARNING, "client spans several ranks: CUDA-graph capture of the training step is disabled for it")
            self.engine = replace(engine, cuda_graphs=False)
        super().setup_client(config)  # type: ignore[misc]
        self._hook_group_averaging()

    def train_step(self, input: Any, target: Any) -> Any:
        self._hook_group_averaging()  # optimizers re-created since set-up (new round, new phase) are picked up here
        return super().train_step(input, target)  # type: ignore[misc]

    def _hook_group_averaging(self) -> None:
        group = self.client_group
        if group is None or group.process_group is None:
            return
        for optimizer in getattr(self, "optimizers", {}).values():
            if getattr(optimizer, "_fl4h_group_hook", None) is None:
                if not hasattr(optimizer, "register_step_pre_hook"):
                    raise TypeError(f"{type(optimizer).__name__} is not a torch.optim.Optimizer: a client spanning several "
                                    "ranks needs step pre-hooks to average its gradients over the group")
                optimizer._fl4h_group_hook = optimizer.register_step_pre_hook(self._average_before_step)

    def _average_before_step(self, optimizer: Any, args: Any, kwargs: Any) -> None:
        group = self.client_group
        if group is None or group.process_group is None:
            return
        params = self._parameters_reduced_for(optimizer)
        average_parameter_gradients(params, group, self._flat_gradient_of(params))

    def _parameters_reduced_for(self, optimizer: Any) -> list[torch.nn.Parameter]:
        """Whose gradients are averaged before ``optimizer`` steps: by default exactly what it updates."""
        return [p for param_group in optimizer.param_groups for p in param_group["params"] if p.requires_grad]

    def _flat_gradient_of(self, params: list[torch.nn.Parameter]) -> torch.Tensor | None:
        """The arena's flat gradient region when ``params`` are exactly one arena-backed module's trainable parameters
        (then one collective on the flat buffer does it)."""
        wanted = {id(p) for p in params}
        modules = [getattr(self, "model", None), getattr(self, "global_model", None)]
        candidates = getattr(self, "_candidate_modules", None)
        if callable(candidates):
            modules = list(candidates())
        for module in modules:
            if isinstance(module, torch.nn.Module):
                arena = arena_of(module)
                flat = getattr(arena, "grad", None) if arena is not None else None
                if flat is not None and wanted == {id(p) for p in module.parameters() if p.requires_grad}:
                    return flat
        return None


def average_gradients(model: torch.nn.Module, group: ClientGroup) -> None:
    """Mean of a model's gradients over the group.  One collective when the gradients live in the arena's flat buffer;
    otherwise the per-parameter gradients are coalesced into one temporary per dtype."""
    arena = arena_of(model)
    average_parameter_gradients([p for p in model.parameters() if p.requires_grad], group,
                                getattr(arena, "grad", None) if arena is not None else None)


def average_parameter_gradients(params: list[torch.nn.Parameter], group: ClientGroup, flat: torch.Tensor | None = None) -> None:
    if group.process_group is None or not params:
        return
    if flat is not None and all(p.grad is not None and _is_view_of(p.grad, flat) for p in params):
        group.all_reduce_mean(flat)
        return
    for p in params:  # a parameter unused on this replica must still take part: the collective is by position
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    by_dtype: dict[torch.dtype, list[torch.Tensor]] = {}
    for p in params:
        by_dtype.setdefault(p.grad.dtype, []).append(p.grad)  # type: ignore[union-attr]
    for bucket in by_dtype.values():
        coalesced = torch.cat([g.reshape(-1) for g in bucket])
        group.all_reduce_mean(coalesced)
        offset = 0
        for g in bucket:
            g.copy_(coalesced[offset:offset + g.numel()].view_as(g))
            offset += g.numel()


def _is_view_of(tensor: torch.Tensor, base: torch.Tensor) -> bool:
    start, end = base.data_ptr(), base.data_ptr() + base.numel() * base.element_size()
    return tensor.device == base.device and start <= tensor.data_ptr() < end


# --------------------------------------------------------------------------------------------------------------------
# ZeRO-1: optimizer state sharded over the replicas of a client
# --------------------------------------------------------------------------------------------------------------------
def partition_parameters(params: list[torch.nn.Parameter], group_size: int) -> list[list[int]]:
    """Greedy balance by element count (largest first, to the least loaded replica); deterministic, so every replica
    computes the same ownership table.  Returns, per replica, the indices into ``params`` it owns."""
    owned: list[list[int]] = [[] for _ in range(group_size)]
    load = [0] * group_size
    for index in sorted(range(len(params)), key=lambda i: (-params[i].numel(), i)):
        target = min(range(group_size), key=lambda r: (load[r], r))
        owned[target].append(index)
        load[target] += params[index].numel()
    return [sorted(indices) for indices in owned]


class Zero1ClientMixin(ReplicatedClientMixin):
    """``ReplicatedClientMixin`` + optimizer-state sharding: every replica keeps momentum / Adam moments only for the
    parameters it owns (1/G of the model), steps those, and the owners then broadcast the new values inside the group.

    The user's ``get_optimizer`` is unchanged: the optimizer it returns is rebuilt over the owned subset with the same
    class and per-group hyper-parameters (and kept as a stock ``torch.optim`` optimizer, see ``get_optimizer``).  With
    fp32 masters in the arena the *masters* are what is broadcast, and the bf16 compute shadow is refreshed afterwards.
    """

    def get_optimizer(self, config: Any) -> Any:
        optimizer = super().get_optimizer(config)  # type: ignore[misc]
        group = self.client_group
        if group is None or group.process_group is None:
            return optimizer
        if isinstance(optimizer, dict):
            raise NotImplementedError("Zero1ClientMixin shards single-optimizer clients; multi-optimizer clients use ReplicatedClientMixin")
        model: torch.nn.Module = self.model  # type: ignore[attr-defined]
        trainable = [p for p in model.parameters() if p.requires_grad]
        table = partition_parameters(trainable, group.group_size)
        mine = {id(trainable[i]) for i in table[group.group_rank]}
        new_groups = []
        for param_group in optimizer.param_groups:
            kept = [p for p in param_group["params"] if id(p) in mine]
            if kept:
                new_groups.append({**{k: v for k, v in param_group.items() if k != "params"}, "params": kept})
        if not new_groups:  # more replicas than tensors: this replica owns nothing but still needs a valid optimizer
            new_groups = [{**{k: v for k, v in optimizer.param_groups[0].items() if k != "params"}, "params": [torch.nn.Parameter(torch.zeros(()))]}]
        names = {id(p): name for name, p in model.named_parameters()}
        self._zero1_owned_names = [[names[id(trainable[i])] for i in indices] for indices in table]
        self._zero1_foreign = [trainable[i] for r, indices in enumerate(table) if r != group.group_rank for i in indices]
        sharded = type(optimizer)(new_groups, **{k: v for k, v in optimizer.defaults.items() if k in type(optimizer).__init__.__code__.co_varnames})
        # The fused flat optimizers index arena-length companion buffers (momentum / moments at the parameter's arena
        # offset), which would allocate the full state on every replica; the stock optimizer over the owned subset is
        # what actually divides the state memory by G.  (A compact companion layout for the fused kernels is future work.)
        sharded.fl4h_keep_stock = True  # type: ignore[attr-defined]
        return sharded

    def _parameters_reduced_for(self, optimizer: Any) -> list[torch.nn.Parameter]:
        # the optimizer only holds this replica's share, but the collective is by position over the whole model
        return [p for p in self.model.parameters() if p.requires_grad]  # type: ignore[attr-defined]

    def update_after_step(self, step: int, current_round: int | None = None) -> None:
        group = self.client_group
        if group is not None and group.process_group is not None and hasattr(self, "_zero1_owned_names"):
            self._zero1_sync_parameters(group)
        super().update_after_step(step, current_round)  # type: ignore[misc]

    @torch.no_grad()
    def _zero1_sync_parameters(self, group: ClientGroup) -> None:
        model: torch.nn.Module = self.model  # type: ignore[attr-defined]
        arena = arena_of(model)
        params = dict(model.named_parameters())
        storage = (lambda name: arena.view(name)) if arena is not None else (lambda name: params[name].data)  # fp32 masters when there are any
        first_rank = group.client_index * group.group_size  # global rank of the group's replica 0
        for owner, names in enumerate(self._zero1_owned_names):
            if not names:
                continue
            tensors = [storage(name) for name in names]
            coalesced = torch.cat([t.reshape(-1).float() for t in tensors])
            dist.broadcast(coalesced, src=first_rank + owner, group=group.process_group)
            if owner != group.group_rank:
                offset = 0
                for t in tensors:
                    t.copy_(coalesced[offset:offset + t.numel()].view_as(t))
                    offset += t.numel()
        if arena is not None:
            arena.refresh_shadow()
        for p in self._zero1_foreign:  # nobody on this replica steps (or clears) these gradients
            if p.grad is not None:
                p.grad.zero_() if _grad_is_persistent(p, arena) else setattr(p, "grad", None)


def _grad_is_persistent(p: torch.nn.Parameter, arena: Any) -> bool:
    flat = getattr(arena, "grad", None) if arena is not None else None
    return flat is not None and p.grad is not None and _is_view_of(p.grad, flat)
