"""In-house DP-SGD engine (replaces the reference's dependency on Opacus; SURVEY §2.8 / hot-op L11).

Pieces, with the names users of the reference already know:

* ``GradSampleModule``  — wraps a model and computes *per-sample gradients* during the ordinary ``loss.backward()``
  with forward/backward hooks (Linear, Conv1d/2d, LayerNorm, GroupNorm, Embedding; anything else must be replaced by
  ``ModuleValidator.fix``, e.g. BatchNorm -> GroupNorm).  State-dict keys carry the ``_module.`` prefix, like Opacus.
* ``DPOptimizer``       — flat clipping: per-sample global L2 norm over all parameters, clip to ``max_grad_norm``, sum,
  add ``N(0, (noise_multiplier * max_grad_norm)^2)``, divide by the expected batch size, then the wrapped optimizer
  steps.  When the parameters' gradients are views of a flat arena the noise is ONE counter-RNG kernel over the flat
  gradient (``ops.flat.add_gaussian_``) and the wrapped optimizer can be the fused flat SGD/AdamW.
* ``DPDataLoader``      — Poisson sampling: each example joins a batch independently with probability
  ``batch_size / N`` (batches may be empty; the client skips those, as the reference does).
* ``PrivacyEngine.make_private`` — glue with the Opacus call signature used by ``InstanceLevelDpClient``.
"""

from __future__ import annotations

from collections.abc import Iterator
from logging import WARNING
from typing import Any

import torch
import torch.nn.functional as F
from torch import nn
from torch.optim import Optimizer

from fl4health_b200.common.logger import log
from fl4health_b200.ops import flat as flat_ops
from fl4health_b200.parallel.arena import arena_of

# ---------------------------------------------------------------------------------------------------------------
# per-sample gradient rules:  (module, activations, backprops) -> {parameter: grad_sample [B, *param.shape]}
# ---------------------------------------------------------------------------------------------------------------


def _linear_rule(layer: nn.Linear, a: torch.Tensor, g: torch.Tensor) -> dict[nn.Parameter, torch.Tensor]:
    out = {}
    if layer.weight.requires_grad:
        out[layer.weight] = torch.einsum("n...i,n...j->nij", g, a)
    if layer.bias is not None and layer.bias.requires_grad:
        out[layer.bias] = g.reshape(g.shape[0], -1, g.shape[-1]).sum(dim=1)
    return out


def _conv_rule(layer: nn.Module, a: torch.Tensor, g: torch.Tensor) -> dict[nn.Parameter, torch.Tensor]:
    one_d = isinstance(layer, nn.Conv1d)
    if one_d:  # treat as a 2-D convolution with height 1
        a, g = a.unsqueeze(-2), g.unsqueeze(-2)
        kernel, stride = (1, layer.kernel_size[0]), (1, layer.stride[0])
        padding = (0, layer.padding[0]) if not isinstance(layer.padding, str) else layer.padding
        dilation = (1, layer.dilation[0])
    else:
        kernel, stride, padding, dilation = layer.kernel_size, layer.stride, layer.padding, layer.dilation
    if isinstance(padding, str):
        raise NotImplementedError("string padding modes are not supported for per-sample gradients")
    n = a.shape[0]
    out = {}
    if layer.weight.requires_grad:
        cols = F.unfold(a, kernel, dilation=dilation, padding=padding, stride=stride)  # [B, Cin*k, L]
        gmat = g.reshape(n, layer.out_channels, -1)  # [B, Cout, L]
        groups = layer.groups
        cols = cols.reshape(n, groups, -1, cols.shape[-1])
        gmat = gmat.reshape(n, groups, layer.out_channels // groups, -1)
        grad = torch.einsum("ngol,ngil->ngoi", gmat, cols)
        out[layer.weight] = grad.reshape(n, *layer.weight.shape)
    if layer.bias is not None and layer.bias.requires_grad:
        out[layer.bias] = g.reshape(n, layer.out_channels, -1).sum(dim=2)
    return out


def _norm_rule(layer: nn.Module, a: torch.Tensor, g: torch.Tensor) -> dict[nn.Parameter, torch.Tensor]:
    out = {}
    if isinstance(layer, nn.LayerNorm):
        if layer.weight is not None and layer.weight.requires_grad:
            normed = F.layer_norm(a, layer.normalized_shape, eps=layer.eps)
            extra = tuple(range(1, a.dim() - len(layer.normalized_shape)))
            prod = normed * g
            out[layer.weight] = prod.sum(dim=extra) if extra else prod
        if layer.bias is not None and layer.bias.requires_grad:
            extra = tuple(range(1, a.dim() - len(layer.normalized_shape)))
            out[layer.bias] = g.sum(dim=extra) if extra else g
        return out
    assert isinstance(layer, nn.GroupNorm)
    if layer.weight is not None and layer.weight.requires_grad:
        normed = F.group_norm(a, layer.num_groups, eps=layer.eps)
        out[layer.weight] = torch.einsum("nc...,nc...->nc", normed, g)
    if layer.bias is not None and layer.bias.requires_grad:
        out[layer.bias] = g.reshape(g.shape[0], g.shape[1], -1).sum(dim=2)
    return out


def _embedding_rule(layer: nn.Embedding, a: torch.Tensor, g: torch.Tensor) -> dict[nn.Parameter, torch.Tensor]:
    if not layer.weight.requires_grad:
        return {}
    n = a.shape[0]
    grad = torch.zeros(n, *layer.weight.shape, device=g.device, dtype=g.dtype)
    index = a.reshape(n, -1, 1).expand(-1, -1, layer.embedding_dim)
    grad.scatter_add_(1, index, g.reshape(n, -1, layer.embedding_dim))
    return {layer.weight: grad}


_RULES: list[tuple[tuple[type, ...], Any]] = [
    ((nn.Linear,), _linear_rule), ((nn.Conv1d, nn.Conv2d), _conv_rule), ((nn.LayerNorm, nn.GroupNorm), _norm_rule),
    ((nn.Embedding,), _embedding_rule),
]


def _rule_for(module: nn.Module) -> Any:
    for types, rule in _RULES:
        if isinstance(module, types):
            return rule
    return None


# ---------------------------------------------------------------------------------------------------------------
# module validation / fixing
# ---------------------------------------------------------------------------------------------------------------
class ModuleValidator:
    """Finds layers whose per-sample gradients are undefined or unsupported (BatchNorm mixes samples) and replaces
    them with DP-compatible equivalents (BatchNorm -> GroupNorm with min(32, C) groups, as Opacus does)."""

    @staticmethod
    def validate(model: nn.Module, strict: bool = False) -> list[str]:
        errors = []
        for name, module in model.named_modules():
            if isinstance(module, nn.modules.batchnorm._BatchNorm):
                errors.append(f"{name}: BatchNorm cannot support training with differential privacy (mixes samples)")
            elif any(p.requires_grad for p in module.parameters(recurse=False)) and _rule_for(module) is None:
                errors.append(f"{name}: no per-sample gradient rule for {type(module).__name__}")
        if strict and errors:
            raise ValueError("\n".join(errors))
        return errors

    @staticmethod
    def is_valid(model: nn.Module) -> bool:
        return not ModuleValidator.validate(model)

    @staticmethod
    def fix(model: nn.Module) -> nn.Module:
        for name, child in list(model.named_children()):
            if isinstance(child, nn.modules.batchnorm._BatchNorm):
                groups = min(32, child.num_features)
                while child.num_features % groups != 0:
                    groups -= 1
                setattr(model, name, nn.GroupNorm(groups, child.num_features, affine=child.affine))
            else:
                ModuleValidator.fix(child)
        return model


# ---------------------------------------------------------------------------------------------------------------
# GradSampleModule
# ---------------------------------------------------------------------------------------------------------------
class GradSampleModule(nn.Module):
    def __init__(self, module: nn.Module, batch_first: bool = True, loss_reduction: str = "mean") -> None:
        super().__init__()
        errors = [e for e in ModuleValidator.validate(module) if "BatchNorm" in e or "no per-sample" in e]
        if errors:
            raise ValueError("Model is not DP-compatible:\n" + "\n".join(errors))
        self._module = module
        self.loss_reduction = loss_reduction
        self.hooks_enabled = True
        self._handles: list[Any] = []
        self._activations: dict[nn.Module, list[torch.Tensor]] = {}
        self._add_hooks()

    def forward(self, *args: Any, **kwargs: Any) -> Any:
        return self._module(*args, **kwargs)

    def _add_hooks(self) -> None:
        for module in self._module.modules():
            rule = _rule_for(module)
            if rule is None or not any(p.requires_grad for p in module.parameters(recurse=False)):
                continue
            self._handles.append(module.register_forward_hook(self._capture_activation))
            self._handles.append(module.register_full_backward_hook(self._capture_backprop))

    def remove_hooks(self) -> None:
        for handle in self._handles:
            handle.remove()
        self._handles = []

    def _capture_activation(self, module: nn.Module, inputs: tuple, output: Any) -> None:
        if not self.hooks_enabled or not module.training or not torch.is_grad_enabled():
            return
        self._activations.setdefault(module, []).append(inputs[0].detach())

    def _capture_backprop(self, module: nn.Module, grad_input: tuple, grad_output: tuple) -> None:
        if not self.hooks_enabled or module not in self._activations or not self._activations[module]:
            return
        activation = self._activations[module].pop()
        backprop = grad_output[0].detach()
        if self.loss_reduction == "mean":
            backprop = backprop * backprop.shape[0]  # undo the 1/B of the mean loss: per-sample grads are unscaled
        for param, grad_sample in _rule_for(module)(module, activation, backprop).items():
            existing = getattr(param, "grad_sample", None)
            param.grad_sample = grad_sample if existing is None else existing + grad_sample  # type: ignore[attr-defined]

    def zero_grad(self, set_to_none: bool = True) -> None:
        for param in self._module.parameters():
            param.grad_sample = None  # type: ignore[attr-defined]
        self._activations = {}
        super().zero_grad(set_to_none)

    def to_standard_module(self) -> nn.Module:
        self.remove_hooks()
        return self._module


def wrap_model(model: nn.Module, grad_sample_mode: str = "hooks", *args: Any, **kwargs: Any) -> GradSampleModule:
    assert grad_sample_mode == "hooks", "only hook-based per-sample gradients are implemented"
    return GradSampleModule(model, *args, **kwargs)


# ---------------------------------------------------------------------------------------------------------------
# DPOptimizer
# ---------------------------------------------------------------------------------------------------------------
class DPOptimizer(Optimizer):
    def __init__(self, optimizer: Optimizer, *, noise_multiplier: float, max_grad_norm: float,
                 expected_batch_size: int | None, loss_reduction: str = "mean", generator: torch.Generator | None = None,
                 module: nn.Module | None = None) -> None:
        self.original_optimizer = optimizer
        self.noise_multiplier = noise_multiplier
        self.max_grad_norm = max_grad_norm
        self.expected_batch_size = expected_batch_size
        self.loss_reduction = loss_reduction
        self.generator = generator
        self.param_groups = optimizer.param_groups
        self.defaults = optimizer.defaults
        self.state = optimizer.state
        self._step_count = 0
        self._module = module

    @property
    def params(self) -> list[nn.Parameter]:
        return [p for group in self.param_groups for p in group["params"] if p.requires_grad]

    def clip_and_accumulate(self) -> bool:
        params = [p for p in self.params if getattr(p, "grad_sample", None) is not None]
        if not params:
            return False
        per_param_norms = [p.grad_sample.reshape(p.grad_sample.shape[0], -1).norm(2, dim=1) for p in params]  # type: ignore[attr-defined]
        per_sample_norms = torch.stack(per_param_norms, dim=1).norm(2, dim=1)
        clip_factor = (self.max_grad_norm / (per_sample_norms + 1e-6)).clamp(max=1.0)
        for p in params:
            summed = torch.einsum("i,i...", clip_factor.to(p.grad_sample.dtype), p.grad_sample)  # type: ignore[attr-defined]
            if p.grad is not None:
                p.grad.copy_(summed.to(p.grad.dtype))  # keeps arena gradient views alive
            else:
                p.grad = summed
            p.grad_sample = None  # type: ignore[attr-defined]
        return True

    def add_noise_and_scale(self) -> None:
        std = self.noise_multiplier * self.max_grad_norm
        scale = 1.0
        if self.loss_reduction == "mean" and self.expected_batch_size:
            scale = 1.0 / float(self.expected_batch_size)
        arena = arena_of(self._module) if self._module is not None else None
        inner = getattr(self._module, "_module", None)
        if arena is None and inner is not None:
            arena = arena_of(inner)
        params = self.params
        if (arena is not None and arena.grad is not None and std > 0
                and all(p.grad is not None and p.grad.untyped_storage().data_ptr() == arena.grad.untyped_storage().data_ptr() for p in params)):
            seed = int(torch.randint(0, 2**62, (1,), generator=self.generator).item())
            flat_ops.add_gaussian_(arena.grad, std, seed)  # one kernel over the whole flat gradient
            arena.grad.mul_(scale)
            return
        for p in params:
            if p.grad is None:
                continue
            if std > 0:
                p.grad.add_(torch.normal(0.0, std, p.grad.shape, device=p.grad.device, dtype=p.grad.dtype, generator=None))
            p.grad.mul_(scale)

    def pre_step(self) -> bool:
        if not self.clip_and_accumulate():
            return False
        self.add_noise_and_scale()
        return True

    def step(self, closure: Any = None) -> Any:  # type: ignore[override]
        loss = closure() if closure is not None else None
        if self.pre_step():
            self._step_count += 1
            return self.original_optimizer.step()
        return loss

    def zero_grad(self, set_to_none: bool = False) -> None:
        for p in self.params:
            p.grad_sample = None  # type: ignore[attr-defined]
        self.original_optimizer.zero_grad(set_to_none) if not _is_flat(self.original_optimizer) else self.original_optimizer.zero_grad()

    def state_dict(self) -> dict[str, Any]:  # type: ignore[override]
        return self.original_optimizer.state_dict()

    def load_state_dict(self, state_dict: dict[str, Any]) -> None:  # type: ignore[override]
        self.original_optimizer.load_state_dict(state_dict)

    def __repr__(self) -> str:
        return f"DPOptimizer({self.original_optimizer!r}, sigma={self.noise_multiplier}, C={self.max_grad_norm})"


def _is_flat(optimizer: Optimizer) -> bool:
    from fl4health_b200.engine.fused_optim import _FlatOptimizer

    return isinstance(optimizer, _FlatOptimizer)


# ---------------------------------------------------------------------------------------------------------------
# Poisson data loader
# ---------------------------------------------------------------------------------------------------------------
class DPDataLoader:
    """Poisson-sampled batches over any loader that exposes ``dataset`` and ``batch_size``."""

    def __init__(self, dataset: Any, sample_rate: float, generator: torch.Generator | None = None, collate_fn: Any = None) -> None:
        self.dataset = dataset
        self.sample_rate = sample_rate
        self.generator = generator
        self.collate_fn = collate_fn
        self.batch_size = max(1, int(round(sample_rate * len(dataset))))

    @classmethod
    def from_data_loader(cls, data_loader: Any, generator: torch.Generator | None = None) -> DPDataLoader:
        n = len(data_loader.dataset)
        return cls(data_loader.dataset, sample_rate=data_loader.batch_size / n, generator=generator,
                   collate_fn=getattr(data_loader, "collate_fn", None))

    def __len__(self) -> int:
        return max(1, int(round(1.0 / self.sample_rate)))

    def __iter__(self) -> Iterator[Any]:
        n = len(self.dataset)
        for _ in range(len(self)):
            mask = torch.rand(n, generator=self.generator) < self.sample_rate
            indices = mask.nonzero(as_tuple=False).reshape(-1)
            get_batch = getattr(self.dataset, "get_batch", None)
            if get_batch is not None and indices.numel() > 0:
                yield get_batch(indices.to(self.dataset.data.device))
                continue
            samples = [self.dataset[int(i)] for i in indices]
            if not samples:
                first = self.dataset[0]
                yield torch.empty((0, *first[0].shape), dtype=first[0].dtype), torch.empty((0,), dtype=torch.as_tensor(first[1]).dtype)
                continue
            if self.collate_fn is not None:
                yield self.collate_fn(samples)
            else:
                yield torch.stack([s[0] for s in samples]), torch.stack([torch.as_tensor(s[1]) for s in samples])


# ---------------------------------------------------------------------------------------------------------------
# PrivacyEngine
# ---------------------------------------------------------------------------------------------------------------
class PrivacyEngine:
    def __init__(self, secure_mode: bool = False) -> None:
        if secure_mode:
            log(WARNING, "secure_mode (cryptographic RNG) is not implemented; using the default generator")

    def make_private(
        self, *, module: nn.Module, optimizer: Optimizer, data_loader: Any, noise_multiplier: float,
        max_grad_norm: float, batch_first: bool = True, loss_reduction: str = "mean", poisson_sampling: bool = True,
        clipping: str = "flat", noise_generator: torch.Generator | None = None, **kwargs: Any,
    ) -> tuple[GradSampleModule, DPOptimizer, Any]:
        assert clipping == "flat", "only flat clipping is implemented"
        wrapped = module if isinstance(module, GradSampleModule) else GradSampleModule(module, batch_first, loss_reduction)
        expected_batch_size = int(data_loader.batch_size) if getattr(data_loader, "batch_size", None) else None
        dp_loader = DPDataLoader.from_data_loader(data_loader, noise_generator) if poisson_sampling else data_loader
        dp_optimizer = DPOptimizer(optimizer, noise_multiplier=noise_multiplier, max_grad_norm=max_grad_norm,
                                   expected_batch_size=expected_batch_size, loss_reduction=loss_reduction,
                                   generator=noise_generator, module=wrapped)
        return wrapped, dp_optimizer, dp_loader
