"""Top-k% sparse exchange in per-tensor COO form.

Parity: ``fl4health/parameter_exchange/sparse_coo_parameter_exchanger.py:18-162``.  The global score threshold is found
with ``torch.kthvalue``/``topk`` on the device rather than a full ``torch.sort`` of every parameter (:94-101), and the
COO pieces stay on the device.
"""

from __future__ import annotations

import math
from collections.abc import Callable
from logging import INFO, WARNING

import numpy as np
import torch
from torch import Tensor, nn

from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import Config, NDArrays, to_tensor
from fl4health_b200.parameter_exchange._state import inject_state
from fl4health_b200.parameter_exchange.parameter_packer import SparseCooParameterPacker
from fl4health_b200.parameter_exchange.partial_parameter_exchanger import PartialParameterExchanger

ScoreGenFunction = Callable[[nn.Module, "nn.Module | None"], dict[str, Tensor]]


class SparseCooParameterExchanger(PartialParameterExchanger[tuple[NDArrays, NDArrays, list[str]]]):
    def __init__(self, sparsity_level: float, score_gen_function: ScoreGenFunction) -> None:
        assert 0 < sparsity_level <= 1
        self.sparsity_level = sparsity_level
        self.parameter_packer: SparseCooParameterPacker = SparseCooParameterPacker()
        self.score_gen_function = score_gen_function

    def generate_parameter_scores(self, model: nn.Module, initial_model: nn.Module | None) -> dict[str, Tensor]:
        return self.score_gen_function(model, initial_model)

    def _check_unique_score(self, param_scores: Tensor) -> None:
        if param_scores.numel() > 0 and bool((param_scores == param_scores.reshape(-1)[0]).all()):
            log(
                WARNING,
                "All parameters have the same score.\nThe number of parameters selected may not match the intended"
                " sparsity level.",
            )

    def select_parameters(
        self, model: nn.Module, initial_model: nn.Module | None = None
    ) -> tuple[NDArrays, tuple[NDArrays, NDArrays, list[str]]]:
        scores = self.generate_parameter_scores(model, initial_model)
        all_scores = torch.cat([s.reshape(-1).float() for s in scores.values()])
        n_top = math.ceil(all_scores.numel() * self.sparsity_level)
        assert n_top >= 1
        # k-th largest == (N - k + 1)-th smallest: selection, not a full sort
        threshold = torch.kthvalue(all_scores, all_scores.numel() - n_top + 1).values

        values, indices, shapes, names = NDArrays(), NDArrays(), NDArrays(), []
        states = model.state_dict()
        for name, param_scores in scores.items():
            tensor = states[name]
            assert tensor.shape == param_scores.shape
            self._check_unique_score(param_scores)
            sparse = torch.where(param_scores >= threshold, tensor, torch.zeros((), dtype=tensor.dtype, device=tensor.device))
            if sparse.dim() == 0 or not bool((sparse != 0).any()):
                continue
            vals, idx, shape = self.parameter_packer.extract_coo_info_from_dense(sparse)
            values.append(vals)
            indices.append(idx)
            shapes.append(shape)
            names.append(name)
        log(INFO, f"Sparsity level used to select parameters for exchange: {self.sparsity_level}")
        return values, (indices, shapes, names)

    def push_parameters(
        self, model: nn.Module, initial_model: nn.Module | None = None, config: Config | None = None
    ) -> NDArrays:
        selected, additional = self.select_parameters(model, initial_model)
        return self.pack_parameters(model_weights=selected, additional_parameters=additional)

    def pull_parameters(self, parameters: NDArrays, model: nn.Module, config: Config | None = None) -> None:
        values, (indices, shapes, names) = self.parameter_packer.unpack_parameters(parameters)
        assert len(values) == len(indices) == len(shapes) == len(names) and len(names) > 0
        state = model.state_dict()
        dense = []
        for vals, idx, shape, name in zip(values, indices, shapes, names):
            target = state[name]
            v = to_tensor(vals, target.device).to(target.dtype)
            i = to_tensor(idx, target.device).long()
            out = torch.zeros(tuple(int(s) for s in np.asarray(shape).tolist()), dtype=target.dtype, device=target.device)
            if i.numel() > 0:
                out[tuple(i.t())] = v
            dense.append(out)
        inject_state(model, names, dense)
