"""Re-export for import-path parity with the reference (implementation: masked_layers.py)."""
from fl4health_b200.model_bases.masked_layers.masked_layers import *  # noqa: F401,F403
from fl4health_b200.model_bases.masked_layers.masked_layers import _MaskedBatchNorm  # noqa: F401

import torch as _torch

TorchShape = int | list[int] | _torch.Size  # accepted ``normalized_shape`` forms of MaskedLayerNorm

BATCH_NORM_3D_INPUT_LENGTH = 5
BATCH_NORM_2D_INPUT_LENGTH = 4
BATCH_NORM_1D_INPUT_LENGTHS = {2, 3}
