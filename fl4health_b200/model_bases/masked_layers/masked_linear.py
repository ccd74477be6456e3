"""Re-export for import-path parity with the reference (implementation: masked_layers.py)."""
from fl4health_b200.model_bases.masked_layers.masked_layers import *  # noqa: F401,F403
from fl4health_b200.model_bases.masked_layers.masked_layers import _MaskedBatchNorm  # noqa: F401
