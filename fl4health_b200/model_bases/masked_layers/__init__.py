from fl4health_b200.model_bases.masked_layers.masked_layers import (  # noqa: F401
    MaskedBatchNorm1d,
    MaskedBatchNorm2d,
    MaskedBatchNorm3d,
    MaskedConv1d,
    MaskedConv2d,
    MaskedConv3d,
    MaskedConvTranspose1d,
    MaskedConvTranspose2d,
    MaskedConvTranspose3d,
    MaskedLayerNorm,
    MaskedLinear,
)
from fl4health_b200.model_bases.masked_layers.masked_layers_utils import convert_to_masked_model, is_masked_module  # noqa: F401
