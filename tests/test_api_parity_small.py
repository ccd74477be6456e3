"""Small public names kept for import-path parity with the reference (found by diffing module-level public names)."""

import pandas as pd
import pytest
import torch

from fl4health_b200.feature_alignment.constants import FeatureType
from fl4health_b200.feature_alignment.handle_types import to_dtype, valid_feature_type


def test_handle_types_helpers() -> None:
    assert valid_feature_type(FeatureType.ORDINAL)
    assert not valid_feature_type(FeatureType.CATEGORICAL_INDICATOR, raise_error=False)
    with pytest.raises(ValueError, match="categorical_indicator"):
        valid_feature_type(FeatureType.CATEGORICAL_INDICATOR)
    series = pd.Series([0, 1, 1, 0])
    assert str(to_dtype(series, FeatureType.BINARY).dtype) == "category"
    assert to_dtype(series, FeatureType.NUMERIC) is series
    strings = pd.Series(["a", "b"])
    assert to_dtype(strings, FeatureType.STRING) is strings  # strings keep whatever dtype the caller chose


def test_skin_cancer_path_and_label_functions() -> None:
    from fl4health_b200.datasets.skin_cancer import preprocess_skin as ps

    ham = pd.Series({"image_id": "ISIC_0001", "dx": "akiec"})
    assert ps.ham_image_path_func(ham).endswith("HAM10000/ISIC_0001.jpg") and ps.ham_label_map_func(ham) == "AK"
    pad = pd.Series({"img_id": "PAT_1.png", "diagnostic": "SEK"})
    assert ps.pad_image_path_func(pad).endswith("PAD-UFES-20/PAT_1.png") and ps.pad_label_map_func(pad) == "BKL"
    derm = pd.Series({"derm": "a/b.jpg", "diagnosis": "melanoma (in situ)"})
    assert ps.derm7pt_image_path_func(derm).endswith("Derm7pt/images/a/b.jpg") and ps.derm7pt_label_map_func(derm) == "MEL"


def test_rxrx1_save_to_pkl(tmp_path) -> None:
    import pickle

    from fl4health_b200.datasets.rxrx1.preprocess import save_to_pkl

    save_to_pkl(torch.arange(6).view(2, 3), str(tmp_path / "t.pkl"))
    assert torch.equal(pickle.load(open(tmp_path / "t.pkl", "rb")), torch.arange(6).view(2, 3))


def test_protocol_compliance_decorator_and_protocol_names() -> None:
    from fl4health_b200.clients.flexible.base import FlexibleClient
    from fl4health_b200.mixins.adaptive_drift_constrained import AdaptiveDriftConstrainedProtocol
    from fl4health_b200.mixins.personalized.ditto import DittoPersonalizedProtocol
    from fl4health_b200.mixins.personalized.mr_mtl import MrMtlPersonalizedProtocol
    from fl4health_b200.mixins.personalized.utils import ensure_protocol_compliance

    assert AdaptiveDriftConstrainedProtocol in DittoPersonalizedProtocol.__mro__
    assert AdaptiveDriftConstrainedProtocol in MrMtlPersonalizedProtocol.__mro__

    class NotAClient:
        @ensure_protocol_compliance
        def method(self) -> int:
            return 1

    with pytest.raises(TypeError, match="Protocol requirements not met"):
        NotAClient().method()

    class Flex(FlexibleClient):
        @ensure_protocol_compliance
        def method(self) -> int:
            return 2

    assert Flex.method(Flex.__new__(Flex)) == 2


def test_contracts_are_checkable_and_name_what_is_missing() -> None:
    from fl4health_b200.clients.basic_client import BasicClient
    from fl4health_b200.clients.flexible.base import FlexibleClient
    from fl4health_b200.mixins.adaptive_drift_constrained import AdaptiveDriftConstrainedMixin
    from fl4health_b200.mixins.core_protocols import (
        AdaptiveDriftConstrainedProtocol,
        DittoPersonalizedProtocol,
        FlexibleClientProtocol,
        MrMtlPersonalizedProtocol,
        NumPyClientMinimalProtocol,
    )
    from fl4health_b200.mixins.personalized import PersonalizedMode, make_it_personal

    assert issubclass(FlexibleClient, FlexibleClientProtocol) and not issubclass(FlexibleClient, DittoPersonalizedProtocol)
    assert issubclass(make_it_personal(FlexibleClient, PersonalizedMode.DITTO), DittoPersonalizedProtocol)
    assert issubclass(make_it_personal(FlexibleClient, PersonalizedMode.MR_MTL), MrMtlPersonalizedProtocol)

    class Constrained(AdaptiveDriftConstrainedMixin, FlexibleClient):
        pass

    assert AdaptiveDriftConstrainedProtocol.missing(Constrained) == []
    # the non-flexible client speaks the NumPy-client part but not the per-model step family
    assert NumPyClientMinimalProtocol.missing(BasicClient) == []
    assert "_train_step_with_model_and_optimizer" in FlexibleClientProtocol.missing(BasicClient)
    assert not isinstance(object(), NumPyClientMinimalProtocol)
    # instances are also held to the attributes; inherited requirements accumulate, each name once
    assert {"device", "model", "global_model"} <= set(DittoPersonalizedProtocol.all_attributes())
    assert len(set(DittoPersonalizedProtocol.all_methods())) == len(DittoPersonalizedProtocol.all_methods())
    with pytest.raises(TypeError, match="lacks .*fit.*NumPyClientMinimalProtocol"):
        NumPyClientMinimalProtocol.require(object())


def test_constants_exist() -> None:
    from fl4health_b200.clients.basic_client import EXPECTED_OUTPUT_TUPLE_SIZE
    from fl4health_b200.model_bases.ensemble_base import EXPECTED_MAX_PRED_N_DIMS
    from fl4health_b200.model_bases.masked_layers.masked_normalization_layers import BATCH_NORM_1D_INPUT_LENGTHS
    from fl4health_b200.model_bases.pca import TWO_D_TENSOR_SHAPE_LENGTH
    from fl4health_b200.servers.nnunet_server import EVAL_CFG_FN, FIT_CFG_FN  # noqa: F401

    assert (EXPECTED_OUTPUT_TUPLE_SIZE, EXPECTED_MAX_PRED_N_DIMS, TWO_D_TENSOR_SHAPE_LENGTH) == (2, 2, 2)
    assert BATCH_NORM_1D_INPUT_LENGTHS == {2, 3}


def test_dp_events_and_mkmmd_decomposition_helpers() -> None:
    from fl4health_b200.losses.mkmmd_loss import MkMmdLoss
    from fl4health_b200.privacy.dp_events import GaussianDpEvent, NeighborRel, PoissonSampledDpEvent, SelfComposedDpEvent
    from fl4health_b200.privacy.moments_accountant import FixedSamplingWithoutReplacement, PoissonSampling

    event = PoissonSampling(0.25).composed_event(1.5, 40)
    assert event == SelfComposedDpEvent(PoissonSampledDpEvent(0.25, GaussianDpEvent(1.5)), 40)
    assert FixedSamplingWithoutReplacement(100, 10).neighbor_relation is NeighborRel.REPLACE_ONE
    assert FixedSamplingWithoutReplacement(100, 10).get_dp_event(GaussianDpEvent(1.0)).sample_size == 10

    torch.manual_seed(0)
    loss = MkMmdLoss(device=torch.device("cpu"), gammas=torch.tensor([0.5, 2.0, 8.0]))
    x, y = torch.randn(8, 5), torch.randn(8, 5) + 0.5
    distances = loss.compute_euclidean_inner_products(x, y)
    all_h = loss.compute_all_h_u_from_inner_products(distances)
    assert torch.allclose(all_h, loss.compute_all_h_u_all_samples(x, y))
    per_kernel = torch.cat([loss.compute_h_u_from_inner_products(distances, g.reshape(1)) for g in loss.gammas])
    assert torch.allclose(per_kernel, all_h, atol=1e-6)
    quads = loss.compute_euclidean_inner_products_linear(loss.construct_quadruples(x, y))
    lin = loss.compute_all_h_u_from_inner_products_linear(quads)
    assert torch.allclose(lin, loss.compute_all_h_u_linear(x, y))
    assert torch.allclose(torch.cat([loss.compute_h_u_from_inner_products_linear(quads, g.reshape(1)) for g in loss.gammas]), lin, atol=1e-6)
    hat_d = loss.compute_hat_d_per_kernel(all_h)
    centred = loss.form_kernel_samples_minus_expectation(all_h, hat_d)
    assert centred.shape == all_h.shape and torch.allclose(centred.reshape(3, -1).mean(dim=1), torch.zeros(3), atol=1e-6)
