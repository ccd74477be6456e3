"""PI-CAI preprocessing on tensors + focal loss (modelled on the reference's ``tests/test_picai/test_case.py`` and
``tests/test_picai/test_preprocess_transforms.py``)."""

from __future__ import annotations

import pytest
import torch

from research.picai.losses import FocalLoss
from research.picai.preprocessing import (
    AlignOriginAndDirection,
    BinarizeAnnotation,
    Case,
    CentreCropAndOrPad,
    PreprocessingError,
    PreprocessingSettings,
    ResampleSpacing,
    ResampleToFirstScan,
    Volume,
    ZScoreNormalise,
    apply_transform,
    centre_crop_or_pad,
    default_transforms,
    preprocess,
    resample,
)


def settings(tmp_path, **kwargs) -> PreprocessingSettings:  # type: ignore[no-untyped-def]
    return PreprocessingSettings(tmp_path / "scans", tmp_path / "labels", **kwargs)


def ramp(size: tuple[int, int, int], spacing=(1.0, 1.0, 1.0), origin=(0.0, 0.0, 0.0)) -> Volume:  # type: ignore[no-untyped-def]
    """f(z, y, x) = 100 z + 10 y + x in PHYSICAL coordinates: linear, so trilinear resampling reproduces it exactly."""
    axes = [origin[a] + torch.arange(size[a], dtype=torch.float32) * spacing[a] for a in range(3)]
    z, y, x = torch.meshgrid(*axes, indexing="ij")
    return Volume(100 * z + 10 * y + x, spacing, origin)


def test_settings_infer_and_validate_geometry(tmp_path) -> None:  # type: ignore[no-untyped-def]
    inferred = settings(tmp_path, size=(20, 256, 256), spacing=(3.0, 0.5, 0.5))
    assert inferred.physical_size == (60.0, 128.0, 128.0)
    assert settings(tmp_path, physical_size=(60.0, 128.0, 128.0), spacing=(3.0, 0.5, 0.5)).size == (20, 256, 256)
    settings(tmp_path, size=(2, 4, 4), spacing=(1.0, 0.5, 0.5), physical_size=(2.0, 2.0, 2.0))
    with pytest.raises(AssertionError):
        settings(tmp_path, size=(2, 4, 4), spacing=(1.0, 0.5, 0.5), physical_size=(2.0, 2.0, 3.0))


def test_resample_reproduces_a_linear_field_and_handles_labels() -> None:
    coarse = ramp((5, 6, 7), spacing=(2.0, 1.5, 1.0), origin=(1.0, -2.0, 3.0))
    fine = resample(coarse, (9, 11, 13), (1.0, 0.75, 0.5), coarse.origin)
    assert torch.allclose(fine.data, ramp((9, 11, 13), (1.0, 0.75, 0.5), coarse.origin).data, atol=1e-3)
    same = resample(coarse, coarse.size, coarse.spacing, coarse.origin)
    assert torch.allclose(same.data, coarse.data, atol=1e-4)
    labels = Volume(torch.zeros(4, 4, 4, dtype=torch.int64), (1.0, 1.0, 1.0))
    labels.data[1:3, 1:3, 1:3] = 3
    up = resample(labels, (8, 8, 8), (0.5, 0.5, 0.5), labels.origin, nearest=True)
    assert up.data.dtype == torch.int64 and set(up.data.unique().tolist()) == {0, 3}  # no interpolated label values
    shifted = resample(coarse, coarse.size, coarse.spacing, (101.0, -2.0, 3.0))  # entirely outside the source
    assert torch.count_nonzero(shifted.data) == 0


def test_resample_to_first_scan_and_alignment(tmp_path) -> None:  # type: ignore[no-untyped-def]
    t2w = ramp((6, 16, 16), (3.0, 0.5, 0.5))
    adc = ramp((6, 4, 4), (3.0, 2.0, 2.0))  # same field of view, 4x coarser in-plane
    annotation = Volume(torch.zeros(6, 4, 4, dtype=torch.int64), (3.0, 2.0, 2.0))
    annotation.data[2:4, 1:3, 1:3] = 2
    case = AlignOriginAndDirection()(ResampleToFirstScan()(Case([t2w, adc], annotation, settings(tmp_path))))
    assert all(s.size == (6, 16, 16) and s.spacing == t2w.spacing for s in case.scans) and case.annotation.size == (6, 16, 16)
    inside = (slice(None), slice(0, 13), slice(0, 13))  # the coarse scan covers physical [0, 6] mm in-plane
    assert torch.allclose(case.scans[1].data[inside], t2w.data[inside], atol=1e-3)
    assert set(case.annotation.data.unique().tolist()) == {0, 2}
    with pytest.raises(PreprocessingError):
        AlignOriginAndDirection()(Case([t2w, adc], annotation, settings(tmp_path)))  # grids were never brought together
    far = Volume(t2w.data.clone(), t2w.spacing, (50.0, 0.0, 0.0))
    with pytest.raises(PreprocessingError):
        AlignOriginAndDirection()(Case([t2w, far], Volume(torch.zeros(6, 16, 16), t2w.spacing), settings(tmp_path)))


def test_resample_spacing_keeps_physical_extent(tmp_path) -> None:  # type: ignore[no-untyped-def]
    scan = ramp((4, 8, 8), (3.0, 1.0, 1.0))
    annotation = Volume((scan.data > 500).long(), scan.spacing)
    case = ResampleSpacing()(Case([scan], annotation, settings(tmp_path, spacing=(1.5, 0.5, 0.5))))
    assert case.scans[0].size == (8, 16, 16) and case.scans[0].spacing == (1.5, 0.5, 0.5)
    assert case.scans[0].physical_size == scan.physical_size
    assert ResampleSpacing()(Case([scan], annotation, settings(tmp_path))).scans[0] is scan  # no target spacing: untouched


def test_centre_crop_and_pad(tmp_path) -> None:  # type: ignore[no-untyped-def]
    scan = ramp((4, 10, 6))
    out = centre_crop_or_pad(scan, (6, 4, 6))
    assert out.size == (6, 4, 6)
    assert torch.equal(out.data[1:5], scan.data[:, 3:7, :])          # depth padded by (1, 1); height cropped from index 3
    assert torch.count_nonzero(out.data[0]) == 0 and torch.count_nonzero(out.data[5]) == 0
    assert out.origin == (-1.0, 3.0, 0.0)                              # physical position of the new voxel (0, 0, 0)
    odd = centre_crop_or_pad(Volume(torch.ones(3, 3, 3)), (4, 3, 3))  # odd padding puts the extra slice at the end
    assert odd.data[:, 0, 0].tolist() == [1.0, 1.0, 1.0, 0.0]
    case = CentreCropAndOrPad()(Case([scan], Volume(torch.ones(4, 10, 6)), settings(tmp_path, size=(2, 12, 6))))
    assert case.scans[0].size == case.annotation.size == (2, 12, 6)


def test_binarise_and_normalise(tmp_path) -> None:  # type: ignore[no-untyped-def]
    annotation = Volume(torch.tensor([[[0, 1], [2, 5]]]))
    scan = Volume(torch.tensor([[[0.0, 2.0], [4.0, 6.0]]]))
    case = ZScoreNormalise()(BinarizeAnnotation()(Case([scan], annotation, settings(tmp_path))))
    assert case.annotation.data.flatten().tolist() == [0, 1, 1, 1] and case.annotation.data.dtype == annotation.data.dtype
    values = case.scans[0].data.flatten()
    assert values[0] == 0 and abs(float(values[1:].mean())) < 1e-6 and float(values[1:].std()) == pytest.approx(1.0)


def test_full_pipeline_writes_and_reads_cases(tmp_path) -> None:  # type: ignore[no-untyped-def]
    target = settings(tmp_path, size=(8, 12, 12), spacing=(1.5, 0.5, 0.5))

    def make(case_id: str, broken: bool = False) -> Case:
        t2w = ramp((4, 16, 16), (3.0, 0.5, 0.5))
        adc = ramp((4, 4, 4), (3.0, 2.0, 2.0), origin=(40.0 if broken else 0.0, 0.0, 0.0))
        annotation = Volume(torch.zeros(4, 16, 16, dtype=torch.int64), (3.0, 0.5, 0.5))
        annotation.data[1:3, 6:10, 6:10] = 4
        return Case([t2w, adc], annotation, target, case_id=case_id)

    written = preprocess([make("10000_1000000"), make("10001_1000001")])
    assert len(written) == 2
    scan_paths, annotation_path = written[0]
    assert [p.name for p in scan_paths] == ["10000_1000000_0000.pt", "10000_1000000_0001.pt"] and annotation_path.name == "10000_1000000.pt"
    case = Case.read(scan_paths, annotation_path, target)
    assert all(s.size == (8, 12, 12) and s.spacing == (1.5, 0.5, 0.5) for s in [*case.scans, case.annotation])
    assert set(case.annotation.data.unique().tolist()) == {0, 1} and case.annotation.data.sum() > 0
    # a sequence acquired somewhere else entirely resamples to zeros — consistent grids, so it is kept; a case whose
    # annotation grid cannot be aligned is skipped with a warning instead of aborting the whole study list
    bad = make("bad")
    bad.annotation = Volume(torch.zeros(4, 16, 16, dtype=torch.int64), (3.0, 0.5, 0.5), origin=(0.0, 30.0, 0.0))
    transforms = [AlignOriginAndDirection(), *default_transforms()]
    assert preprocess([bad], transforms) == []
    with pytest.raises(PreprocessingError):
        apply_transform(bad, transforms)


def test_focal_loss_matches_the_textbook_formula() -> None:
    torch.manual_seed(0)
    logits, targets = torch.randn(2, 1, 4, 5, 5), (torch.rand(2, 1, 4, 5, 5) > 0.8).float()
    p = torch.sigmoid(logits)
    ce = -(targets * p.log() + (1 - targets) * (1 - p).log())
    p_t = p * targets + (1 - p) * (1 - targets)
    for alpha, gamma in ((1.0, 1.0), (0.75, 2.0), (-1.0, 0.5)):
        expected = ce * (1 - p_t) ** gamma
        if alpha >= 0:
            expected = (alpha * targets + (1 - alpha) * (1 - targets)) * expected
        assert torch.allclose(FocalLoss(alpha, gamma)(logits, targets), expected.sum(), rtol=1e-4)
        assert torch.allclose(FocalLoss(alpha, gamma, "mean")(logits, targets), expected.mean(), rtol=1e-4)
    assert float(FocalLoss(-1.0, 0.0)(logits, targets)) == pytest.approx(float(ce.sum()), rel=1e-4)  # gamma 0: plain BCE
    extreme = FocalLoss()(torch.tensor([80.0, -80.0]), torch.tensor([0.0, 1.0]))  # stable where sigmoid saturates
    assert torch.isfinite(extreme) and float(extreme) > 0  # alpha=1 weights only the positive (second) voxel
    with pytest.raises(NotImplementedError):
        FocalLoss(reduction="none")
