"""The onboarding walkthrough (docs/onboarding/, examples/onboarding_tabular.py) is executed, so its text cannot drift
from the code: federated training across skewed sites must clearly beat every site training alone and land near the
pooled-data model."""

from __future__ import annotations

from examples.onboarding_tabular import N_CLASSES, make_cohort, walkthrough

import torch


def test_cohorts_are_skewed_but_share_the_relationship() -> None:
    generator = torch.Generator().manual_seed(0)
    low, high = make_cohort(2000, -0.8, generator), make_cohort(2000, 0.8, generator)
    mix_low = torch.bincount(low[1], minlength=N_CLASSES) / 2000
    mix_high = torch.bincount(high[1], minlength=N_CLASSES) / 2000
    assert (mix_low - mix_high).abs().max() > 0.2  # the outcome mix differs a lot between sites


def test_walkthrough_numbers_tell_the_story() -> None:
    results = walkthrough(hospitals=3, rounds=10, patients_per_hospital=400)
    central = results["centralized (pooled data)"]
    alone = results["each hospital alone (mean)"]
    federated = results["federated (FedAvg, data stays put)"]
    assert federated > alone + 0.05, results
    assert federated > central - 0.08, results
