"""Onboarding walkthrough: the same tabular classifier trained (1) centrally on pooled data and (2) federated across
hospitals that each hold a skewed slice of it (the text is in ``docs/onboarding/``; role of the reference's
``vector-bootcamp-2025/tabular_centralized_training`` notebook, data generated instead of shipped).

    python -m examples.onboarding_tabular [--hospitals 4] [--rounds 15] [--device cuda:0]
"""

from __future__ import annotations

import argparse
from pathlib import Path

import torch
from torch import nn

from fl4health_b200.clients.basic_client import BasicClient
from fl4health_b200.common.typing import Config
from fl4health_b200.engine.data import BatchedTensorLoader
from fl4health_b200.metrics import Accuracy
from fl4health_b200.metrics.metric_aggregation import evaluate_metrics_aggregation_fn, fit_metrics_aggregation_fn
from fl4health_b200.servers.base_server import FlServer
from fl4health_b200.servers.client_manager import SimpleClientManager
from fl4health_b200.simulation import run_simulation
from fl4health_b200.strategies.basic_fedavg import BasicFedAvg
from fl4health_b200.utils.dataset import TensorDataset
from fl4health_b200.utils.random import set_all_random_seeds

N_FEATURES, N_CLASSES = 12, 3


def make_cohort(n_patients: int, site_shift: float, generator: torch.Generator) -> tuple[torch.Tensor, torch.Tensor]:
    """Vitals / lab values -> one of three outcomes.  Every site measures the same relationship, but its population is
    shifted (``site_shift`` moves the feature means, which also changes the outcome mix the site sees)."""
    features = torch.randn(n_patients, N_FEATURES, generator=generator) + site_shift
    scores = torch.stack([
        features[:, :4].sum(dim=1),
        features[:, 4:8].sum(dim=1) * 0.9 + 0.5,
        (features[:, 8:] ** 2).sum(dim=1) * 0.35,
    ], dim=1)
    return features, scores.argmax(dim=1)


def build_model() -> nn.Module:
    return nn.Sequential(nn.Linear(N_FEATURES, 64), nn.ReLU(), nn.Linear(64, 32), nn.ReLU(), nn.Linear(32, N_CLASSES))


def accuracy(model: nn.Module, features: torch.Tensor, labels: torch.Tensor) -> float:
    with torch.no_grad():
        return float((model(features).argmax(dim=1) == labels).float().mean())


def train_centrally(train: tuple[torch.Tensor, torch.Tensor], steps: int, batch_size: int, lr: float) -> nn.Module:
    """Step 1 of the walkthrough: the plain PyTorch loop everybody starts from."""
    model, features, labels = build_model(), *train
    optimizer = torch.optim.SGD(model.parameters(), lr=lr, momentum=0.9)
    for step in range(steps):
        rows = torch.randint(0, len(features), (batch_size,))
        optimizer.zero_grad()
        nn.functional.cross_entropy(model(features[rows]), labels[rows]).backward()
        optimizer.step()
    return model


class HospitalClient(BasicClient):
    """Step 2: the same model / optimizer / loss, handed to the framework through four hooks.  The data never leaves
    the object; only parameters travel."""

    def __init__(self, cohort: tuple[torch.Tensor, torch.Tensor], **kwargs) -> None:  # noqa: ANN003
        super().__init__(**kwargs)
        n_val = len(cohort[0]) // 5
        self.cohort_train = (cohort[0][n_val:], cohort[1][n_val:])
        self.cohort_val = (cohort[0][:n_val], cohort[1][:n_val])

    def get_model(self, config: Config) -> nn.Module:
        return build_model()

    def get_data_loaders(self, config: Config):  # noqa: ANN201
        batch_size = int(config["batch_size"])
        return (BatchedTensorLoader(TensorDataset(*self.cohort_train), batch_size, shuffle=True, device=self.device),
                BatchedTensorLoader(TensorDataset(*self.cohort_val), batch_size, device=self.device))

    def get_optimizer(self, config: Config) -> torch.optim.Optimizer:
        return torch.optim.SGD(self.model.parameters(), lr=float(config["lr"]), momentum=0.9)

    def get_criterion(self, config: Config) -> nn.Module:
        return nn.CrossEntropyLoss()


def train_federated(cohorts: list[tuple[torch.Tensor, torch.Tensor]], rounds: int, local_steps: int, batch_size: int,
                    lr: float, device: torch.device) -> tuple[nn.Module, list[HospitalClient]]:
    def round_config(server_round: int) -> Config:
        return {"current_server_round": server_round, "local_steps": local_steps, "batch_size": batch_size, "lr": lr}

    clients = [
        HospitalClient(cohort, data_path=Path("."), metrics=[Accuracy()], device=device, client_name=f"hospital_{index}")
        for index, cohort in enumerate(cohorts)
    ]
    strategy = BasicFedAvg(
        min_fit_clients=len(clients), min_evaluate_clients=len(clients), min_available_clients=len(clients),
        on_fit_config_fn=round_config, on_evaluate_config_fn=round_config,
        fit_metrics_aggregation_fn=fit_metrics_aggregation_fn,
        evaluate_metrics_aggregation_fn=evaluate_metrics_aggregation_fn,
    )
    server = FlServer(SimpleClientManager(), {"n_server_rounds": rounds}, strategy, on_init_parameters_config_fn=round_config)
    run_simulation(server, clients, num_rounds=rounds)
    # after the last round every client holds the aggregated model it just evaluated
    return clients[0].model, clients


def walkthrough(hospitals: int = 4, rounds: int = 15, patients_per_hospital: int = 600, device: str = "cpu",
                seed: int = 2025) -> dict[str, float]:
    set_all_random_seeds(seed)
    generator = torch.Generator().manual_seed(seed)
    shifts = torch.linspace(-0.8, 0.8, hospitals).tolist()
    cohorts = [make_cohort(patients_per_hospital, shift, generator) for shift in shifts]
    held_out = [make_cohort(400, shift, generator) for shift in shifts]
    test_x, test_y = torch.cat([c[0] for c in held_out]), torch.cat([c[1] for c in held_out])

    local_steps, batch_size, lr = 10, 32, 0.05
    budget = rounds * local_steps  # every arm sees the same number of optimizer steps per model

    pooled = (torch.cat([c[0] for c in cohorts]), torch.cat([c[1] for c in cohorts]))
    central = train_centrally(pooled, budget, batch_size, lr)
    alone = [train_centrally(cohort, budget, batch_size, lr) for cohort in cohorts]
    federated, _ = train_federated(cohorts, rounds, local_steps, batch_size, lr, torch.device(device))
    federated = federated.to("cpu")

    return {
        "centralized (pooled data)": accuracy(central, test_x, test_y),
        "each hospital alone (mean)": sum(accuracy(m, test_x, test_y) for m in alone) / len(alone),
        "federated (FedAvg, data stays put)": accuracy(federated, test_x, test_y),
    }


def main() -> None:
    parser = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    parser.add_argument("--hospitals", type=int, default=4)
    parser.add_argument("--rounds", type=int, default=15)
    parser.add_argument("--device", default="cuda:0" if torch.cuda.is_available() else "cpu")
    args = parser.parse_args()
    for arm, value in walkthrough(args.hospitals, args.rounds, device=args.device).items():
        print(f"{arm:40s} accuracy on all hospitals' held-out patients: {value:.3f}")


if __name__ == "__main__":
    main()
