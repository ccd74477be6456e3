"""Whole federations vs the reference: the same clients (model, data, optimizer, seeds) and the same server set-up, run
(1) by the UNMODIFIED reference -- its ``FlServer`` / strategy / client classes over the stand-alone Flower transport on
localhost -- and (2) by this framework's in-process simulation.  The per-round histories (aggregated validation loss,
aggregated fit and evaluation metrics) must coincide."""
import importlib
import socket
import sys
import threading
from pathlib import Path

import torch
from torch import nn
from torch.utils.data import DataLoader

ROUNDS, CLIENTS, LOCAL_STEPS, BATCH = 3, 3, 4, 16
agreed = 0
# The reference's clients run as threads of this process and share the global generator with everything else (every
# DataLoader iterator draws a base seed from it), so models are NOT initialised from it: after construction every
# randomly initialised layer is refilled from a private generator.
def pin_initialisation(module: nn.Module, seed: int = 7) -> nn.Module:
    generator = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for layer in module.modules():
            if isinstance(layer, (nn.Linear, nn.Embedding, nn.Conv1d, nn.Conv2d)):
                for parameter in layer.parameters(recurse=False):
                    parameter.copy_(torch.randn(parameter.shape, generator=generator) * (0.25 if parameter.dim() > 1 else 0.05))
    return module


def cohort(index: int) -> tuple[torch.Tensor, torch.Tensor]:
    generator = torch.Generator().manual_seed(100 + index)
    features = torch.randn(96, 10, generator=generator) + 0.4 * index
    labels = (features[:, :3].sum(dim=1) + 0.3 * torch.randn(96, generator=generator) > 0.4 * index * 3).long()
    return features, labels


class Net(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.body = nn.Sequential(nn.Linear(10, 16), nn.ReLU())
        self.head = nn.Linear(16, 2)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.head(self.body(x))


def user_hooks(side, index: int, model_factory=Net, lr: float = 0.05) -> dict:
    """The four user hooks, identical on both sides (``side`` resolves the package the dataset class comes from)."""
    dataset_module = side("utils.dataset")

    def get_model(self, config):
        return pin_initialisation(model_factory()).to(self.device)  # every client starts from the same initialisation

    def get_data_loaders(self, config):
        features, labels = cohort(index)
        train = dataset_module.TensorDataset(features[:64], labels[:64])
        val = dataset_module.TensorDataset(features[64:], labels[64:])
        # private generators: a DataLoader iterator draws its base seed from the generator it is given, else the global one
        return (DataLoader(train, batch_size=BATCH, shuffle=False, generator=torch.Generator().manual_seed(1)),
                DataLoader(val, batch_size=BATCH, shuffle=False, generator=torch.Generator().manual_seed(2)))

    def get_optimizer(self, config):
        return torch.optim.SGD(self.model.parameters(), lr=lr, momentum=0.9)

    def get_criterion(self, config):
        return nn.CrossEntropyLoss()

    return {"get_model": get_model, "get_data_loaders": get_data_loaders, "get_optimizer": get_optimizer, "get_criterion": get_criterion}


def resolver(prefix: str):
    return lambda dotted: importlib.import_module(f"{prefix}.{dotted}")


def round_config(extra: dict):
    def config_fn(server_round: int) -> dict:
        config = {"current_server_round": server_round, "local_steps": LOCAL_STEPS, "batch_size": BATCH, "n_server_rounds": ROUNDS, **extra}
        if "local_epochs" in extra or "local_head_steps" in extra or "local_head_epochs" in extra:
            del config["local_steps"]  # mutually exclusive ways of saying how long to train
        return config

    return config_fn


def build(side, scenario: dict, ours: bool):
    """Server and clients of one scenario from the package ``side`` resolves."""
    client_cls = getattr(side(scenario["client"][0]), scenario["client"][1])
    accuracy = side("metrics").Accuracy if hasattr(side("metrics"), "Accuracy") else side("metrics.metrics").Accuracy
    clients = []
    for index in range(CLIENTS):
        hooks = user_hooks(side, index, **scenario.get("hook_args", {}))
        hooks.update(scenario.get("extra_hooks", lambda side, index: {})(side, index))
        cls = type(f"Client{index}", (client_cls,), hooks)
        if scenario.get("personalize"):  # dynamically personalised flexible clients
            personalized = side("mixins.personalized")
            cls = personalized.make_it_personal(cls, getattr(personalized.PersonalizedMode, scenario["personalize"]))
        clients.append(cls(data_path=Path("."), metrics=[accuracy()], device=torch.device("cpu"), client_name=f"client_{index}",
                           **scenario.get("client_args", lambda side: {})(side)))
    aggregation = side("metrics.metric_aggregation")
    config_fn = round_config(scenario.get("config", {}))
    strategy_cls = getattr(side(scenario["strategy"][0]), scenario["strategy"][1])
    strategy_args = dict(on_fit_config_fn=config_fn, on_evaluate_config_fn=config_fn,
                         fit_metrics_aggregation_fn=aggregation.fit_metrics_aggregation_fn,
                         evaluate_metrics_aggregation_fn=aggregation.evaluate_metrics_aggregation_fn, min_available_clients=CLIENTS)
    if scenario.get("min_fit", True):
        strategy_args.update(min_fit_clients=CLIENTS, min_evaluate_clients=CLIENTS)
    strategy_args.update(scenario.get("strategy_args", lambda side, ours: {})(side, ours))
    strategy = strategy_cls(**strategy_args)
    server_cls = getattr(side(scenario["server"][0]), scenario["server"][1])
    manager = scenario.get("manager", lambda side, ours: (side("servers.client_manager") if ours else importlib.import_module("flwr.server.client_manager")).SimpleClientManager())(side, ours)
    server = server_cls(client_manager=manager, fl_config={"n_server_rounds": ROUNDS}, strategy=strategy,
                        on_init_parameters_config_fn=config_fn, accept_failures=scenario.get("accept_failures", False),
                        **scenario.get("server_args", lambda side: {})(side))
    return server, clients


def run_reference(scenario: dict):
    import flwr

    server, clients = build(resolver("fl4health"), scenario, ours=False)
    with socket.socket() as probe:
        probe.bind(("127.0.0.1", 0))
        port = probe.getsockname()[1]
    address = f"127.0.0.1:{port}"
    threads = [threading.Thread(target=flwr.client.start_client, kwargs=dict(server_address=address, client=client.to_client(), cid=client.client_name), daemon=True)
               for client in clients]
    for thread in threads:
        thread.start()
    history = flwr.server.start_server(server=server, server_address=address, config=flwr.server.ServerConfig(num_rounds=ROUNDS))
    for thread in threads:
        thread.join(60)
    server.shutdown()  # flushes the reporters, as the reference's example servers do after ``start_server`` returns
    return history


def run_ours(scenario: dict):
    from fl4health_b200.simulation import run_simulation

    server, clients = build(resolver("fl4health_b200"), scenario, ours=True)
    return run_simulation(server, clients, num_rounds=ROUNDS)


def compare(name: str, theirs, ours, tol: float) -> None:
    global agreed
    assert [r for r, _ in theirs.losses_distributed] == [r for r, _ in ours.losses_distributed], name
    for (server_round, a), (_, b) in zip(theirs.losses_distributed, ours.losses_distributed):
        assert abs(a - b) <= tol * max(1.0, abs(a)), (name, "loss", server_round, a, b)
    central_theirs, central_ours = getattr(theirs, "losses_centralized", []), getattr(ours, "losses_centralized", [])
    assert len(central_theirs) == len(central_ours), (name, central_theirs, central_ours)
    for (round_a, a), (round_b, b) in zip(central_theirs, central_ours):
        assert round_a == round_b and abs(a - b) <= tol * max(1.0, abs(a)), (name, "central loss", central_theirs, central_ours)
    for label, left, right in (("fit", theirs.metrics_distributed_fit, ours.metrics_distributed_fit), ("eval", theirs.metrics_distributed, ours.metrics_distributed),
                               ("central", getattr(theirs, "metrics_centralized", {}), getattr(ours, "metrics_centralized", {}))):
        assert set(left) == set(right), (name, label, sorted(left), sorted(right))
        for key in left:
            for (server_round, a), (_, b) in zip(left[key], right[key]):
                assert abs(float(a) - float(b)) <= tol * max(1.0, abs(float(a))), (name, label, key, server_round, a, b)
    print(f"  {name}: {len(theirs.losses_distributed)} rounds, {len(theirs.metrics_distributed_fit)} fit / {len(theirs.metrics_distributed)} eval metric series agree", file=sys.stderr)
    agreed += 1


def optional_hooks(side, index):
    """A test loader, a step-wise learning-rate schedule and a bounded validation pass."""
    dataset_module = side("utils.dataset")

    def get_test_data_loader(self, config):
        features, labels = cohort(index + 10)
        return DataLoader(dataset_module.TensorDataset(features[:40], labels[:40]), batch_size=BATCH, shuffle=False)

    def get_lr_scheduler(self, optimizer_key, config):
        return torch.optim.lr_scheduler.StepLR(self.optimizers[optimizer_key], step_size=3, gamma=0.5)

    return {"get_test_data_loader": get_test_data_loader, "get_lr_scheduler": get_lr_scheduler}


# (Early stopping is not compared: in this reference version ``EarlyStopper.should_stop`` calls the client's validation
# with a logging mode its own assertion rejects -- ``basic_client.py:844`` -- so the reference arm cannot run it.)


def failing_client_hooks(side, index):
    """Client 2 raises during its round-2 fit; the others carry on."""
    base = side("clients.basic_client").BasicClient

    def fit(self, parameters, config):
        if index == 2 and config["current_server_round"] == 2:
            raise RuntimeError("simulated client failure")
        return base.fit(self, parameters, config)

    return {"fit": fit}


def central_evaluation(side, ours):
    """Server-side evaluation of the aggregate on held-out data after every round (and of the initial model)."""
    features, labels = cohort(9)

    def evaluate_fn(server_round, arrays, config):
        model = Net()
        tensors = [torch.as_tensor(a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a) for a in arrays]
        model.load_state_dict(dict(zip(model.state_dict(), tensors)))
        with torch.no_grad():
            logits = model(features)
        return float(nn.functional.cross_entropy(logits, labels)), {"central - accuracy": float((logits.argmax(1) == labels).float().mean())}

    return {**initial_parameters()(side, ours), "evaluate_fn": evaluate_fn}


def seeded(factory):
    return pin_initialisation(factory())


def initial_parameters(factory=Net):
    """The server-side initial model: the same seeded initialisation the clients build."""
    return lambda side, ours: {"initial_parameters": side("utils.parameter_extraction").get_all_model_parameters(seeded(factory))}


def adaptive_constraint(factory=Net, **kwargs):
    settings = dict(initial_loss_weight=0.1, adapt_loss_weight=True, loss_weight_delta=0.05, loss_weight_patience=1)
    settings.update(kwargs)
    return lambda side, ours: {**initial_parameters(factory)(side, ours), **settings}


def two_optimizers(first: str, second: str, first_of, second_of):
    def hooks(side, index):
        def get_optimizer(self, config):
            return {first: torch.optim.SGD(first_of(self).parameters(), lr=0.05, momentum=0.9),
                    second: torch.optim.SGD(second_of(self).parameters(), lr=0.05, momentum=0.9)}

        return {"get_optimizer": get_optimizer}

    return hooks


def model_hook(build_model):
    """Replace ``get_model`` by a seeded factory that needs classes from the side's own package."""
    def hooks(side, index):
        def get_model(self, config):
            return pin_initialisation(build_model(side)).to(self.device)

        return {"get_model": get_model}

    return hooks


def merged(*hook_makers):
    def hooks(side, index):
        out = {}
        for make in hook_makers:
            out.update(make(side, index))
        return out

    return hooks


class NormNet(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.first, self.norm, self.head = nn.Linear(10, 16), nn.BatchNorm1d(16), nn.Linear(16, 2)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.head(torch.relu(self.norm(self.first(x))))


class Body(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.layers = nn.Sequential(nn.Linear(10, 16), nn.ReLU())

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.layers(x)


def parallel_head(side):
    module = side("model_bases.parallel_split_models")

    class Head(module.ParallelSplitHeadModule):
        def __init__(self) -> None:
            super().__init__(module.ParallelFeatureJoinMode.CONCATENATE)
            self.classifier = nn.Linear(32, 2)

        def parallel_output_join(self, local_tensor, global_tensor):
            return torch.cat([local_tensor, global_tensor], dim=1)

        def head_forward(self, input_tensor):
            return self.classifier(input_tensor)

    return Head()


FEDAVG = dict(strategy=("strategies.basic_fedavg", "BasicFedAvg"), server=("servers.base_server", "FlServer"))
SCENARIOS = {
    "fedavg": dict(client=("clients.basic_client", "BasicClient"), **FEDAVG),
    "fedavg_epochs": dict(client=("clients.basic_client", "BasicClient"), **FEDAVG, config={"local_epochs": 2}),
    "fedprox": dict(client=("clients.fed_prox_client", "FedProxClient"), strategy=("strategies.fedavg_with_adaptive_constraint", "FedAvgWithAdaptiveConstraint"),
                    server=("servers.adaptive_constraint_servers.fedprox_server", "FedProxServer"), strategy_args=adaptive_constraint()),
    "ditto": dict(client=("clients.ditto_client", "DittoClient"), strategy=("strategies.fedavg_with_adaptive_constraint", "FedAvgWithAdaptiveConstraint"),
                  server=("servers.adaptive_constraint_servers.ditto_server", "DittoServer"), strategy_args=adaptive_constraint(initial_loss_weight=0.5),
                  extra_hooks=two_optimizers("global", "local", lambda c: c.global_model, lambda c: c.model)),
    "mr_mtl": dict(client=("clients.mr_mtl_client", "MrMtlClient"), strategy=("strategies.fedavg_with_adaptive_constraint", "FedAvgWithAdaptiveConstraint"),
                   server=("servers.adaptive_constraint_servers.mrmtl_server", "MrMtlServer"), strategy_args=adaptive_constraint(adapt_loss_weight=False)),
    "scaffold": dict(client=("clients.scaffold_client", "ScaffoldClient"), strategy=("strategies.scaffold", "Scaffold"), server=("servers.scaffold_server", "ScaffoldServer"),
                     min_fit=False, hook_args={"lr": 0.05}, server_args=lambda side: {"warm_start": True},
                     manager=lambda side, ours: side("client_managers.fixed_without_replacement_manager").FixedSamplingByFractionClientManager(),
                     strategy_args=lambda side, ours: {**initial_parameters()(side, ours), "model": seeded(Net), "learning_rate": 1.0}),
    "apfl": dict(client=("clients.apfl_client", "ApflClient"), **FEDAVG,
                 extra_hooks=merged(model_hook(lambda side: side("model_bases.apfl_base").ApflModule(Net(), alpha_lr=0.05)),
                                    two_optimizers("local", "global", lambda c: c.model.local_model, lambda c: c.model.global_model))),
    "moon": dict(client=("clients.moon_client", "MoonClient"), **FEDAVG, client_args=lambda side: {"contrastive_weight": 2.0},
                 extra_hooks=model_hook(lambda side: side("model_bases.moon_base").MoonModel(Body(), nn.Linear(8, 2), nn.Linear(16, 8)))),
    "fedper": dict(client=("clients.fedper_client", "FedPerClient"), **FEDAVG,
                   extra_hooks=model_hook(lambda side: side("model_bases.sequential_split_models").SequentiallySplitExchangeBaseModel(Body(), nn.Linear(16, 2)))),
    "fenda": dict(client=("clients.fenda_client", "FendaClient"), **FEDAVG,
                  extra_hooks=model_hook(lambda side: side("model_bases.fenda_base").FendaModel(Body(), Body(), parallel_head(side)))),
    "fedrep": dict(client=("clients.fedrep_client", "FedRepClient"), **FEDAVG, config={"local_head_steps": 2, "local_rep_steps": 3},
                   extra_hooks=merged(model_hook(lambda side: side("model_bases.fedrep_base").FedRepModel(Body(), nn.Linear(16, 2))),
                                      two_optimizers("representation", "head", lambda c: c.model.base_module, lambda c: c.model.head_module))),
    "fedbn": dict(client=("clients.fedbn_client", "FedBnClient"), **FEDAVG, hook_args={"model_factory": lambda: NormNet()},
                  extra_hooks=lambda side, index: {"get_parameter_exchanger": lambda self, config: side("parameter_exchange.layer_exchanger").LayerExchangerWithExclusions(self.model, {nn.BatchNorm1d})}),
    "perfcl": dict(client=("clients.perfcl_client", "PerFclClient"), **FEDAVG,
                   client_args=lambda side: {"global_feature_contrastive_loss_weight": 0.5, "local_feature_contrastive_loss_weight": 2.0},
                   extra_hooks=model_hook(lambda side: side("model_bases.perfcl_base").PerFclModel(Body(), Body(), parallel_head(side)))),
    "gpfl": dict(client=("clients.gpfl_client", "GpflClient"), **FEDAVG, client_args=lambda side: {"lam": 0.05, "mu": 0.02},
                 extra_hooks=merged(model_hook(lambda side: side("model_bases.gpfl_base").GpflModel(Body(), nn.Linear(16, 2), feature_dim=16, num_classes=2)),
                                    lambda side, index: {"get_optimizer": lambda self, config: {
                                        "model": torch.optim.SGD(self.model.gpfl_main_module.parameters(), lr=0.05),
                                        "gce": torch.optim.SGD(self.model.gce.embedding.parameters(), lr=0.05, weight_decay=0.02),
                                        "cov": torch.optim.SGD(self.model.cov.parameters(), lr=0.05, weight_decay=0.02)}})),
    "ensemble": dict(client=("clients.ensemble_client", "EnsembleClient"), **FEDAVG,
                     extra_hooks=merged(model_hook(lambda side: side("model_bases.ensemble_base").EnsembleModel({"model_0": Net(), "model_1": Net()})),
                                        lambda side, index: {"get_optimizer": lambda self, config: {
                                            name: torch.optim.SGD(member.parameters(), lr=0.05, momentum=0.9) for name, member in self.model.ensemble_models.items()}})),
    "dynamic_layers": dict(client=("clients.partial_weight_exchange_client", "PartialWeightExchangeClient"), strategy=("strategies.fedavg_dynamic_layer", "FedAvgDynamicLayer"),
                           server=("servers.base_server", "FlServer"), strategy_args=initial_parameters(), client_args=lambda side: {"store_initial_model": True},
                           extra_hooks=lambda side, index: {"get_parameter_exchanger": lambda self, config: side("parameter_exchange.layer_exchanger").DynamicLayerExchanger(
                               side("parameter_exchange.parameter_selection_criteria").LayerSelectionFunctionConstructor(0.01, 0.5, True, True).select_by_percentage())}),
    "sparse_coo": dict(client=("clients.partial_weight_exchange_client", "PartialWeightExchangeClient"), strategy=("strategies.fedavg_sparse_coo_tensor", "FedAvgSparseCooTensor"),
                       server=("servers.base_server", "FlServer"), strategy_args=initial_parameters(), client_args=lambda side: {"store_initial_model": True},
                       extra_hooks=lambda side, index: {"get_parameter_exchanger": lambda self, config: side("parameter_exchange.sparse_coo_parameter_exchanger").SparseCooParameterExchanger(
                           0.3, side("parameter_exchange.parameter_selection_criteria").largest_final_magnitude_scores)}),
    "feddg_ga": dict(client=("clients.basic_client", "BasicClient"), strategy=("strategies.feddg_ga", "FedDgGa"), server=("servers.base_server", "FlServer"),
                     strategy_args=initial_parameters(), config={"evaluate_after_fit": True, "pack_losses_with_val_metrics": True},
                     manager=lambda side, ours: side("client_managers.fixed_sampling_client_manager").FixedSamplingClientManager()),
    "fenda_ditto": dict(client=("clients.fenda_ditto_client", "FendaDittoClient"), strategy=("strategies.fedavg_with_adaptive_constraint", "FedAvgWithAdaptiveConstraint"),
                        server=("servers.base_server", "FlServer"),
                        strategy_args=lambda side, ours: {**adaptive_constraint(initial_loss_weight=0.5)(side, ours), "initial_parameters": side("utils.parameter_extraction").get_all_model_parameters(
                            seeded(lambda: side("model_bases.sequential_split_models").SequentiallySplitExchangeBaseModel(Body(), nn.Linear(16, 2))))},
                        extra_hooks=merged(model_hook(lambda side: side("model_bases.fenda_base").FendaModel(Body(), Body(), parallel_head(side))),
                                           lambda side, index: {"get_global_model": lambda self, config: seeded(
                                               lambda: side("model_bases.sequential_split_models").SequentiallySplitModel(Body(), nn.Linear(16, 2))).to(self.device)},
                                           two_optimizers("global", "local", lambda c: c.global_model, lambda c: c.model))),
    "constrained_fenda": dict(client=("clients.constrained_fenda_client", "ConstrainedFendaClient"), **FEDAVG,
                              client_args=lambda side: {"loss_container": side("losses.fenda_loss_config").ConstrainedFendaLossContainer(
                                  None, side("losses.fenda_loss_config").CosineSimilarityLossContainer(torch.device("cpu"), 0.5),
                                  side("losses.fenda_loss_config").MoonContrastiveLossContainer(torch.device("cpu"), 1.5))},
                              extra_hooks=model_hook(lambda side: side("model_bases.fenda_base").FendaModelWithFeatureState(Body(), Body(), parallel_head(side), flatten_features=True))),
    "constrained_fenda_perfcl": dict(client=("clients.constrained_fenda_client", "ConstrainedFendaClient"), **FEDAVG,
                                     client_args=lambda side: {"loss_container": side("losses.fenda_loss_config").ConstrainedFendaLossContainer(
                                         side("losses.fenda_loss_config").PerFclLossContainer(torch.device("cpu"), 0.7, 1.3), None, None)},
                                     extra_hooks=model_hook(lambda side: side("model_bases.fenda_base").FendaModelWithFeatureState(Body(), Body(), parallel_head(side), flatten_features=True))),
    "flexible": dict(client=("clients.flexible.base", "FlexibleClient"), **FEDAVG),
    "flexible_ditto": dict(client=("clients.flexible.base", "FlexibleClient"), personalize="DITTO", strategy=("strategies.fedavg_with_adaptive_constraint", "FedAvgWithAdaptiveConstraint"),
                           server=("servers.adaptive_constraint_servers.ditto_server", "DittoServer"), strategy_args=adaptive_constraint(initial_loss_weight=0.5),
                           extra_hooks=lambda side, index: {"get_optimizer": lambda self, config: {"local": torch.optim.SGD(self.model.parameters(), lr=0.05, momentum=0.9)}}),
    "flexible_mr_mtl": dict(client=("clients.flexible.base", "FlexibleClient"), personalize="MR_MTL", strategy=("strategies.fedavg_with_adaptive_constraint", "FedAvgWithAdaptiveConstraint"),
                            server=("servers.adaptive_constraint_servers.mrmtl_server", "MrMtlServer"), strategy_args=adaptive_constraint(adapt_loss_weight=False)),
    "options": dict(client=("clients.basic_client", "BasicClient"), **FEDAVG, extra_hooks=optional_hooks, config={"num_validation_steps": 1}),
    # the same algorithms under other round configurations
    "fedavg_unweighted_eval_after_fit": dict(client=("clients.basic_client", "BasicClient"), **FEDAVG, config={"evaluate_after_fit": True, "pack_losses_with_val_metrics": True},
                                             strategy_args=lambda side, ours: {"weighted_aggregation": False, "weighted_eval_losses": False}),
    "fedprox_epochs_unweighted": dict(client=("clients.fed_prox_client", "FedProxClient"), strategy=("strategies.fedavg_with_adaptive_constraint", "FedAvgWithAdaptiveConstraint"),
                                      server=("servers.adaptive_constraint_servers.fedprox_server", "FedProxServer"), config={"local_epochs": 1},
                                      strategy_args=adaptive_constraint(weighted_aggregation=False, weighted_train_losses=False, loss_weight_patience=2)),
    "ditto_epochs": dict(client=("clients.ditto_client", "DittoClient"), strategy=("strategies.fedavg_with_adaptive_constraint", "FedAvgWithAdaptiveConstraint"),
                         server=("servers.adaptive_constraint_servers.ditto_server", "DittoServer"), strategy_args=adaptive_constraint(initial_loss_weight=1.0, adapt_loss_weight=False),
                         config={"local_epochs": 2}, extra_hooks=two_optimizers("global", "local", lambda c: c.global_model, lambda c: c.model)),
    "apfl_epochs": dict(client=("clients.apfl_client", "ApflClient"), **FEDAVG, config={"local_epochs": 1},
                        extra_hooks=merged(model_hook(lambda side: side("model_bases.apfl_base").ApflModule(Net(), adaptive_alpha=False, alpha=0.25)),
                                           two_optimizers("local", "global", lambda c: c.model.local_model, lambda c: c.model.global_model))),
    "scaffold_cold_start_epochs": dict(client=("clients.scaffold_client", "ScaffoldClient"), strategy=("strategies.scaffold", "Scaffold"), server=("servers.scaffold_server", "ScaffoldServer"),
                                       min_fit=False, hook_args={"lr": 0.05}, config={"local_epochs": 1},
                                       manager=lambda side, ours: side("client_managers.fixed_without_replacement_manager").FixedSamplingByFractionClientManager(),
                                       strategy_args=lambda side, ours: {**initial_parameters()(side, ours), "model": seeded(Net), "learning_rate": 0.5}),
    "fedrep_epochs": dict(client=("clients.fedrep_client", "FedRepClient"), **FEDAVG, config={"local_head_epochs": 1, "local_rep_epochs": 1},
                          extra_hooks=merged(model_hook(lambda side: side("model_bases.fedrep_base").FedRepModel(Body(), nn.Linear(16, 2))),
                                             two_optimizers("representation", "head", lambda c: c.model.base_module, lambda c: c.model.head_module))),
    # constructor options of the personalised clients
    "moon_two_old_models": dict(client=("clients.moon_client", "MoonClient"), **FEDAVG, client_args=lambda side: {"contrastive_weight": 1.0, "temperature": 0.2, "len_old_models_buffer": 2},
                                extra_hooks=model_hook(lambda side: side("model_bases.moon_base").MoonModel(Body(), nn.Linear(16, 2)))),
    "fenda_ditto_frozen_extractor": dict(client=("clients.fenda_ditto_client", "FendaDittoClient"), strategy=("strategies.fedavg_with_adaptive_constraint", "FedAvgWithAdaptiveConstraint"),
                                         server=("servers.base_server", "FlServer"), client_args=lambda side: {"freeze_global_feature_extractor": True},
                                         strategy_args=lambda side, ours: {**adaptive_constraint(initial_loss_weight=0.5)(side, ours), "initial_parameters": side("utils.parameter_extraction").get_all_model_parameters(
                                             seeded(lambda: side("model_bases.sequential_split_models").SequentiallySplitExchangeBaseModel(Body(), nn.Linear(16, 2))))},
                                         extra_hooks=merged(model_hook(lambda side: side("model_bases.fenda_base").FendaModel(Body(), Body(), parallel_head(side))),
                                                            lambda side, index: {"get_global_model": lambda self, config: seeded(
                                                                lambda: side("model_bases.sequential_split_models").SequentiallySplitModel(Body(), nn.Linear(16, 2))).to(self.device)},
                                                            two_optimizers("global", "local", lambda c: c.global_model, lambda c: c.model))),
    "perfcl_temperatures": dict(client=("clients.perfcl_client", "PerFclClient"), **FEDAVG,
                                client_args=lambda side: {"global_feature_loss_temperature": 0.1, "local_feature_loss_temperature": 1.5,
                                                          "global_feature_contrastive_loss_weight": 1.0, "local_feature_contrastive_loss_weight": 0.25},
                                extra_hooks=model_hook(lambda side: side("model_bases.perfcl_base").PerFclModel(Body(), Body(), parallel_head(side)))),
    # server behaviour: a client that fails mid-run (tolerated), centralised evaluation next to the federated one
    "client_failure_tolerated": dict(client=("clients.basic_client", "BasicClient"), **FEDAVG, extra_hooks=failing_client_hooks, accept_failures=True,
                                     # (sample sizes stay 3 -- Flower sizes a sample from the clients connected at that instant -- and the round goes on
                                     # with the two results that arrive)
                                     strategy_args=lambda side, ours: {"accept_failures": True}),
    "central_evaluation": dict(client=("clients.basic_client", "BasicClient"), **FEDAVG, strategy_args=central_evaluation),
    "flash": dict(client=("clients.flash_client", "FlashClient"), strategy=("strategies.flash", "Flash"), server=("servers.base_server", "FlServer"),
                  strategy_args=lambda side, ours: {**initial_parameters()(side, ours), "eta": 0.1, "eta_l": 0.05}, config={"local_epochs": 1, "gamma": 0.5}),
}

if __name__ == "__main__":
    wanted = sys.argv[1:] or list(SCENARIOS)
    for name in wanted:
        compare(name, run_reference(SCENARIOS[name]), run_ours(SCENARIOS[name]), tol=2e-4)
    print("configs agree:", agreed)
