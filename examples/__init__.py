"""Runnable federated-learning scenarios (the counterpart of the reference's ``examples/`` tree).

``python -m examples.run <scenario> [--rounds N] [--clients K] [--device cuda] [--config path.yaml]`` builds the server,
the strategy and K clients of one scenario from ``examples/scenarios.py`` and runs them in-process
(``fl4health_b200.simulation``); ``torchrun ... -m examples.run <scenario> --spmd`` runs one client per GPU.
"""
