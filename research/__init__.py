"""Experiment harnesses built on ``fl4health_b200`` (the counterpart of the reference's ``research/`` tree, which is not
part of its wheel either).  One harness (``research.harness``) replaces the reference's per-method ``server.py`` /
``client.py`` / ``*.slrm`` triplets: a *task* (data + model zoo) and a *method* (client class + server + strategy) are
combined by name, and sweeps run in-process on one GPU or SPMD under ``torchrun``."""
