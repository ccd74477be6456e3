"""Fused peer-memory collectives on >=2 GPUs (skipped on single-GPU boxes)."""

import json
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]


def _torchrun(script: str, args: list[str], nproc: int, port: int, timeout: int = 600,
              env: dict | None = None) -> subprocess.CompletedProcess:
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), str(ROOT / "tests" / script), *args]
    return subprocess.run(cmd, env=dict(os.environ, FL4H_LOG_LEVEL="WARNING", **(env or {})), capture_output=True, text=True,
                          timeout=timeout, cwd=ROOT)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("nvls", ["1", "0"])
def test_fused_collectives_match_nccl(tmp_path: Path, nvls: str) -> None:
    """NVLS (multimem) and fixed-order P2P data paths against an NCCL all-gather + single-GPU reduction."""
    nproc = min(torch.cuda.device_count(), 8)
    out = tmp_path / "fused.json"
    proc = _torchrun("fused_worker.py", [str(out), str(1 << 22)], nproc, 29731 + int(nvls), env={"FL4H_NVLS": nvls})
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]
    report = json.loads(out.read_text())
    if nvls == "0":
        assert not report["nvls"] and report["agg_bit_exact"], report
    else:
        assert report["nvls"], report  # B200 + NVSwitch: the multicast path must be the one that ran
        assert report["agg_max_abs_err"] < 1e-5, report  # in-switch reduction order differs from the fixed order
    assert report["agg_uniform_max_abs_err"] < 1e-5, report
    assert report["int_ok"], report
    assert report["adam_max_abs_err"] < 1e-4, report
    assert report["bcast_ok"], report


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_spmd_fedavg_fused_equals_nccl(tmp_path: Path) -> None:
    results = {}
    for mode, port in (("nccl", 29741), ("fused", 29742)):
        out = tmp_path / f"{mode}.json"
        env_args = [str(out), "fedavg"]
        cmd_env = dict(os.environ, FL4H_COLLECTIVES=mode, FL4H_LOG_LEVEL="WARNING")
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
               "127.0.0.1", "--master-port", str(port), str(ROOT / "tests" / "spmd_worker.py"), *env_args]
        proc = subprocess.run(cmd, env=cmd_env, capture_output=True, text=True, timeout=600, cwd=ROOT)
        assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]
        results[mode] = json.loads(out.read_text())
    for key, value in results["nccl"]["state"].items():
        assert abs(results["fused"]["state"][key] - value) < 1e-3 * max(1.0, abs(value)), key
