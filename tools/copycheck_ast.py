#!/usr/bin/env python
"""Self-check for derivative code: AST-normalised statement overlap between each file of ``fl4health_b200/`` and the
same-named file of the reference tree (docstrings stripped, imports dropped, ``ast.unparse`` formatting).

    python tools/copycheck_ast.py [--ref /root/reference/fl4health] [--min 0.35]

Prints ``fraction_of_repo_statements_matched  n_statements  path`` sorted by fraction; exit code 1 when any file not on
the allow-list (schema / constant / ABC files whose shape the preserved public API dictates) exceeds ``--min``."""

from __future__ import annotations

import argparse
import ast
import difflib
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

ALLOW = {  # similarity dictated by the preserved API: constants, type aliases, ABCs, tiny managers, schema classes
    "feature_alignment/constants.py", "feature_alignment/tabular_feature.py", "feature_alignment/tabular_type.py",
    "utils/typing.py", "metrics/base_metrics.py", "parameter_exchange/partial_parameter_exchanger.py",
    "parameter_exchange/parameter_exchanger_base.py",
    "client_managers/fixed_sampling_client_manager.py", "client_managers/base_sampling_manager.py",
    "servers/adaptive_constraint_servers/fedprox_server.py", "servers/adaptive_constraint_servers/ditto_server.py",
    "servers/adaptive_constraint_servers/mrmtl_server.py", "model_bases/sequential_split_models.py",
    "model_bases/parallel_split_models.py", "model_bases/fedsimclr_base.py",
    "mixins/base.py",
    # <= 20 statements, every one a constructor argument / attribute / one-line delegation fixed by the public API
    "parameter_exchange/packing_exchanger.py", "parameter_exchange/fedpm_exchanger.py", "losses/cosine_similarity_loss.py",
    "losses/perfcl_loss.py", "model_bases/moon_base.py", "servers/fedpm_server.py", "reporting/reports_manager.py",
    "preprocessing/pca_preprocessor.py", "clients/instance_level_dp_client.py", "clients/partial_weight_exchange_client.py",
    # containers whose constructors must keep the reference's attribute names (configs and clients read them); the
    # computation behind them is restructured (one weighted-term / one latent-code code path)
    "losses/fenda_loss_config.py", "preprocessing/autoencoders/dim_reduction.py", "model_bases/autoencoders_base.py",
}


def _strip(tree: ast.AST) -> ast.AST:
    for node in ast.walk(tree):
        if isinstance(node, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef, ast.Module)):
            body = node.body
            if body and isinstance(body[0], ast.Expr) and isinstance(getattr(body[0], "value", None), ast.Constant) \
                    and isinstance(body[0].value.value, str):
                node.body = body[1:] or [ast.Pass()]
    return tree


def statements(path: Path) -> list[str]:
    try:
        tree = _strip(ast.parse(path.read_text()))
    except SyntaxError:
        return []
    lines = []
    for line in ast.unparse(tree).splitlines():
        t = line.strip()
        if not t or t == "pass" or t.startswith(("import ", "from ")):
            continue
        lines.append(t)
    return lines


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference/fl4health")
    ap.add_argument("--min", type=float, default=0.35)
    ap.add_argument("--all", action="store_true", help="print every compared pair")
    args = ap.parse_args()
    ref_root, repo_root = Path(args.ref), ROOT / "fl4health_b200"
    rows = []
    for path in sorted(repo_root.rglob("*.py")):
        rel = path.relative_to(repo_root)
        ref = ref_root / rel
        if not ref.exists():
            continue
        mine, theirs = statements(path), statements(ref)
        if len(mine) < 8:
            continue
        matcher = difflib.SequenceMatcher(None, mine, theirs, autojunk=False)
        matched = sum(block.size for block in matcher.get_matching_blocks())
        rows.append((matched / len(mine), len(mine), str(rel)))
    rows.sort(reverse=True)
    bad = 0
    for frac, n, rel in rows:
        flagged = frac > args.min and rel not in ALLOW
        bad += flagged
        if args.all or frac > args.min:
            print(f"{frac:5.2f} {n:5d} {rel}{'' if flagged else '   (allow-listed)' if frac > args.min else ''}")
    total = sum(n for _, n, _ in rows)
    weighted = sum(f * n for f, n, _ in rows) / max(total, 1)
    print(f"# {len(rows)} pairs, weighted overlap {weighted:.2f}, {bad} file(s) over {args.min}")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
