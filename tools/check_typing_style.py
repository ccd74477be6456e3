#!/usr/bin/env python
"""Pre-commit hook: reject pre-PEP-585/604 annotations (``typing.List``, ``Optional[X]``, ``Union[X, Y]`` …) in favour
of ``list[...]`` / ``X | None``.  Works on the AST, so strings and comments never trigger it.

    python tools/check_typing_style.py fl4health_b200/**/*.py
"""

from __future__ import annotations

import ast
import sys
from pathlib import Path

LEGACY = {"List", "Dict", "Tuple", "Set", "FrozenSet", "Type", "Optional", "Union", "Deque", "DefaultDict"}


def offences(path: Path) -> list[str]:
    tree = ast.parse(path.read_text(), filename=str(path))
    found = []
    for node in ast.walk(tree):
        if isinstance(node, ast.ImportFrom) and node.module == "typing":
            for alias in node.names:
                if alias.name in LEGACY:
                    found.append(f"{path}:{node.lineno}: `from typing import {alias.name}` (use the builtin / `|` form)")
        elif isinstance(node, ast.Attribute) and isinstance(node.value, ast.Name) and node.value.id == "typing" and node.attr in LEGACY:
            found.append(f"{path}:{node.lineno}: `typing.{node.attr}` (use the builtin / `|` form)")
    return found


def main(argv: list[str]) -> int:
    files = [Path(a) for a in argv] or sorted(Path("fl4health_b200").rglob("*.py"))
    problems = [line for f in files if f.suffix == ".py" for line in offences(f)]
    print("\n".join(problems))
    return 1 if problems else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
