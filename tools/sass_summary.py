"""Per-kernel SASS evidence: which tensor-core / TMA / TMEM / multimem / cluster mnemonics each kernel of the in-tree
extension contains (``cuobjdump -sass``; runs without a GPU).

    python tools/sass_summary.py > profiles/r2/sass_mnemonics_by_kernel.txt
"""

from __future__ import annotations

import re
import subprocess
import sys
from collections import Counter
from pathlib import Path

LIB = Path(__file__).resolve().parents[1] / "fl4health_b200" / "ops" / "libfl4h_ops.so"
# mnemonic families that prove a hardware path is used (B200_PROFILING.md's list)
FAMILIES = {
    "tcgen05.mma": r"^UTC[HQIO]?MMA", "tmem ld/st": r"^(LDTM|STTM)", "tmem alloc": r"^UTCATOMSWS", "tcgen05.commit": r"^UTCBAR",
    "TMA load": r"^UTMALDG", "TMA store": r"^UTMASTG", "TMA reduce": r"^UTMAREDG", "mbarrier": r"^SYNCS",
    "multimem ld_reduce": r"^LDGMC", "sys-scope 128-bit store (peer st / multimem.st)": r"^STG\.E\.128\.STRONG\.SYS", "multimem red": r"^REDG?MC|^REDMC",
    "cluster barrier": r"^UCGABAR", "dsmem ld": r"^LDS\S*\.CLUSTER|^LD\.E\S*\.SHARED", "mapa": r"^MAPA", "PDL": r"^ACQBULK|^PREEXIT|^DEPBAR\.PRE",
    "legacy mma.sync": r"^(HMMA|IMMA|DMMA|QMMA)",
}


def main() -> None:
    text = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    kernel, per_kernel = None, {}
    for line in text.splitlines():
        found = re.search(r"Function : (\S+)", line)
        if found:
            kernel = subprocess.run(["c++filt", found.group(1)], capture_output=True, text=True).stdout.strip()
            per_kernel[kernel] = Counter()
            continue
        op = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", line)
        if op and kernel:
            per_kernel[kernel][op.group(1)] += 1
    print(f"# {LIB.name}: {len(per_kernel)} kernels; counts of instructions per mnemonic family (static SASS)\n")
    totals: Counter = Counter()
    for kernel, ops in sorted(per_kernel.items()):
        hits = {family: sum(count for op, count in ops.items() if re.match(pattern, op)) for family, pattern in FAMILIES.items()}
        hits = {family: count for family, count in hits.items() if count}
        totals.update(hits)
        if hits:
            short = re.sub(r"\((?!anonymous).*", "", kernel.replace("(anonymous namespace)::", ""))
            print(f"{short}\n    " + ", ".join(f"{family}={count}" for family, count in hits.items()))
    print("\n# totals\n" + "\n".join(f"{family}: {count}" for family, count in totals.items()))
    if "--ops" in sys.argv:  # full mnemonic histogram of the kernels named after the flag
        for want in sys.argv[sys.argv.index("--ops") + 1:]:
            for kernel, ops in per_kernel.items():
                if want in kernel:
                    print(f"\n## {kernel}\n" + "\n".join(f"{count:6d} {op}" for op, count in sorted(ops.items())))


if __name__ == "__main__":
    main()
