"""Filter ``ncu -i report.ncu-rep --page raw --csv`` (stdin) down to the columns the profiling recipe asks for
(/opt/skills/guides/B200_PROFILING.md): kernel, duration, DRAM bytes / throughput, tensor-pipe activity, SM throughput,
achieved occupancy, registers.  Writes CSV to stdout (header row, units row, one row per captured launch)."""

from __future__ import annotations

import csv
import re
import sys

WANTED = re.compile(
    r"^(ID|Kernel Name|Block Size|Grid Size|gpu__time_duration\.sum|dram__bytes_(read|write)\.sum|dram__cycles_active\.avg\.pct.*|"
    r"gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|sm__pipe_tensor.*cycles_active\.avg\.pct_of_peak_sustained_active|"
    r"sm__inst_executed_pipe_tensor.*pct.*|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"sm__warps_active\.avg\.pct_of_peak_sustained_active|launch__registers_per_thread|launch__occupancy_limit.*|"
    r"l1tex__m_xbar2l1tex_read_bytes\.sum|lts__t_bytes\.sum|smsp__cycles_active\.avg|launch__shared_mem_per_block.*)$")


def main() -> None:
    rows = list(csv.reader(sys.stdin))
    rows = [r for r in rows if r]
    header_index = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    header = rows[header_index]
    keep = [i for i, name in enumerate(header) if WANTED.match(name)]
    writer = csv.writer(sys.stdout)
    for row in rows[header_index:]:
        writer.writerow([row[i] if i < len(row) else "" for i in keep])


if __name__ == "__main__":
    main()
