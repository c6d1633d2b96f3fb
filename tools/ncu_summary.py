#!/usr/bin/env python3
"""ncu_summary.py <report.ncu-rep> <out.json> [kernel-regex]  — the metrics DESIGN.md quotes, per profiled launch
(read here, on the GPU-less box, from a report captured under gpurun)."""
import csv
import io
import json
import re
import subprocess
import sys

KEYS = ("gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum",
        "sm__inst_executed.sum.per_cycle_elapsed", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed_op_local_st.sum", "smsp__inst_executed_op_local_ld.sum",
        "sm__sass_inst_executed_op_local_st.sum", "sm__sass_inst_executed_op_local_ld.sum")


def main():
    rep, out = sys.argv[1], sys.argv[2]
    pat = re.compile(sys.argv[3]) if len(sys.argv) > 3 else None
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d.get("Kernel Name", "")
        if pat and not pat.search(name):
            continue
        rec = {"kernel": name}
        for i, h in enumerate(hdr):
            if h in KEYS or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                if h.startswith("smsp__average_warps_issue_stalled") and v < 0.05:
                    continue
                rec[h + (f" [{units[i]}]" if units[i] else "")] = v
        res.append(rec)
    json.dump({"report": rep, "launches": res}, open(out, "w"), indent=1)
    print(f"{len(res)} launch(es) -> {out}")


if __name__ == "__main__":
    main()
