#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu -i ... --page raw --csv) into the few counters the roofline discussion needs.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [> profiles/xxx.md]"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
]
STALL = "smsp__average_warps_issue_stalled_"
OPS = ("dfma", "dmul", "dadd")


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print(f"## {r[col['Kernel Name']]}")
        for k in KEYS:
            if k in col:
                print(f"  {k}: {r[col[k]]} {units[col[k]]}")
        # FP64 thread-instruction counts: this ncu version exports them per elapsed cycle (summed over the SM
        # sub-partitions); times the elapsed cycles of one sub-partition = instructions of the launch
        cyc = num(r[col["smsp__cycles_elapsed.max"]]) if "smsp__cycles_elapsed.max" in col else None
        for op in OPS:
            k = f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed"
            if k in col and cyc:
                v = num(r[col[k]])
                if v is not None:
                    print(f"  thread instructions {op}: {v * cyc:.4g}  ({v:.1f} per cycle)")
        # stall reasons: warps stalled per issue-active cycle; printed as the share of all stalled-warp samples
        st = []
        for h, i in col.items():
            if h.startswith(STALL) and h.endswith("_per_issue_active.ratio"):
                v = num(r[i])
                if v is not None:
                    st.append((v, h[len(STALL):-len("_per_issue_active.ratio")]))
        tot = sum(v for v, _ in st) or 1.0
        st.sort(reverse=True)
        print("  stall reasons (share of warp-cycles, all reasons = 100 %): " + ", ".join(f"{n}={100.0 * v / tot:.1f}" for v, n in st[:9]))
        print()


if __name__ == "__main__":
    main()
