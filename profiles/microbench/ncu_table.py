"""Markdown table of selected counters from an `ncu --set full` report (read with `ncu -i REP --page raw --csv`), one
column per captured launch.  `python profiles/microbench/ncu_table.py REP.ncu-rep [substring ...]` — extra arguments add
metric-name substrings to the default selection."""
import csv
import subprocess
import sys

DEFAULT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
    "sm__cycles_active.avg", "gpc__cycles_elapsed.max", "gpc__cycles_elapsed.avg.per_second",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
]


def main():
    rep, extra = sys.argv[1], sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head, units, data = rows[0], rows[1], rows[2:]
    ki = head.index("Kernel Name")
    names = []
    for r in data:
        n = r[ki].replace("unnamed>::", "").replace("void ", "")
        names.append(n.split("(")[0][:48])
    print("| metric | " + " | ".join(f"`{n}`" for n in names) + " |")
    print("|---|" + "---|" * len(names))
    for i, m in enumerate(head):
        if m in DEFAULT or any(e in m for e in extra):
            vals = []
            for r in data:
                v = r[i]
                try:
                    f = float(v)
                    v = f"{f:.6g}"
                except ValueError:
                    pass
                vals.append(v)
            unit = f" [{units[i]}]" if units[i] else ""
            print(f"| {m}{unit} | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main()
