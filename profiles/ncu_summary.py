"""Summarise an `ncu --set full` report (exported with --page raw --csv) into a per-kernel table."""
import csv
import sys

WANT = [
    ("time_us", "gpu__time_duration.sum"),
    ("dram_rd_MB", "dram__bytes_read.sum"),
    ("dram_wr_MB", "dram__bytes_write.sum"),
    ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("sm_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l1_pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l2_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("issue_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("fma_pipe_pct", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    ("alu_pipe_pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    ("lsu_pipe_pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("regs", "launch__registers_per_thread"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("smem_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    ("sm_mhz", "sm__cycles_elapsed.avg.per_second"),
]


def to_float(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    cols = [(n, hdr.index(m)) for n, m in WANT if m in hdr]
    print("| kernel | " + " | ".join(n for n, _ in cols) + " |")
    print("|---|" + "---|" * len(cols))
    for r in data:
        vals = []
        for n, i in cols:
            v, u = to_float(r[i]), units[i]
            if v is None:
                vals.append(r[i]); continue
            if n == "time_us":
                v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
            if n.endswith("_MB"):
                scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
                v = v * scale
            if n == "sm_mhz":
                v = v * {"hz": 1e-6, "Khz": 1e-3, "Mhz": 1.0, "Ghz": 1e3}.get(u, 1e-6)
            vals.append(f"{v:.1f}" if abs(v) < 1e6 else f"{v:.3g}")
        name = r[ki].split("(")[0].replace("void ", "").replace("b200::", "")
        print(f"| {name[:44]} | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
