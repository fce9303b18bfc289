"""Per-kernel warp-stall breakdown (pc sampling) + instruction counts from an `ncu --set full` raw csv."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
ki = hdr.index("Kernel Name")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
extra = ["smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "sm__inst_executed_pipe_tc.sum", "smsp__warps_eligible.avg.per_cycle_active",
         "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_shared_mem",
         "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
         "sm__pipe_tensor_subunit_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.sum"]
for r in rows[2:]:
    print("==", r[ki].split("(")[0][-60:])
    st = []
    for i, h in stall_cols:
        try:
            st.append((float(r[i].replace(",", "")), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
        except ValueError:
            pass
    tot = sum(v for v, _ in st) or 1.0
    st.sort(reverse=True)
    print("   stalls:", ", ".join(f"{n} {100*v/tot:.0f}%" for v, n in st[:8]))
    for e in extra:
        if e in hdr:
            print("   ", e, r[hdr.index(e)])
