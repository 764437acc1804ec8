"""Key metrics of an `ncu --set full` report (ncu -i X.ncu-rep --page raw --csv > X.csv; python summarize_ncu.py X.csv)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
WANT = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__sass_inst_executed_op_tmem_ldt.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_active.avg", "sm__cycles_elapsed.max"]
stall = [h for h in hdr if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h]
for r in rows[2:]:
    print("----")
    for w in WANT:
        if w in idx:
            print(f"{w:72s} {r[idx[w]]} {units[idx[w]]}")
    tot = sum(float(r[idx[h]]) for h in stall) or 1.0
    top = sorted(stall, key=lambda h: -float(r[idx[h]]))[:8]
    print("stall samples: " + ", ".join(f"{h.replace('smsp__pcsamp_warps_issue_stalled_', '')} {float(r[idx[h]]) / tot * 100:.1f}%" for h in top))
