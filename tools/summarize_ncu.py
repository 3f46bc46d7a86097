"""Condense `ncu --page raw --csv` exports into the few metrics the design notes quote."""
import csv, sys
KEYS = [
 ("gpu__time_duration.sum", "duration"),
 ("dram__bytes_read.sum", "dram_read"),
 ("dram__bytes_write.sum", "dram_write"),
 ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
 ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
 ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
 ("launch__registers_per_thread", "regs"),
 ("launch__grid_size", "grid"),
 ("launch__block_size", "block"),
 ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
 ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_cycles_pct"),
 ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
 ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
 ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor_hmma_pct"),
 ("l1tex__t_bytes.sum", "l1_bytes"),
 ("lts__t_bytes.sum", "l2_bytes"),
 ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
 ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
 ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall_lg_throttle"),
 ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
]
def main(path):
    rows = list(csv.reader(open(path)))
    # find header
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units, data = rows[h], rows[h + 1], rows[h + 2:]
    col = {n: i for i, n in enumerate(hdr)}
    tens = [n for n in hdr if "tensor" in n]
    for r in data:
        if len(r) < len(hdr): continue
        out = [r[col["Kernel Name"]][:60]]
        for k, short in KEYS:
            if k in col and r[col[k]] not in ("", "n/a"):
                out.append("%s=%s%s" % (short, r[col[k]], (" " + units[col[k]]) if units[col[k]] else ""))
        print(" | ".join(out))
    if "-t" in sys.argv:
        for r in data:
            print(r[col["Kernel Name"]][:40], {n: r[col[n]] for n in tens if r[col[n]] not in ("0", "", "n/a")})
for p in sys.argv[1:]:
    if p.startswith("-"): continue
    print("==", p); main(p)
