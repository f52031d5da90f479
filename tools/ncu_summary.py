#!/usr/bin/env python3
"""Print the roofline-relevant raw metrics of every kernel in an .ncu-rep as one compact line each."""
import csv, subprocess, sys
WANT = [("gpu__time_duration.sum", "dur"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("launch__registers_per_thread", "regs"), ("launch__occupancy_limit_registers", "occ_reg"),
        ("launch__occupancy_limit_shared_mem", "occ_smem"), ("launch__grid_size", "grid"),
        ("smsp__inst_executed.sum", "winst"), ("sm__cycles_elapsed.max", "cyc"), ("sm__cycles_active.avg", "cyc_act"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wf"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_confl"),
        ("l1tex__data_pipe_lsu_wavefronts.sum", "lsu_wf"),
        ("lts__t_bytes.sum", "l2_bytes"),
        ("smsp__average_warp_latency_issue_stalled_barrier.pct", "st_bar"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_barrier"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "st_notsel"),
        ]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")][:60]
    parts = []
    for m, short in WANT:
        if m in hdr:
            i = hdr.index(m)
            v = r[i]
            try:
                v = f"{float(v.replace(',', '')):.4g}"
            except ValueError:
                pass
            parts.append(f"{short}={v}{units[i] if units[i] not in ('', '%') else ''}")
    print(name, "|", " ".join(parts))
