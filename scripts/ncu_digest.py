#!/usr/bin/env python
"""Digest of an .ncu-rep (read here, on the CPU): per kernel the numbers the profiles/ summaries quote.
usage: python scripts/ncu_digest.py file.ncu-rep [more metric-name substrings]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
        ("lts__t_bytes.sum", "L2 bytes"), ("l1tex__t_bytes.sum", "L1 bytes"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
        ("launch__shared_mem_per_block_dynamic", "dyn smem"), ("launch__occupancy_limit_registers", "occ limit regs"),
        ("launch__occupancy_limit_shared_mem", "occ limit smem"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occ %"),
        ("smsp__inst_executed.sum", "warp insts"), ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue busy %"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
        ("l1tex__data_pipe_lsu_wavefronts.sum", "L1 wavefronts"), ("l1tex__data_bank_conflicts_pipe_lsu.sum", "smem bank conflicts"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_sb"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_throttle"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
        ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall membar"),
        ("smsp__average_warps_issue_stalled_drain_per_issue_active.ratio", "stall drain")]
for extra in sys.argv[2:]:
    want += [(h, h) for h in hdr if extra in h]
col = {h: q for q, h in enumerate(hdr)}
kn = col.get("Kernel Name")
for r in data:
    print("====", r[kn][:90])
    seen = set()
    for m, label in want:
        if m in col and label not in seen and r[col[m]] != "":
            seen.add(label)
            print("  %-22s %s %s" % (label, r[col[m]], units[col[m]]))
