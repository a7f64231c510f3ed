#!/usr/bin/env python
"""Key figures of an `ncu --page raw --csv` export (one line per metric of interest), and the hottest source lines of a
`--page source --csv` export.  python tools/ncu_summary.py raw.csv [source.csv]"""
import csv
import sys

KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "launch__occupancy_limit_registers",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "lts__t_sector_op_write_hit_rate.pct",
        "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fmaheavy.sum", "sm__inst_executed_pipe_uniform.sum",
        "smsp__pcsamp_sample_buffer_full", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    for k in KEYS:
        for i, h in enumerate(hdr):
            if h == k or h.endswith("." + k):
                print(f"{k:<88} {units[i]:>14}  " + "  ".join(r[i] for r in rows[2:5]))
                break
    if len(sys.argv) > 2:
        src = list(csv.reader(open(sys.argv[2])))
        h = src[0]
        ci = {n: i for i, n in enumerate(h)}
        print("columns:", [n for n in h][:40])


def hot_lines(path, top=25):
    src = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(src) if r and r[0] == "Address")
    h = src[hi]
    ci = {n: i for i, n in enumerate(h)}
    rows = [r for r in src[hi + 1:] if len(r) == len(h)]
    tot = sum(int(r[ci["# Samples"]] or 0) for r in rows)
    inst = sum(int(r[ci["Instructions Executed"]] or 0) for r in rows)
    print(f"total samples {tot}, warp instructions executed {inst}, static instructions {len(rows)}")
    stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    agg = {n: sum(int(r[ci[n]] or 0) for r in rows) for n in stall_cols}
    print("stall samples:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
    rows.sort(key=lambda r: -int(r[ci["# Samples"]] or 0))
    for r in rows[:top]:
        st = {n[6:]: int(r[ci[n]] or 0) for n in stall_cols if int(r[ci[n]] or 0)}
        st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print(f"{r[ci['# Samples']]:>7} {r[ci['Instructions Executed']]:>10}  {r[ci['Source']][:70]:<70} {st}")


if __name__ == "__main__":
    main()
    if len(sys.argv) > 2:
        hot_lines(sys.argv[2])
