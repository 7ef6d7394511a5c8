"""Condense an `ncu --page raw --csv` export into the per-launch metrics the docs quote.
    python scripts/ncu_summary.py gpurun_out/prof_r02_raw.csv > profiles/ncu_full_r02_summary.txt"""
import csv
import sys

KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.max", "smsp__inst_executed_op_global_atom.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_atom.sum", "lts__t_sectors_op_atom.sum"]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units = rows[hdr], rows[hdr + 1]
    col = {n: i for i, n in enumerate(names)}
    for r in rows[hdr + 2:]:
        if len(r) < len(names):
            continue
        kname = r[col["Kernel Name"]]
        short = kname.split("(")[0].replace("<unnamed>::", "")
        print("== launch %s: %s   grid %s block %s" % (r[col["ID"]], short, r[col.get("Grid Size", 0)], r[col.get("Block Size", 0)]))
        for k in KEEP:
            if k in col:
                print("   %-70s %s %s" % (k, r[col[k]], units[col[k]]))
        stalls = []
        for n, i in col.items():
            if n.startswith(STALL) and n.endswith("_per_warp_active.pct") is False and n.endswith(".ratio"):
                try:
                    stalls.append((float(r[i]), n[len(STALL):].replace("_per_warp_active.ratio", "").replace(".ratio", "")))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        if stalls:
            print("   stall cycles per issued instruction: " + ", ".join("%s %.2f" % (n, v) for v, n in stalls[:8]))
        print()


if __name__ == "__main__":
    main()
