"""Print selected metrics from `ncu -i X.ncu-rep --page raw --csv` (path of the csv as argv[1])."""
import sys, csv
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct", "dram__throughput.avg.pct",
        "sm__inst_executed_pipe_tensor", "sm__pipe_tensor", "launch__grid_size", "launch__registers_per_thread", "launch__block_size", "sm__warps_active.avg.pct",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared", "smsp__average_warps_issue_stalled", "smsp__average_warp_latency_issue_stalled",
        "smsp__issue_active.avg.pct", "smsp__inst_executed.sum"] + sys.argv[2:]
for r in rows[2:]:
    for i, h in enumerate(hdr):
        if any(h.startswith(k) for k in keys):
            print(f"{h:90s} {rows[1][i]:14s} {r[i]}")
    print("-" * 40)
