"""Build the profiles/ summary of `ncu --set full` captures: one row per captured unit, selected metrics.
usage: python tools/ncu_full_table.py out.csv "label 1" raw1.csv ["label 2" raw2.csv ...]   (raw = `ncu -i rep --page raw --csv`)"""
import csv
import sys

WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__cluster_size", "lts__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "sm__cycles_active.avg"]

out, args = sys.argv[1], sys.argv[2:]
table, units, cols = [], None, None
for label, path in zip(args[0::2], args[1::2]):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    if cols is None:
        cols = [w for w in WANT if w in idx]
        units = [rows[1][idx[c]] for c in cols]
    for r in rows[2:]:
        table.append([label] + [r[idx[c]] if c in idx else "" for c in cols])
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["unit (batch 64 x 249 frames)"] + cols)
    w.writerow([""] + units)
    w.writerows(table)
print(f"{len(table)} rows -> {out}")
