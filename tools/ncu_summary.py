"""Text summary of an `ncu --set full` report for profiles/: python tools/ncu_summary.py <file.ncu-rep> "<title>" [top_n_stalls]"""
import csv
import subprocess
import sys

rep, title = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 12
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, unit, vals = rows[0], rows[1], rows[-1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__block_size", "launch__cluster_size",
        "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.avg.per_second", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum"]
print(f"# {title}")
print("# (cold-cache, serialised replay under ncu: durations here are NOT bench values)")
for w in want:
    for h, u, v in zip(hdr, unit, vals):
        if h == w:
            print(f"{h} = {v} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(src.splitlines()))
if len(srows) > 2:
    sh = srows[1]
    idx = {h: i for i, h in enumerate(sh)}
    data = [r for r in srows[2:] if len(r) == len(sh)]
    tot = sum(int(r[idx["# Samples"]]) for r in data) or 1
    stall_cols = [h for h in sh if h.startswith("stall_") and "Not Issued" not in h]
    agg = {h: sum(int(r[idx[h]] or 0) for r in data) for h in stall_cols}
    print("# warp-state samples by reason (all warps, % of samples):")
    print("  " + "  ".join(f"{k[6:]} {100 * v / tot:.1f}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    print(f"# top {top_n} SASS instructions by samples:")
    for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]]))[:top_n]:
        st = sorted(((h[6:], int(r[idx[h]] or 0)) for h in stall_cols), key=lambda kv: -kv[1])[:2]
        print(f"  {100 * int(r[idx['# Samples']]) / tot:5.1f}%  {r[idx['Source']].strip()[:60]:60s} {st}")
