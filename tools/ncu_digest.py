"""Digest an .ncu-rep (read here, no GPU): headline metrics + instructions/stalls by source line.
   python tools/ncu_digest.py gpurun_out/x.ncu-rep [top_n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]
for h, u, v in zip(hdr, units, vals):
    if h in want or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
        try:
            if h.startswith("smsp__average_warps") and float(v) < 0.3:
                continue
        except ValueError:
            pass
        print("%-90s %-14s %s" % (h, u, v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = None
cur = None
out = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) > 2 and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0] != "":
        out.append((int(r[7] or 0), int(r[8] or 0), cur, r[0], r[1].strip()[:95], int(r[6] or 0)))
tot = sum(o[0] for o in out) or 1
ts = sum(o[5] for o in out) or 1
print("total warp-instructions", tot, "stall samples", ts)
print("--- by instructions")
for o in sorted(out, reverse=True)[:top]:
    print("%5.1f%% inst  thr/inst %4.1f  %4.1f%% smp  %s:%s  %s" % (100 * o[0] / tot, o[1] / max(o[0], 1), 100 * o[5] / ts, o[2], o[3], o[4]))
print("--- by stall samples")
for o in sorted(out, key=lambda x: -x[5])[:12]:
    print("%5.1f%% smp  %4.1f%% inst  thr/inst %4.1f  %s:%s  %s" % (100 * o[5] / ts, 100 * o[0] / tot, o[1] / max(o[0], 1), o[2], o[3], o[4]))
