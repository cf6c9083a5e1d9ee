#!/usr/bin/env python
"""Summarise an ncu report offline: key raw metrics and the source-page hot spots of one kernel.
    python tools/ncu_hot.py report.ncu-rep kernel_regex [top_n]"""
import csv, subprocess, sys, io, collections
rep, pat = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name", "regex:" + pat], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")])
    for k in keys:
        if k in hdr:
            print("   %-75s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and "per_issue_active" in h and float(r[i] or 0) > 0.15:
            print("   %-75s %s" % (h.replace("smsp__average_warps_issue_stalled_", "stall ").replace("_per_issue_active.ratio", ""), r[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
seen, data = set(), []
for r in rows[2:]:
    try:
        a = int(r[ia], 16)
    except Exception:
        continue
    if a in seen:
        continue
    seen.add(a)
    data.append((a, r[isrc].strip(), int(r[isamp]), int(r[iex])))
base = data[0][0]
tot = sum(d[2] for d in data)
print("samples", tot, "warp instructions", sum(d[3] for d in data))
win = collections.OrderedDict()
for a, s, n, e in data:
    w = win.setdefault((a - base) // 0x400, [0, 0]); w[0] += n; w[1] += e
print("-- 1 KB code windows with > 1.5 % of the samples")
for k, (n, e) in win.items():
    if n > tot * 0.015:
        print("   0x%05x  %5.1f %%  exec %d" % (k * 0x400, 100.0 * n / tot, e))
print("-- top instructions")
for a, s, n, e in sorted(data, key=lambda d: -d[2])[:topn]:
    print("   0x%05x %5.2f %% exec %10d  %s" % (a - base, 100.0 * n / tot, e, s[:100]))
