#!/usr/bin/env python
"""Print headline metrics + hottest SASS lines of an ncu-rep (needs `ncu` on PATH; no GPU)."""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
minsamp = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
h, v = r[0], r[2]
want = ["Kernel Name", "Grid Size", "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed.avg.per_cycle_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "launch__registers_per_thread",
        "SM_A.TriageCompute.sm__inst_executed_pipe_xu_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warps_active.avg.per_cycle_active"]
for n in want:
    if n in h:
        print("%-90s %s %s" % (n, v[h.index(n)], r[1][h.index(n)]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hh, data = rows[1], rows[2:]
isrc, isamp, iex = hh.index("Source"), hh.index("# Samples"), hh.index("Instructions Executed")
stall = [i for i, n in enumerate(hh) if n.startswith("stall_") and "Not Issued" not in n]
tot = sum(int(x[isamp] or 0) for x in data)
texe = sum(int(x[iex] or 0) for x in data)
print("total samples", tot, "warp-instructions", texe)
agg = collections.Counter()
for x in data:
    for i in stall:
        agg[hh[i][6:]] += int(x[i] or 0)
print("stall mix:", [(k, c) for k, c in agg.most_common(8)])
for i, x in enumerate(data):
    s = int(x[isamp] or 0)
    if s >= minsamp:
        st = sorted([(int(x[j] or 0), hh[j][6:]) for j in stall], reverse=True)[:2]
        print(str(i).rjust(5), str(s).rjust(5), x[iex].rjust(8), x[isrc][:84].ljust(84), st)
