#!/usr/bin/env python
"""Summarise an .ncu-rep (details + stall reasons + dynamic opcode mix) -- used to write profiles/*.md"""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
def page(name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))

rows = page("raw")
hdr, units, r = rows[0], rows[1], rows[2]
def get(k):
    return r[hdr.index(k)] if k in hdr else "n/a"
print("kernel:", get("Kernel Name")[:90])
for k in ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
          "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
          "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__average_warp_latency_per_inst_issued.ratio",
          "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "smsp__inst_executed_op_branch.sum"]:
    if k in hdr:
        print(f"  {k:70s} {get(k):>18s} {units[hdr.index(k)]}")
print("stalls per issued instruction:")
st = []
for i, h in enumerate(hdr):
    m = re.match(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active.ratio", h)
    if m and r[i] not in ("", "0"):
        st.append((float(r[i]), m.group(1)))
for v, n in sorted(st, reverse=True)[:9]:
    print(f"  {n:28s} {v:6.3f}")
src = page("source")
h2 = src[1]
isrc, iex = h2.index("Source"), h2.index("Instructions Executed")
agg, tot = collections.Counter(), 0
for row in src[2:]:
    if len(row) <= iex: continue
    s = re.sub(r"^@!?U?P\d+\s+", "", row[isrc].strip())
    op = s.split()[0].split(".")[0] if s else "?"
    try:
        n = int(row[iex])
    except ValueError:
        break  # next kernel's header: only the first captured launch is summarised
    agg[op] += n; tot += n
print(f"dynamic opcode mix (warp instructions, total {tot}, static {len(src) - 2}):")
fp64 = sum(agg[o] for o in ("DFMA", "DMUL", "DADD", "DSETP"))
print(f"  FP64 (DFMA+DMUL+DADD+DSETP) {fp64 / tot * 100:5.1f}%")
print("  " + "  ".join(f"{o} {n / tot * 100:.1f}%" for o, n in agg.most_common(22)))
