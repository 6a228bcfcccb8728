#!/usr/bin/env python
"""Join the SASS page of an .ncu-rep (per-instruction stall samples / executed counts) with the line table of the
object that ran (nvdisasm -g), and print where the warp time goes per source line and per section of the kernel.

    python tools/ncu_hotspots.py REPORT.ncu-rep build/obj/pyh_march_nq1.o 'k_stage_marchILi0ELi0ELi0ELi1' [top]

The object must be the build that was profiled (same SASS); only the first captured launch is read."""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, obj, sym = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40

with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, check=True, capture_output=True)
    cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cubin)], capture_output=True, text=True).stdout

lines, cur, on = {}, ("?", 0), False
for ln in dis.splitlines():
    if ln.startswith(".text.") and ln.endswith(":"):
        on = sym in ln
        continue
    if not on:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        lines[int(m.group(1), 16)] = (cur, m.group(2).strip())

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ia, isamp, iex = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = None
per_line = collections.defaultdict(lambda: [0, 0, collections.Counter()])
tot_s = tot_e = 0
for r in rows[2:]:
    if len(r) <= iex or not r[ia].startswith("0x"):
        break
    a = int(r[ia], 16)
    if base is None:
        base = a
    key, _ = lines.get(a - base, (("?", 0), ""))
    s, e = int(r[isamp]), int(r[iex])
    per_line[key][0] += s
    per_line[key][1] += e
    for i, n in stall_cols:
        if r[i] not in ("", "0"):
            per_line[key][2][n] += int(r[i])
    tot_s += s
    tot_e += e
print(f"total samples {tot_s}, warp instructions {tot_e}")
print(f"{'file:line':34s} {'samples%':>8s} {'instr%':>7s} {'smp/instr':>9s}  top stalls")
for key, (s, e, st) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
    rel = (s / tot_s) / (e / tot_e) if e else 0.0
    tops = " ".join(f"{n}:{c * 100 // max(s, 1)}%" for n, c in st.most_common(3))
    print(f"{key[0] + ':' + str(key[1]):34s} {s / tot_s * 100:8.2f} {e / tot_e * 100:7.2f} {rel:9.2f}  {tops}")
byfile = collections.defaultdict(lambda: [0, 0])
for key, (s, e, st) in per_line.items():
    byfile[key[0]][0] += s
    byfile[key[0]][1] += e
print("per file:")
for f, (s, e) in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print(f"  {f:30s} samples {s / tot_s * 100:6.2f}%  instructions {e / tot_e * 100:6.2f}%")
