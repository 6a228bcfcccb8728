#!/usr/bin/env python
"""Dynamic opcode mix per section of the stage kernel, per cell (joins the SASS page of an .ncu-rep with the line table of
the object that ran; the line ranges below are those of the source at the commit that was profiled -- edit `secs`
when the files move).  Usage: python tools/ncu_section_mix.py REPORT.ncu-rep OBJECT.o SYMBOL_SUBSTRING CELLS_PER_LAUNCH"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, obj, sym = sys.argv[1], sys.argv[2], sys.argv[3]
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump","-xelf","all",os.path.abspath(obj)],cwd=td,check=True,capture_output=True)
    cubin=[f for f in os.listdir(td) if f.endswith(".cubin")][0]
    dis=subprocess.run(["nvdisasm","-g","-c",os.path.join(td,cubin)],capture_output=True,text=True).stdout
lines,cur,on={}, ("?",0), False
for ln in dis.splitlines():
    if ln.startswith(".text.") and ln.endswith(":"): on = sym in ln; continue
    if not on: continue
    m=re.match(r'\s*//## File "([^"]+)", line (\d+)',ln)
    if m: cur=(os.path.basename(m.group(1)),int(m.group(2))); continue
    m=re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);",ln)
    if m: lines[int(m.group(1),16)]=(cur,m.group(2).strip())
out=subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(out))); hdr=rows[1]
ia,iex=hdr.index("Address"),hdr.index("Instructions Executed")
base=None
F="pyh_fastdiv.cuh"; M="pyh_math.cuh"; S="pyh_stage_march.cuh"
secs=[("top/prologue",S,1,168),("B geom",S,169,195),("B pass1",S,196,216),("B pass2",S,217,262),("barrier..C setup",S,263,325),("C west",S,326,372),("C south",S,373,428),("D",S,429,480),
("seeds",F,24,39),("mid_range+scalar",F,40,113),("RangeAcc",F,115,146),("div4_fast",F,147,171),("recipN/divN_r/rcpN",F,172,215),("sqrtN",F,216,236),
("dmax/dmin",M,68,69),("rot",M,72,100),("limiter4_fast",M,156,209),("harten",M,259,269),("roe_face_fast",M,331,476),("fabs etc","math_functions.hpp",0,99999)]
mix={n:collections.Counter() for n,*_ in secs}; mix["other"]=collections.Counter()
tot=0
for r in rows[2:]:
    if len(r)<=iex or not r[ia].startswith("0x"): break
    a=int(r[ia],16)
    if base is None: base=a
    (f,l),ins=lines.get(a-base,(("?",0),"?"))
    ins=re.sub(r"^@!?U?P\d+\s+","",ins); op=ins.split()[0].split(".")[0]
    if op=="IMAD" and ".MOV" in ins.split()[0]: op="IMAD.MOV"
    n=int(r[iex]); tot+=n
    for name,ff,a0,b0 in secs:
        if f==ff and a0<=l<=b0: mix[name][op]+=n; break
    else: mix["other"][op]+=n
W = (int(sys.argv[4]) if len(sys.argv) > 4 else 33554432) / 32
for name,c in mix.items():
    t=sum(c.values())
    print(f"{name:22s} {t/W:7.1f}/cell  "+" ".join(f"{o}:{v/W:.0f}" for o,v in c.most_common(9)))
print("total per cell", tot/W)
