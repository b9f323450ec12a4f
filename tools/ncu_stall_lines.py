"""Source lines ranked by one stall reason of an ncu report (e.g. long_sb, wait, no_inst, math, short_sb).
usage: python tools/ncu_stall_lines.py report.ncu-rep REASON"""
import csv, subprocess, sys, collections
rep, which = sys.argv[1], sys.argv[2]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
# find header of the CUDA-source table
fname=None; hdr=None; agg=collections.Counter(); tot=0; allsamp=0
for r in rows:
    if not r: continue
    if r[0]=="File Path": fname=r[1].split("/")[-1]; continue
    if r[0]=="Line No": hdr=r; continue
    if hdr and r[0].isdigit() and fname:
        col={h:i for i,h in enumerate(hdr)}
        k=[h for h in hdr if h.startswith("stall_"+which) and "Not Issued" not in h]
        if not k: continue
        try:
            v=int(r[col[k[0]]] or 0); s=int(r[col["# Samples"]] or 0)
        except (ValueError, IndexError): continue
        agg[(fname,int(r[0]),r[1].strip()[:90])]+=v; tot+=v; allsamp+=s
print("total", which, tot, "of samples", allsamp)
for (f,l,t),v in agg.most_common(25):
    print(f"{f}:{l:5d} {100*v/max(1,tot):5.1f}% | {t}")
