"""Executed code footprint of an ncu report: how many distinct instructions ran at least once, and per contiguous executed
region its share of the dynamic instructions, of the samples, and the no_instruction share of those samples.
usage: python tools/ncu_footprint.py report.ncu-rep"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; col = {h: i for i, h in enumerate(hdr)}
ins = []
for r in rows[2:]:
    try: int(r[0], 16)
    except (ValueError, IndexError): continue
    ins.append(r)
ex = [int(r[col["Instructions Executed"]] or 0) for r in ins]
sm = [int(r[col["# Samples"]] or 0) for r in ins]
ni = [int(r[col["stall_no_inst"]] or 0) if "stall_no_inst" in col else 0 for r in ins]
live = sum(1 for e in ex if e > 0)
print("static", len(ins), "executed-at-least-once", live, "=", live * 16 / 1024, "KB")
# contiguous executed regions
regs = []; start = None
for i, e in enumerate(ex + [0]):
    if e > 0 and start is None: start = i
    if e == 0 and start is not None:
        regs.append((start, i)); start = None
big = [(a, b) for a, b in regs if b - a > 50]
for a, b in big:
    print(f"  [{a:6d},{b:6d}) n={b-a:5d} dyn={sum(ex[a:b])/sum(ex)*100:5.1f}% samples={sum(sm[a:b])/sum(sm)*100:5.1f}% no_inst share of its samples={100*sum(ni[a:b])/max(1,sum(sm[a:b])):4.0f}%")
