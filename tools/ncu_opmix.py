"""Opcode mix of an ncu report, weighted by executed warp-instructions.
usage: python tools/ncu_opmix.py report.ncu-rep [warp_steps]   (warp_steps: warps x steps of the launch, default 32768)"""
import csv, subprocess, sys, collections, re
rep = sys.argv[1]
warp_steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1048576 / 32
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; col = {h: i for i, h in enumerate(hdr)}
cnt = collections.Counter(); tot = 0
for r in rows[2:]:
    try: int(r[0], 16)
    except (ValueError, IndexError): continue
    ex = int(r[col["Instructions Executed"]] or 0)
    src = r[1].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = m.group(2).split(".")[0] if m else src[:10]
    cnt[op] += ex; tot += ex
for op, c in cnt.most_common(28):
    print(f"{op:12s} {100*c/tot:6.2f}%  {c / warp_steps:8.1f} per warp-step")
