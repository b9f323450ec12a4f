"""Per-basic-block view of an ncu report: SASS in address order, grouped where the executed count changes.
usage: python tools/ncu_sass.py report.ncu-rep [--full]"""
import csv, subprocess, sys
rep = sys.argv[1]
full = "--full" in sys.argv
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
ins = {}
fname = None; cur = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] in ("Function Name", "Line No"): continue
    if r[0] != "":
        cur = (fname, r[0]); continue
    try:
        addr = int(r[2], 16)
    except ValueError:
        continue
    ins[addr] = (r[3].strip(), int(r[4] or 0), int(r[7] or 0), cur)
addrs = sorted(ins)
tot_s = sum(v[1] for v in ins.values()); tot_e = sum(v[2] for v in ins.values())
print(f"static instrs {len(addrs)} ({len(addrs)*16/1024:.0f} KB), samples {tot_s}, executed {tot_e}")
# group
blocks = []
for a in addrs:
    t, s, e, cur = ins[a]
    if blocks and blocks[-1]["e"] == e and not blocks[-1]["closed"]:
        b = blocks[-1]
    else:
        b = {"start": a, "e": e, "n": 0, "s": 0, "fp64": 0, "lds": 0, "ldg": 0, "lines": {}, "closed": False}
        blocks.append(b)
    b["n"] += 1; b["s"] += s
    op = t.split()[0] if t and not t.startswith("@") else (t.split()[1] if len(t.split()) > 1 else "")
    if op.startswith(("DFMA", "DMUL", "DADD", "DSETP", "DMMA")): b["fp64"] += 1
    if op.startswith(("LDS", "STS")): b["lds"] += 1
    if op.startswith(("LDG", "STG", "LD.", "ST.", "LDL", "STL")): b["ldg"] += 1
    b["lines"][cur] = b["lines"].get(cur, 0) + 1
    if op.startswith(("BRA", "RET", "CALL", "EXIT", "BSYNC")): b["closed"] = True
a0 = addrs[0]
for b in blocks:
    if b["e"] == 0 and not full: continue
    share = 100 * b["s"] / max(1, tot_s)
    if share < 0.3 and not full: continue
    top = sorted(b["lines"].items(), key=lambda kv: -kv[1])[:3]
    tops = " ".join(f"{k[0]}:{k[1]}x{v}" for k, v in top)
    print(f"+{(b['start']-a0)//16:6d} n={b['n']:5d} exec={b['e']:9d} dyn%={100*b['n']*b['e']/tot_e:5.1f} samp%={share:5.1f} cyc/inst={b['s']/max(1,b['n']*b['e'])*1e3:6.2f} fp64={b['fp64']:4d} sm={b['lds']:4d} gl={b['ldg']:4d} | {tops}")
