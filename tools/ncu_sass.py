"""Per-basic-block view of an ncu report: SASS in address order, grouped where the executed count changes, with the
stall-reason mix of each block.
usage: python tools/ncu_sass.py report.ncu-rep [--full] [--dump START END]"""
import csv, subprocess, sys
rep = sys.argv[1]
full = "--full" in sys.argv
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
ins = []
for r in rows[2:]:
    try:
        int(r[0], 16)
    except (ValueError, IndexError):
        continue
    ins.append(r)
tot_s = sum(int(r[col["# Samples"]] or 0) for r in ins)
tot_e = sum(int(r[col["Instructions Executed"]] or 0) for r in ins)
print(f"static instrs {len(ins)} ({len(ins)*16/1024:.0f} KB), samples {tot_s}, executed {tot_e}")
tot_st = {h: sum(int(r[col[h]] or 0) for r in ins) for h in stall_cols}
print("stall mix: " + " ".join(f"{h[6:]}={100*v/max(1,tot_s):.1f}%" for h, v in sorted(tot_st.items(), key=lambda kv: -kv[1]) if v))
if "--dump" in sys.argv:
    i = sys.argv.index("--dump"); a, b = int(sys.argv[i + 1]), int(sys.argv[i + 2])
    for k in range(a, min(b, len(ins))):
        r = ins[k]
        st = {h[6:]: int(r[col[h]] or 0) for h in stall_cols if int(r[col[h]] or 0)}
        print(f"+{k:5d} {r[col['# Samples']]:>5s} {r[col['Instructions Executed']]:>8s} {r[1].strip():70s} {st}")
    sys.exit(0)
blocks = []
for k, r in enumerate(ins):
    t = r[1].strip(); s = int(r[col["# Samples"]] or 0); e = int(r[col["Instructions Executed"]] or 0)
    if blocks and blocks[-1]["e"] == e and not blocks[-1]["closed"]:
        b = blocks[-1]
    else:
        b = {"start": k, "e": e, "n": 0, "s": 0, "fp64": 0, "lds": 0, "ldg": 0, "loc": 0, "st": {}, "closed": False}
        blocks.append(b)
    b["n"] += 1; b["s"] += s
    toks = t.split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "")
    if op.startswith(("DFMA", "DMUL", "DADD", "DSETP", "DMMA")): b["fp64"] += 1
    if op.startswith(("LDS", "STS")): b["lds"] += 1
    if op.startswith(("LDG", "STG", "LD.", "ST.")): b["ldg"] += 1
    if op.startswith(("LDL", "STL")): b["loc"] += 1
    for h in stall_cols:
        v = int(r[col[h]] or 0)
        if v: b["st"][h[6:]] = b["st"].get(h[6:], 0) + v
    if op.startswith(("BRA", "RET", "CALL", "EXIT", "BSYNC")): b["closed"] = True
for b in blocks:
    if b["e"] == 0 and not full: continue
    share = 100 * b["s"] / max(1, tot_s)
    if share < 0.4 and not full: continue
    st = " ".join(f"{k}={100*v/max(1,b['s']):.0f}" for k, v in sorted(b["st"].items(), key=lambda kv: -kv[1])[:5])
    print(f"+{b['start']:6d} n={b['n']:5d} exec={b['e']:9d} dyn%={100*b['n']*b['e']/tot_e:5.1f} samp%={share:5.1f} fp64={b['fp64']:4d} sm={b['lds']:4d} gl={b['ldg']:4d} loc={b['loc']:3d} | {st}")
