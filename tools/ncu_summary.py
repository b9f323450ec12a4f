"""Summarise an ncu report: headline metrics + executed warp-instructions per source line.
usage: python tools/ncu_summary.py report.ncu-rep FILTERS_PER_LAUNCH [top]"""
import csv
import subprocess
import sys

rep, nf = sys.argv[1], float(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, r = rows[0], rows[1], rows[2]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__occupancy_limit",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct", "sm__warps_active.avg.pct", "launch__shared_mem_per_block_dynamic",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct", "per_issue_active.ratio", "sm__inst_executed_pipe_tensor", "pipe_fp64_op_dmma", "sm__inst_executed_pipe_uniform"]
print("kernel:", r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "")
def val(name):
    return float(r[hdr.index(name)].replace(",", ""))
try:
    cyc = val("sm__cycles_elapsed.max")
    ops = {k: val(f"smsp__sass_thread_inst_executed_op_{k}_pred_on.sum.per_cycle_elapsed") * cyc for k in ("dadd", "dmul", "dfma")}
    flops = ops["dadd"] + ops["dmul"] + 2 * ops["dfma"]
    inst = ops["dadd"] + ops["dmul"] + ops["dfma"]
    ms = val("gpu__time_duration.sum") * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[hdr.index("gpu__time_duration.sum")].strip(), 1.0)
    print(f"  executed FP64 thread-instructions per filter: dadd {ops['dadd']/nf:.0f} dmul {ops['dmul']/nf:.0f} dfma {ops['dfma']/nf:.0f}"
          f"  = {inst/nf:.0f} instr, {flops/nf:.0f} flops (FMA = 2)")
    print(f"  executed FP64 rate: {flops/(ms*1e-3)/1e12:.2f} TFLOP/s; FP64 issue slots used: {100*inst/(cyc*148*64):.1f}% of 64 lanes/clk/SM")
except (ValueError, KeyError) as e:
    print("  (no op counts:", e, ")")
for h, u, v in zip(hdr, units, r):
    if any(k in h for k in keys) and "min" not in h and "max" not in h:
        try:
            if float(v.replace(",", "")) == 0: continue
        except ValueError:
            pass
        print(f"  {h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
agg, fname, seen_kernel = {}, None, 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": continue
    if r[0] != "":
        try:
            key = (fname, int(r[0]), r[1].strip()[:95])
            a = agg.setdefault(key, [0, 0, 0])
            a[0] += int(r[6] or 0); a[1] += int(r[7] or 0); a[2] += int(r[8] or 0)
        except (ValueError, IndexError):
            pass
tot = sum(a[1] for a in agg.values()); tots = sum(a[0] for a in agg.values())
print(f"total warp-instructions {tot}  = {tot / nf:.0f} per filter; samples {tots}")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{key[0]:15s}{key[1]:5d} {a[1] / nf:7.1f}/filt {100 * a[1] / tot:5.1f}% thr {a[2] / max(1, a[1]):5.1f} samp {100 * a[0] / tots:5.1f}% | {key[2]}")
