#!/usr/bin/env python
"""tools/make_traffic_json.py POSE_SUMMARY ORI_SUMMARY -- profiles/traffic.json from the ncu summaries
(tools/ncu_summary.py output of `ncu --set full` captures), stamped with the hash of the kernel sources the profiled
library was built from.  bench.py reports `roofline.traffic` / `roofline.executed` from this file only while the stamp
matches the sources of the library it runs."""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from slam_pose_estimation_b200 import _build  # noqa: E402


def parse(path):
    t = open(path).read()
    def num(pat, scale=1.0):
        m = re.search(pat, t)
        return float(m.group(1)) * scale if m else None
    unit = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3}
    def bytes_(name):
        m = re.search(name + r" \[(\w+)\] = ([0-9.]+)", t)
        return float(m.group(2)) * unit[m.group(1)]
    m = re.search(r"dadd (\d+) dmul (\d+) dfma (\d+)\s+= (\d+) instr, (\d+) flops", t)
    return {"kernel": re.search(r"kernel: void (.+?)\(", t).group(1), "dram_bytes_read_per_launch": int(bytes_(r"dram__bytes_read\.sum")),
            "dram_bytes_write_per_launch": int(bytes_(r"dram__bytes_write\.sum")), "fp64_instructions": int(m.group(4)),
            "fp64_flops": int(m.group(5)),
            "fp64_pipe_active_pct": num(r"sm__pipe_fp64_cycles_active\.avg\.pct_of_peak_sustained_active \[%\] = ([0-9.]+)"),
            "registers": int(num(r"launch__registers_per_thread \[register/thread\] = ([0-9.]+)"))}


pose, ori = parse(sys.argv[1]), parse(sys.argv[2])
out = {
    "source_sha16": _build.source_hash(),
    "kernel": pose["kernel"],
    "workload": "bench.py default: 1,048,576 PoseUKF filters, one fused predict + AngularVelocity update per launch, per-filter R",
    "dram_bytes_read_per_launch": pose["dram_bytes_read_per_launch"],
    "dram_bytes_write_per_launch": pose["dram_bytes_write_per_launch"],
    "dram_bytes_per_launch": pose["dram_bytes_read_per_launch"] + pose["dram_bytes_write_per_launch"],
    "algorithmic_bytes_per_launch": int(2616 * (1 << 20)),
    "executed_fp64_instructions_per_step": pose["fp64_instructions"],
    "executed_fp64_flops_per_step": pose["fp64_flops"],
    "fp64_pipe_active_pct": pose["fp64_pipe_active_pct"],
    "source": os.path.relpath(sys.argv[1], ROOT) + " (ncu --set full, one launch: dram__bytes_read.sum + dram__bytes_write.sum; "
              "smsp__sass_thread_inst_executed_op_{dadd,dmul,dfma}_pred_on)",
    "orientation_kernel": {
        "kernel": ori["kernel"],
        "workload": "tools/bench_c2.py: 1,048,576 OrientationUKF, 10 ticks per launch (10 predicts + 1 velocity update)",
        "executed_fp64_instructions_per_launch_per_filter": ori["fp64_instructions"],
        "executed_fp64_flops_per_launch_per_filter": ori["fp64_flops"],
        "fp64_pipe_active_pct": ori["fp64_pipe_active_pct"],
        "dram_bytes_per_launch": ori["dram_bytes_read_per_launch"] + ori["dram_bytes_write_per_launch"],
        "source": os.path.relpath(sys.argv[2], ROOT),
    },
}
old = {}
try:
    old = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
except Exception:
    pass
if "literal_kernel" in old:
    out["literal_kernel"] = old["literal_kernel"]  # ukf_thread.cuh: unchanged since its capture (round 1)
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=2)
print(json.dumps(out, indent=1)[:600])
