#!/usr/bin/env python
"""tools/build_variant.py NAME [nvcc flags ...] -- compiles the engine with extra flags (usually -D tuning knobs of the
kernels) into slam_pose_estimation_b200/lib/variants/NAME.so; tools/bench_libs.sh / UKFB_LIB select such a build."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from slam_pose_estimation_b200 import _build  # noqa: E402

name, extra = sys.argv[1], sys.argv[2:]
out = os.path.join(ROOT, "slam_pose_estimation_b200", "lib", "variants", name + ".so")
os.makedirs(os.path.dirname(out), exist_ok=True)
cmd = [_build.nvcc(), *_build.NVCC_FLAGS, *extra, "-o", out, *[os.path.join(_build.CSRC, s) for s in _build.SOURCES]]
subprocess.run(cmd, check=True, cwd=_build.CSRC)
r = subprocess.run(["cuobjdump", "-res-usage", out], capture_output=True, text=True).stdout.splitlines()
for i, line in enumerate(r):
    if "fast_kernelILb0" in line:
        print(name, line.split("_ZN4ukfb")[-1][:24], r[i + 1].strip()[:40])
