"""Where does a kernel spill?  usage: python tools/spill_lines.py KERNEL_SUBSTRING [top]
Disassembles lib/libukfb.so with line info (cuobjdump -xelf + nvdisasm -g) and counts STL / LDL per source line."""
import collections
import glob
import os
import re
import subprocess
import sys
import tempfile

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.environ.get("UKFB_LIB") or os.path.join(root, "slam_pose_estimation_b200", "lib", "libukfb.so")
name, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, stdout=subprocess.DEVNULL)
    dis = subprocess.run(["nvdisasm", "-g", "-c", glob.glob(d + "/*.cubin")[0]], capture_output=True, text=True).stdout
inside, cur, cnt, total = False, None, collections.Counter(), 0
for l in dis.splitlines():
    if l.startswith("\t.section\t.text."):
        inside = name in l
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
    elif re.search(r"\b(STL|LDL)\b", l):
        cnt[(cur, "STL" if "STL" in l else "LDL")] += 1
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        total += 1
print(f"{name}: {total} instructions, {sum(cnt.values())} local loads/stores")
for (k, t), v in cnt.most_common(top):
    src = ""
    try:
        src = open(os.path.join(root, "slam_pose_estimation_b200", "csrc", k[0])).read().splitlines()[k[1] - 1].strip()[:90]
    except Exception:
        pass
    print(f"{v:4d} {t} {k[0]}:{k[1]}  {src}")
