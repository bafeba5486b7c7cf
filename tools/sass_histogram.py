"""Regenerates profiles/r2_sass_opcode_histogram.txt: `cuobjdump -sass keypoint_learning_b200/libkpl_b200.so`, one opcode histogram per kernel."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "keypoint_learning_b200", "libkpl_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, hist = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); hist[cur] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        hist[cur][m.group(1).split(".")[0]] += 1


def demangle(n):
    try:
        return re.sub(r"\(.*", "", subprocess.check_output(["c++filt", n], text=True).strip())[:110]
    except Exception:
        return n


out = ["# SASS opcode histogram of every kernel in keypoint_learning_b200/libkpl_b200.so (cuobjdump -sass, sm_100a cubins only;",
       "# static instruction counts, not execution counts).  Packed FP32 = FADD2/FMUL2/FFMA2; no UTMA*/UBLKCP (bulk copies), no UTC*MMA /",
       "# tcgen05 (tensor cores): nothing on this path is a dense contraction (DESIGN.md s5).  Regenerate: python tools/sass_histogram.py", ""]
tot = collections.Counter()
for k, c in hist.items():
    if not c:
        continue
    n = sum(c.values())
    out.append("%-110s total %5d  packed-f32 %4d  LDS %4d STS %4d LDG %4d MUFU %3d F2I/I2F %3d" % (
        demangle(k), n, c["FADD2"] + c["FMUL2"] + c["FFMA2"], c["LDS"], c["STS"], c["LDG"], c["MUFU"], c["F2I"] + c["I2F"] + c["I2FP"]))
    out.append("    " + ", ".join("%s %d" % (o, v) for o, v in c.most_common(14)))
    tot.update(c)
out += ["", "ALL KERNELS: " + ", ".join("%s %d" % (o, v) for o, v in tot.most_common(40))]
bad = [o for o in tot if o.startswith(("UTMA", "UBLKCP", "UTC", "HMMA", "IMMA", "QGMMA"))]
out.append("tensor-core / bulk-copy opcodes present: %s" % (bad or "none"))
dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_sass_opcode_histogram.txt")
open(dst, "w").write("\n".join(out) + "\n")
