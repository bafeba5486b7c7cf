"""GPU-box check of the C++ host path over several GPUs: writes a closed-surface scene as a binary PCD, runs
TestDetector with --gpus 1 and --gpus N (one host thread per rank, NCCL inside libkpl_b200.so) and compares the
keypoint files.   python tools/cli_multi_gpu_demo.py N [points]"""
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from keypoint_learning_b200 import synth  # noqa: E402

gpus = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
xyz, _ = synth.scene_closed_surfaces(n, seed=4321)
tmp = os.environ.get("TMPDIR", "/tmp")
pcd = os.path.join(tmp, "scene.pcd")
with open(pcd, "wb") as f:
    f.write(("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH %d\nHEIGHT 1\n"
             "VIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA binary\n" % (n, n)).encode())
    f.write(np.ascontiguousarray(xyz, np.float32).tobytes())
td = os.path.join(ROOT, "keypoint_learning_b200", "TestDetector")
forest = os.path.join(ROOT, "tests", "golden", "forests", "synthetic-T100-D15.yaml.gz")
out = {}
for g in (1, gpus):
    kp = os.path.join(tmp, "kp%d.pcd" % g)
    t0 = time.perf_counter()
    r = subprocess.run([td, "--pathCloud", pcd, "--pathRF", forest, "--pathKP", kp, "--stats", "--gpus", str(g)], capture_output=True, text=True, timeout=900)
    dt = time.perf_counter() - t0
    print("=== TestDetector --gpus %d: rc %d, %.1f s wall (PCD load + detection + keypoint file)" % (g, r.returncode, dt))
    print("\n".join(l for l in r.stdout.splitlines() if l.startswith(("points", "rank", "device"))))
    if r.returncode:
        print(r.stderr[-2000:])
        sys.exit(1)
    out[g] = open(kp).read().splitlines()[11:]
print("keypoints: %d vs %d, files identical: %s" % (len(out[1]), len(out[gpus]), out[1] == out[gpus]))
sys.exit(0 if out[1] == out[gpus] else 2)
