"""GPU-box experiment: where the host-buffer entry point spends its time beyond the device work (10 M-point scene)."""
import ctypes as C, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import keypoint_learning_b200 as K
from keypoint_learning_b200 import synth
import bench
xyz, vp = synth.scene_closed_surfaces(10_000_000, seed=4321)
n = len(xyz)
det = bench.make_detector(K, 0, vp, 4)
x4 = np.ones((n, 4), np.float32); x4[:, :3] = xyz
hx = torch.from_numpy(x4).pin_memory().numpy()
sc = torch.empty(n, dtype=torch.float32).pin_memory().numpy()
kp = torch.empty(n, dtype=torch.int32).pin_memory().numpy()
det.setInputCloud(hx)
for _ in range(2):
    det.compute(scores_out=sc, kp_out=kp)
L = det._L
f32p, i32p = C.POINTER(C.c_float), C.POINTER(C.c_int32)
nk = C.c_int64(0)
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rc = L.kpl_detect(det._h, hx.ctypes.data_as(f32p), 16, None, 0, None, n, sc.ctypes.data_as(f32p), kp.ctypes.data_as(i32p), C.byref(nk))
    t1 = time.perf_counter()
    t = det.timings()
    print("kpl_detect wall %.2f ms, device total %.2f ms (grid %.2f normals %.2f features %.2f nms %.2f), rc %d" % ((t1 - t0) * 1e3, t["total_ms"], t["grid_ms"], t["normals_ms"], t["features_ms"], t["nms_ms"], rc))
    t0 = time.perf_counter()
    det.compute(scores_out=sc, kp_out=kp)
    print("   det.compute() wall %.2f ms" % ((time.perf_counter() - t0) * 1e3))
# unpinned for comparison
up = x4.copy(); us = np.empty(n, np.float32); uk = np.empty(n, np.int32)
t0 = time.perf_counter()
L.kpl_detect(det._h, up.ctypes.data_as(f32p), 16, None, 0, None, n, us.ctypes.data_as(f32p), uk.ctypes.data_as(i32p), C.byref(nk))
print("pageable buffers: wall %.2f ms" % ((time.perf_counter() - t0) * 1e3))
