"""Small end-to-end run of every kernel for compute-sanitizer (memcheck / racecheck / initcheck):
    compute-sanitizer --tool memcheck python tools/sanitize.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import keypoint_learning_b200 as K  # noqa: E402
from keypoint_learning_b200 import synth  # noqa: E402

forest = os.path.join(ROOT, "tests", "golden", "forests", "synthetic-SHOT-like-T50-D10.yaml.gz")
xyz, vp = synth.view_25d(96, 64, seed=11)
d = K.KeypointLearningDetector()
d.setNAnnulus(5); d.setNBins(10); d.setNonMaxima(True); d.setNonMaxRadius(4.0); d.setNonMaximaDrawsRemove(False)
d.setPredictionThreshold(float(np.float32(0.5))); d.setRadiusSearch(20.0)
assert d.loadForest(forest)
for mode in (1, 2):                                     # kNN-10 and radius normals
    d.setNormalsMode(mode, k=10, viewpoint=vp)
    d.setInputCloud(xyz); d.setNormals(None)
    kp, idx = d.compute()
    print("mode", mode, "keypoints", len(idx), "launches", d.stats()["kernel_launches"])
d.setNormalsMode(1, k=7, viewpoint=vp)                  # per-thread kNN kernel
d.compute()
d.setNormalsMode(1, k=10, viewpoint=vp)
d.setNonMaximaDrawsRemove(True); d.setNonMaximaDrawsThreshold(2.0)
print("draws-remove keypoints", len(d.compute()[1]))
d.setNonMaximaDrawsRemove(False)
d.keepIntermediates(True); d.compute(); d.fetch("features", len(xyz), 50); d.fetch("normals", len(xyz), 4)
nrm = d.computeNormals(xyz)
f = d.computePointsForTrainingFeatures(np.arange(0, len(xyz), 7, dtype=np.int32))
print("training features", f.shape)
print("subsampled", len(d.uniformSample(xyz, 3.0)))
cnt, h = d.radiusStats(xyz, 4.0)
role = np.full(len(xyz), 3, np.uint8); role[::5] = 0; role[1::5] = 1
d.setInputCloud(xyz); d.setNormals(nrm)
print("with roles", len(d.compute(role=role)[1]))
# batch of views in one pass, fragile-split report, nearest-point snap
views = [xyz, synth.view_25d(80, 48, seed=12)[0], np.ascontiguousarray(xyz[:9] + np.float32(200.0))]
d.setNormals(None); d.setReportFragile(True)
sc_b, kp_b = d.computeBatch(views)
print("batch keypoints", [len(k) for k in kp_b], "fragile", d.stats()["n_fragile_points"])
d.setReportFragile(False)
d.nearest(xyz, xyz[:100] + np.float32(0.3))
d.close()
# slab sharding: three ranks of an in-process group on this device (pack / halo-score / to-global / record kernels,
# runs cut at the ownership faces, longest-first work list)
from keypoint_learning_b200 import shard  # noqa: E402
big, vp2 = synth.view_25d(400, 96, seed=13)
dets = []
for _ in range(3):
    t = K.KeypointLearningDetector()
    t.setNAnnulus(5); t.setNBins(10); t.setNonMaxima(True); t.setNonMaxRadius(4.0); t.setNonMaximaDrawsRemove(False)
    t.setPredictionThreshold(float(np.float32(0.5))); t.setRadiusSearch(20.0); t.setNormalsMode(1, k=10, viewpoint=vp2)
    assert t.loadForest(forest)
    dets.append(t)
plan = shard.plan_slabs(big, 20.0, 4.0, 4, 3)
jobs = [shard.SlabJob(t, big, plan, r, None) for r, t in enumerate(dets)]
kp, scores = shard.detect_group_widening(jobs, big, 20.0, 4.0, 4)
print("sharded keypoints", len(kp))
for j in jobs:
    j.close()
for t in dets:
    t.close()
print("sanitize run complete")
