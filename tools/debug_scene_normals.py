"""GPU-box diagnostic: where do the device normals of the closed-surface scene crop differ from the oracle's?"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import keypoint_learning_b200 as K
from keypoint_learning_b200 import synth
from oracle import oracle as O

xyz, vp = synth.scene_closed_surfaces(1_250_000, seed=4321)
xyz = synth.cube_crop(xyz, 300_000)
d = K.KeypointLearningDetector()
d.setNormalsMode(1, k=10, viewpoint=vp)
g = d.computeNormals(xyz)
o = O.normals_knn(xyz, 10, vp)
gu, ou = g.view(np.uint32), o.view(np.uint32)
diff = np.nonzero((gu != ou).any(axis=1))[0]
print("points", len(xyz), "differing rows", len(diff))
both_nan = np.isnan(g).any(axis=1) & np.isnan(o).any(axis=1)
print("rows where both hold NaN:", int(both_nan.sum()), " differing rows that are NOT both-NaN:", int((~both_nan[diff]).sum()))
for i in diff[~both_nan[diff]][:10]:
    print(i, xyz[i], "gpu", g[i], gu[i], "oracle", o[i], ou[i])
for i in diff[both_nan[diff]][:3]:
    print("nan row", i, xyz[i], "gpu", gu[i], "oracle", ou[i])
