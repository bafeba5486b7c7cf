"""GPU-box experiment: stand-alone k-NN normals (data-sized grid) vs the normals stage of the fused pipeline (feature grid)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import keypoint_learning_b200 as K
from keypoint_learning_b200 import synth
for name, (xyz, vp) in (("scene10m", synth.scene_closed_surfaces(10_000_000, seed=4321)), ("view1m", synth.view_25d(1250, 800, seed=1234))):
    d = K.KeypointLearningDetector()
    d.setNormalsMode(1, k=10, viewpoint=vp)
    for rep in range(3):
        d.computeNormals(xyz)
    t = d.timings(); st = d.stats()
    print(name, "stand-alone kpl_normals: grid %.3f ms normals %.3f ms, cells %d (cell %.3f)" % (t["grid_ms"], t["normals_ms"], st["grid_cells"], st["grid_cell"]))
    d.close()
