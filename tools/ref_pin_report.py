"""Writes profiles/r1_reference_pin.txt: how many values of the oracle were compared with oracle/_ref (the
reference's own code compiled from the mounted tree) and how many differed.  CPU only."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

assert O.ref_lib() is not None, "oracle/_ref is not built (reference tree not mounted)"
out = []
rng = np.random.default_rng(5)
n_cmp = n_bad = 0
for A, support in ((5, 20.0), (4, 7.5), (8, 40.0), (10, 13.37), (16, 20.0)):
    d = rng.uniform(0, support, 1_000_000).astype(np.float32)
    g, r = O.annulus_sweep(A, support, d), O.ref_annulus_sweep(A, support, d)
    n_cmp += 3 * len(d); n_bad += sum(int(np.sum(a.view(np.uint32) != b.view(np.uint32))) for a, b in zip(g, r))
for B in (10, 5, 8, 16, 3):
    c = rng.uniform(-0.5, 2.5, 1_000_000).astype(np.float32)
    g, r = O.bin_sweep(B, c), O.ref_bin_sweep(B, c)
    n_cmp += 3 * len(c); n_bad += sum(int(np.sum(a.view(np.uint32) != b.view(np.uint32))) for a, b in zip(g, r))
out.append("findAnnulusPair / findBinPair (src/KeypointLearning.cpp:41-92): %d values compared, %d differ" % (n_cmp, n_bad))

views = {v: np.load(os.path.join(ROOT, "tests", "golden", "views", v + ".npz"))["xyz"] for v in ("cheff000", "cheff001", "cheff002")}
forest = O.load_forest_yaml(os.path.join(ROOT, "tests", "golden", "forests", "synthetic-T100-D15.yaml.gz"))
for name, xyz in views.items():
    xyz = np.ascontiguousarray(xyz[:6000])
    nrm = O.normals_knn(xyz, 10)
    lf = O.ref_neighbour_lists(xyz, 20.0, 1)
    f_ref = O.ref_features(xyz, nrm, 20.0, 5, 10, lf)
    f_orc = O.features(xyz, nrm, 20.0, 5, 10, order=1)
    sc = O.scores(forest, f_orc, nrm)
    _, sc_ref = O.ref_detect(xyz, nrm, forest, 20.0, 4.0, 0.85, 5, 10, lf, None, non_maxima=False)
    ln = O.ref_neighbour_lists(xyz, 4.0, 0)
    kp_ref, _ = O.ref_detect(xyz, nrm, forest, 20.0, 4.0, 0.85, 5, 10, lf, ln)
    kp = O.nms(xyz, sc, 4.0, 0.85)
    kpd_ref, _ = O.ref_detect(xyz, nrm, forest, 20.0, 4.0, 0.85, 5, 10, lf, ln, draws_remove=True, draws_thr=1.5)
    kpd = O.nms(xyz, sc, 4.0, 0.85, draws_remove=True, draws_thr=1.5)
    out.append("%s[:6000]: computePointFeatures %d floats, %d differ; runForest %d scores, %d differ; detectKeypoints %d keypoints (reference %d), "
               "identical: %s; draws-remove %d (reference %d), identical: %s"
               % (name, f_ref.size, int(np.sum(f_ref.view(np.uint32) != f_orc.view(np.uint32))), len(sc), int(np.sum(sc.view(np.uint32) != sc_ref.view(np.uint32))),
                  len(kp), len(kp_ref), bool(np.array_equal(kp, kp_ref)), len(kpd), len(kpd_ref), bool(np.array_equal(kpd, kpd_ref))))
text = ("# oracle/kpl_oracle.c against oracle/_ref/libkpl_ref.so (the reference's own detector templates and binning helpers, compiled\n"
        "# from /root/reference against oracle/ref_stubs/kplref_env.h).  Bit-level comparison, canonical neighbour order, T100-D15 forest.\n"
        + "\n".join(out) + "\n")
open(os.path.join(ROOT, "profiles", "r1_reference_pin.txt"), "w").write(text)
print(text)
