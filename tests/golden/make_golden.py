"""Generates the committed fixtures under tests/golden/.  Run ONCE in the build container:

    python tests/golden/make_golden.py

It is the only script that reads /root/reference (the three bundled views,
data/point_cloud_test/cheff00{0,1,2}.pcd) -- nothing in tests/, smoke() or bench.py does.

Outputs
  views/cheff00X.npz          the bundled views as float32 arrays (PCD parsed by oracle.read_pcd_xyz)
  forests/*.yaml.gz           synthetic stand-ins for the missing data/forest/*.gz, trained with
                              cv2.ml.RTrees using TrainDetector's parameter set
                              (src/main_train_detector.cpp:267-278) on ORACLE features of cheff000
  golden_cheff00X.npz         oracle outputs: neighbour counts, sha256 of the full normal / feature /
                              score arrays, every 97th row of them, and the full keypoint list
  kat_bins.npz                known answers of findAnnulusPair / findBinPair (SURVEY.md App. B inputs)
The reference publishes no golden outputs (parity unpinned); these vectors pin the ORACLE so that a
later change to it cannot go unnoticed, and give the GPU tests a fixed target.
"""
import gzip
import hashlib
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

REF = "/root/reference/data/point_cloud_test"
VIEWS = ("cheff000", "cheff001", "cheff002")
R_FEAT, R_NMS, TH = 20.0, 4.0, 0.85
STRIDE = 97


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def train_forest(feat, labels, ntrees, depth, msc, path):
    import cv2
    rt = cv2.ml.RTrees_create()
    rt.setMaxDepth(depth); rt.setMinSampleCount(msc); rt.setRegressionAccuracy(0); rt.setMaxCategories(15)
    rt.setUseSurrogates(False); rt.setCalculateVarImportance(True); rt.setActiveVarCount(0)
    rt.setTermCriteria((cv2.TERM_CRITERIA_MAX_ITER + cv2.TERM_CRITERIA_EPS, ntrees, 1e-6))
    n = len(feat)
    n_train = int(n * 0.8)   # main_train_detector.cpp:496-499: first 80 % of the rows are the train subset
    sample_idx = np.zeros(n, np.uint8); sample_idx[:n_train] = 1
    var_type = np.zeros(feat.shape[1] + 1, np.uint8); var_type[-1] = cv2.ml.VAR_CATEGORICAL
    td = cv2.ml.TrainData_create(feat.astype(np.float32), cv2.ml.ROW_SAMPLE, labels.astype(np.int32), sampleIdx=sample_idx, varType=var_type)
    rt.train(td)
    tmp = path[:-3] if path.endswith(".gz") else path
    rt.save(tmp)
    if path.endswith(".gz"):
        with open(tmp, "rb") as f, gzip.GzipFile(path, "wb", mtime=0) as g:
            shutil.copyfileobj(f, g)
        os.remove(tmp)
    return path


def saliency_labels(feat, A, B, rng, frac=0.2, noise=0.05):
    """label 0 = keypoint, 1 = not keypoint (main_train_detector.cpp:405-407).  'Salient' = a lot of
    normal variation in the inner annuli (mass away from the first cosine bins), plus label noise."""
    f = feat.reshape(len(feat), A, B)
    w = np.arange(B, dtype=np.float64)
    s = (f[:, : max(1, A // 2 + 1), :] * w).sum(axis=(1, 2))
    thr = np.quantile(s, 1.0 - frac)
    lab = np.where(s >= thr, 0, 1)
    flip = rng.random(len(lab)) < noise
    return np.where(flip, 1 - lab, lab)


def main():
    os.makedirs(os.path.join(HERE, "views"), exist_ok=True)
    os.makedirs(os.path.join(HERE, "forests"), exist_ok=True)
    clouds = {}
    for v in VIEWS:
        xyz = O.read_pcd_xyz(os.path.join(REF, v + ".pcd"))
        clouds[v] = xyz
        np.savez_compressed(os.path.join(HERE, "views", v + ".npz"), xyz=xyz)
        print(v, xyz.shape)

    # ---- forests, trained on oracle features of cheff000 --------------------------------------
    rng = np.random.default_rng(7)
    xyz0 = clouds["cheff000"]
    nrm0 = O.normals_knn(xyz0, 10)
    sub = rng.permutation(len(xyz0))[:24000].astype(np.int32)
    specs = [  # name, A, B, r_feat, ntrees, depth, msc
        ("synthetic-T100-D15", 5, 10, 20.0, 100, 15, 10),
        ("synthetic-SHOT-like-T50-D10", 5, 10, 20.0, 50, 10, 10),
        ("synthetic-FPFH-like-T30-D25", 5, 10, 20.0, 30, 25, 5),
        ("synthetic-A10xB5-T20-D8", 10, 5, 20.0, 20, 8, 10),
        ("synthetic-A4xB8-T20-D8", 4, 8, 20.0, 20, 8, 10),
        ("synthetic-A8xB16-T20-D8", 8, 16, 20.0, 20, 8, 10),
    ]
    feats_cache = {}
    for name, A, B, r, T, D, msc in specs:
        key = (A, B, r)
        if key not in feats_cache:
            feats_cache[key] = O.features(xyz0, nrm0, r, A, B, order=1, qidx=sub)
        feat = feats_cache[key]
        lab = saliency_labels(feat, A, B, np.random.default_rng(11))
        p = train_forest(feat, lab, T, D, msc, os.path.join(HERE, "forests", name + ".yaml.gz"))
        F = O.load_forest_yaml(p)
        print(name, "nodes", len(F["var"]), "bytes", os.path.getsize(p))

    # ---- golden oracle outputs on the three views, TestDetector defaults ----------------------
    forest = O.load_forest_yaml(os.path.join(HERE, "forests", "synthetic-T100-D15.yaml.gz"))
    for v in VIEWS:
        xyz = clouds[v]
        c20 = O.radius_counts(xyz, R_FEAT)
        c4 = O.radius_counts(xyz, R_NMS)
        nrm = O.normals_knn(xyz, 10)
        feat = O.features(xyz, nrm, R_FEAT, 5, 10, order=1)
        feat0 = O.features(xyz, nrm, R_FEAT, 5, 10, order=0)
        sums = O.forest_sum(forest, feat)
        sc = O.scores_from_sums(sums, forest["ntrees"])
        kp = O.nms(xyz, sc, R_NMS, TH)
        sc0 = O.scores_from_sums(O.forest_sum(forest, feat0), forest["ntrees"])
        kp0 = O.nms(xyz, sc0, R_NMS, TH)
        np.savez_compressed(
            os.path.join(HERE, "golden_%s.npz" % v),
            counts_r20=c20.astype(np.int32), counts_r4=c4.astype(np.int32), pairs_r20=np.int64(c20.sum()),
            normals_sha=sha(nrm), features_sha=sha(feat), scores_sha=sha(sc),
            normals_rows=nrm[::STRIDE], features_rows=feat[::STRIDE], scores=sc,
            keypoints=kp, n_above=np.int64(((sc.astype(np.float64)) >= float(np.float32(TH))).sum()),
            features_order0_maxdiff=np.float32(np.abs(feat - feat0).max()),
            scores_order0_ndiff=np.int64((sc != sc0).sum()), keypoints_order0=kp0,
            stride=np.int64(STRIDE))
        print(v, "pairs", c20.sum(), "kp", len(kp), "above", int((sc >= np.float32(TH)).sum()),
              "| order0: max feat diff %.3g, score diffs %d, kp %d" % (np.abs(feat - feat0).max(), (sc != sc0).sum(), len(kp0)))

    # ---- binning KATs -----------------------------------------------------------------------
    d = np.array([0, 1, 2, 3, 3.9999998, 4, 5, 6, 10, 17.9, 18, 19, 19.999998, 20], np.float32)
    c = np.array([-0.25, 0, 5.96e-8, 0.05, 0.1, 0.15, 0.2, 0.3, 0.99999994, 1.0, 1.1, 1.9, 1.95, 2.0, 2.5], np.float32)
    ann = np.array([O.find_annulus_pair(5, x, 20.0) for x in d], np.float64)
    bins = np.array([O.find_bin_pair(10, x) for x in c], np.float64)
    np.savez(os.path.join(HERE, "kat_bins.npz"), distance=d, cosine=c, annulus=ann, bins=bins)


if __name__ == "__main__":
    main()
