"""ctypes front-end of the CPU oracle (oracle/kpl_oracle.c) + Python-side test helpers.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under keypoint_learning_b200/ imports this module.
PARITY UNPINNED with respect to the reference binary (see kpl_oracle.c header); the pins that do
exist (cv2.ml.RTrees, scipy cKDTree, float64 PCA, hand-derived KATs) live in tests/.
"""
from __future__ import annotations

import ctypes as C
import gzip
import os
import re
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libkpl_oracle.so")
    src = os.path.join(_HERE, "kpl_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return so


REFERENCE = os.environ.get("KPL_REFERENCE", "/root/reference")
_REF = None


def build_ref(force: bool = False):
    """oracle/_ref/libkpl_ref.so: the reference's OWN code on the detection path -- src/KeypointLearning.cpp and
    the detector templates of include/KeypointLearning.h + impl/KeypointLearning.hpp -- compiled from the mounted
    reference tree against the stand-in environment oracle/ref_stubs/ (oracle/Makefile, target `ref`).  Returns
    the path, or None where the reference tree is not mounted and no prebuilt library travelled with the repo."""
    so = os.path.join(_HERE, "_ref", "libkpl_ref.so")
    src = os.path.join(REFERENCE, "src", "KeypointLearning.cpp")
    deps = [os.path.join(_HERE, "ref_wrap.cpp"), os.path.join(_HERE, "ref_stubs", "kplref_env.h")]
    if os.path.exists(src) and (force or not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(d) for d in deps)):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref", "REFERENCE=" + REFERENCE] + (["-B"] if force else []))
    return so if os.path.exists(so) else None


def ref_lib():
    """ctypes handle of oracle/_ref (the reference's own code), or None when it was never built."""
    global _REF
    if _REF is None:
        so = build_ref()
        if so is None:
            return None
        L = C.CDLL(so)
        f32p, ip, i32p, i64p = C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
        L.kplref_find_annulus_pair.argtypes = [C.c_int, C.c_float, C.c_float, ip, ip, f32p]
        L.kplref_find_bin_pair.argtypes = [C.c_int, C.c_float, ip, ip, f32p]
        L.kplref_annulus_sweep.argtypes = [C.c_int, C.c_float, f32p, C.c_long, ip, ip, f32p]
        L.kplref_bin_sweep.argtypes = [C.c_int, f32p, C.c_long, ip, ip, f32p]
        L.kplref_set_neighbours.argtypes = [C.c_double, i64p, i32p, C.c_double, i64p, i32p]
        L.kplref_set_forest.argtypes = [C.c_int, i32p, i32p, f32p, i32p, i32p, f32p]
        L.kplref_features.argtypes = [f32p, f32p, C.c_int64, C.c_double, C.c_int, C.c_int, i32p, C.c_int64, f32p]
        L.kplref_detect.argtypes = [f32p, f32p, C.c_int64, C.c_double, C.c_double, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, i32p, f32p]
        L.kplref_detect.restype = C.c_int64
        _REF = L
    return _REF


def ref_annulus_sweep(n_annulus, support, distances):
    d = np.ascontiguousarray(distances, np.float32)
    i = np.empty(len(d), np.int32); p = np.empty(len(d), np.int32); w = np.empty(len(d), np.float32)
    ref_lib().kplref_annulus_sweep(int(n_annulus), np.float32(support), _p(d, C.c_float), len(d), _p(i, C.c_int), _p(p, C.c_int), _p(w, C.c_float))
    return i, p, w


def ref_bin_sweep(n_bins, cosines):
    c = np.ascontiguousarray(cosines, np.float32)
    i = np.empty(len(c), np.int32); p = np.empty(len(c), np.int32); w = np.empty(len(c), np.float32)
    ref_lib().kplref_bin_sweep(int(n_bins), _p(c, C.c_float), len(c), _p(i, C.c_int), _p(p, C.c_int), _p(w, C.c_float))
    return i, p, w


def ref_neighbour_lists(xyz, radius, order, r_feat=None, cpr=4):
    """CSR neighbour lists for the stand-in pcl::search::KdTree of oracle/_ref: every point's radius neighbours
    (d2 < (float)(r*r), FLANN's FP32 expression) with the query itself FIRST -- slot 0 of a sorted search, the
    slot computePointFeatures skips -- and the others in `order`: 0 ascending point index, 1 canonical
    (cell key, index) of the grid built for r_feat, 2 ascending (d2, index) like a sorted kd-tree."""
    xyz = _xyz(xyz)
    n = len(xyz)
    off, idx = radius_neighbors(xyz, radius, np.arange(n, dtype=np.int32))
    q = np.repeat(np.arange(n, dtype=np.int64), np.diff(off))
    not_self = (idx != q).astype(np.int8)
    if order == 1:
        org, cell, dims = canon_grid(xyz, radius if r_feat is None else r_feat, cpr)
        sec = canon_keys(xyz, org, cell, dims)[idx]
    elif order == 2:
        d = xyz[q] - xyz[idx]
        sec = ((d[:, 0] * d[:, 0]) + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
    else:
        sec = np.zeros(len(idx), np.int8)
    perm = np.lexsort((idx, sec, not_self, q))
    return np.ascontiguousarray(off, np.int64), np.ascontiguousarray(idx[perm], np.int32)


def ref_features(xyz, normals4, r_feat, A, B, lists, qidx=None):
    """The reference's computePointsForTrainingFeatures / computePointFeatures (hpp:299-376) through oracle/_ref."""
    xyz = _xyz(xyz); nrm = np.ascontiguousarray(normals4, np.float32)
    q = np.arange(len(xyz), dtype=np.int32) if qidx is None else np.ascontiguousarray(qidx, np.int32)
    off, idx = lists
    L = ref_lib()
    L.kplref_set_neighbours(float(r_feat), _p(off, C.c_int64), _p(idx, C.c_int32), -1.0, None, None)
    out = np.empty((len(q), A * B), np.float32)
    rc = L.kplref_features(_p(xyz, C.c_float), _p(nrm, C.c_float), len(xyz), float(r_feat), A, B, _p(q, C.c_int32), len(q), _p(out, C.c_float))
    assert rc == 0, rc
    return out


def ref_detect(xyz, normals4, forest, r_feat, r_nms, th, A, B, lists_feat, lists_nms, non_maxima=True, draws_remove=False, draws_thr=0.0):
    """The reference's compute() -> detectKeypoints -> runForest (hpp:179-296) through oracle/_ref.
    Returns (keypoint indices, their scores); with non_maxima=False every point and its response."""
    xyz = _xyz(xyz); nrm = np.ascontiguousarray(normals4, np.float32)
    L = ref_lib()
    L.kplref_set_forest(forest["ntrees"], _p(forest["roots"], C.c_int32), _p(forest["var"], C.c_int32), _p(forest["thr"], C.c_float),
                        _p(forest["left"], C.c_int32), _p(forest["right"], C.c_int32), _p(forest["value"], C.c_float))
    of, xf = lists_feat
    on, xn = lists_nms if lists_nms is not None else (None, None)
    L.kplref_set_neighbours(float(r_feat), _p(of, C.c_int64), _p(xf, C.c_int32), float(r_nms),
                            _p(on, C.c_int64) if on is not None else None, _p(xn, C.c_int32) if xn is not None else None)
    kp = np.empty(len(xyz), np.int32); sc = np.empty(len(xyz), np.float32)
    cnt = L.kplref_detect(_p(xyz, C.c_float), _p(nrm, C.c_float), len(xyz), float(r_feat), float(r_nms), np.float32(th), A, B,
                          int(non_maxima), int(draws_remove), np.float32(draws_thr), _p(kp, C.c_int32), _p(sc, C.c_float))
    assert cnt >= 0, cnt
    return kp[:cnt].copy(), sc[:cnt].copy()


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        f32p, i32p, i64p, f64p = (C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_double))
        L.kplo_atan2f.restype = C.c_float; L.kplo_atan2f.argtypes = [C.c_float, C.c_float]
        L.kplo_cosf.restype = C.c_float; L.kplo_cosf.argtypes = [C.c_float]
        L.kplo_sinf.restype = C.c_float; L.kplo_sinf.argtypes = [C.c_float]
        L.kplo_find_annulus_pair.argtypes = [C.c_int, C.c_float, C.c_float, C.POINTER(C.c_int), C.POINTER(C.c_int), f32p]
        L.kplo_find_bin_pair.argtypes = [C.c_int, C.c_float, C.POINTER(C.c_int), C.POINTER(C.c_int), f32p]
        L.kplo_radius_counts.argtypes = [f32p, C.c_int64, C.c_double, i32p]
        L.kplo_radius_counts_brute.argtypes = [f32p, C.c_int64, C.c_double, i32p]
        L.kplo_radius_neighbors.argtypes = [f32p, C.c_int64, C.c_double, i32p, C.c_int64, i64p, i32p]
        L.kplo_normals_knn.argtypes = [f32p, C.c_int64, C.c_int, f32p, C.c_double, f32p]
        L.kplo_knn_indices.argtypes = [f32p, C.c_int64, C.c_int, C.c_double, i32p, f32p]
        L.kplo_normals_radius.argtypes = [f32p, C.c_int64, C.c_double, f32p, f32p]
        L.kplo_normals_radius_ordered.argtypes = [f32p, C.c_int64, C.c_double, f32p, C.c_int, C.c_int, f32p]
        L.kplo_canon_grid.argtypes = [f32p, C.c_int64, C.c_double, C.c_int, f64p, f64p, i32p]
        L.kplo_canon_keys.argtypes = [f32p, C.c_int64, f64p, C.c_double, i32p, i64p]
        L.kplo_features.argtypes = [f32p, f32p, C.c_int64, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int,
                                    f64p, C.c_double, i32p, i32p, C.c_int64, f32p]
        L.kplo_forest_sum.argtypes = [i32p, C.c_int, i32p, f32p, i32p, i32p, f32p, f32p, C.c_int64, C.c_int, f32p]
        L.kplo_scores.argtypes = [f32p, C.c_int64, C.c_int, f32p]
        L.kplo_mask_unscored.restype = C.c_int64
        L.kplo_mask_unscored.argtypes = [f32p, C.c_int64, f32p]
        L.kplo_nms.restype = C.c_int64
        L.kplo_nms.argtypes = [f32p, f32p, C.c_int64, C.c_double, C.c_double, i32p]
        L.kplo_nms_draws.restype = C.c_int64
        L.kplo_nms_draws.argtypes = [f32p, f32p, C.c_int64, C.c_double, C.c_double, C.c_float, i32p]
        L.kplo_detect.restype = C.c_int64
        L.kplo_detect.argtypes = [f32p, f32p, C.c_int64, C.c_int, C.c_int, f32p, C.c_int,
                                  C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int,
                                  i32p, C.c_int, i32p, f32p, i32p, i32p, f32p, f32p, f32p, i32p, f64p]
        L.kplo_num_threads.restype = C.c_int
        L.kplo_set_threads.argtypes = [C.c_int]
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _xyz(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] == 3, a.shape
    return a


# ---------------------------------------------------------------------------------------------
# scalar helpers
# ---------------------------------------------------------------------------------------------
def annulus_sweep(n_annulus, support, distances):
    d = np.ascontiguousarray(distances, np.float32)
    i = np.empty(len(d), np.int32); p = np.empty(len(d), np.int32); w = np.empty(len(d), np.float32)
    L = lib()
    L.kplo_annulus_sweep.argtypes = [C.c_int, C.c_float, C.POINTER(C.c_float), C.c_int64, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_float)]
    L.kplo_annulus_sweep(int(n_annulus), np.float32(support), _p(d, C.c_float), len(d), _p(i, C.c_int), _p(p, C.c_int), _p(w, C.c_float))
    return i, p, w


def bin_sweep(n_bins, cosines):
    c = np.ascontiguousarray(cosines, np.float32)
    i = np.empty(len(c), np.int32); p = np.empty(len(c), np.int32); w = np.empty(len(c), np.float32)
    L = lib()
    L.kplo_bin_sweep.argtypes = [C.c_int, C.POINTER(C.c_float), C.c_int64, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_float)]
    L.kplo_bin_sweep(int(n_bins), _p(c, C.c_float), len(c), _p(i, C.c_int), _p(p, C.c_int), _p(w, C.c_float))
    return i, p, w


def find_annulus_pair(n_annulus, distance, support):
    i, p, w = C.c_int(), C.c_int(), C.c_float()
    lib().kplo_find_annulus_pair(n_annulus, np.float32(distance), np.float32(support), C.byref(i), C.byref(p), C.byref(w))
    return i.value, p.value, np.float32(w.value)


def find_bin_pair(n_bins, cosine):
    i, p, w = C.c_int(), C.c_int(), C.c_float()
    lib().kplo_find_bin_pair(n_bins, np.float32(cosine), C.byref(i), C.byref(p), C.byref(w))
    return i.value, p.value, np.float32(w.value)


def atan2f(y, x):
    return np.float32(lib().kplo_atan2f(np.float32(y), np.float32(x)))


def cosf(x):
    return np.float32(lib().kplo_cosf(np.float32(x)))


def sinf(x):
    return np.float32(lib().kplo_sinf(np.float32(x)))


# ---------------------------------------------------------------------------------------------
# stages
# ---------------------------------------------------------------------------------------------
def radius_counts(xyz, radius, brute=False):
    xyz = _xyz(xyz)
    out = np.empty(len(xyz), np.int32)
    if brute:
        lib().kplo_radius_counts_brute(_p(xyz, C.c_float), len(xyz), float(radius), _p(out, C.c_int32))
    else:
        rc = lib().kplo_radius_counts(_p(xyz, C.c_float), len(xyz), float(radius), _p(out, C.c_int32))
        assert rc == 0, rc
    return out


def radius_neighbors(xyz, radius, qidx=None):
    """-> (offsets[m+1], indices) ; neighbour lists ascending by index, self included."""
    xyz = _xyz(xyz)
    q = None if qidx is None else np.ascontiguousarray(qidx, np.int32)
    m = len(xyz) if q is None else len(q)
    off = np.empty(m + 1, np.int64)
    rc = lib().kplo_radius_neighbors(_p(xyz, C.c_float), len(xyz), float(radius), _p(q, C.c_int32), m, _p(off, C.c_int64), None)
    assert rc == 0, rc
    idx = np.empty(int(off[-1]), np.int32)
    rc = lib().kplo_radius_neighbors(_p(xyz, C.c_float), len(xyz), float(radius), _p(q, C.c_int32), m, _p(off, C.c_int64), _p(idx, C.c_int32))
    assert rc == 0, rc
    return off, idx


def normals_knn(xyz, k=10, viewpoint=(0.0, 0.0, 0.0), cell=0.0):
    xyz = _xyz(xyz)
    vp = np.asarray(viewpoint, np.float32)
    out = np.empty((len(xyz), 4), np.float32)
    rc = lib().kplo_normals_knn(_p(xyz, C.c_float), len(xyz), int(k), _p(vp, C.c_float), float(cell), _p(out, C.c_float))
    assert rc == 0, rc
    return out


def knn_indices(xyz, k=10, cell=0.0):
    xyz = _xyz(xyz)
    idx = np.empty((len(xyz), k), np.int32)
    d2 = np.empty((len(xyz), k), np.float32)
    rc = lib().kplo_knn_indices(_p(xyz, C.c_float), len(xyz), int(k), float(cell), _p(idx, C.c_int32), _p(d2, C.c_float))
    assert rc == 0, rc
    return idx, d2


def normals_radius(xyz, radius, viewpoint=(0.0, 0.0, 0.0), order=0, cpr=4):
    """order 0: PCL's sorted-search order (d2, index); order 1: canonical (cell key, index), the device order."""
    xyz = _xyz(xyz)
    vp = np.asarray(viewpoint, np.float32)
    out = np.empty((len(xyz), 4), np.float32)
    rc = lib().kplo_normals_radius_ordered(_p(xyz, C.c_float), len(xyz), float(radius), _p(vp, C.c_float), int(order), int(cpr), _p(out, C.c_float))
    assert rc == 0, rc
    return out


def normals_integral_image(xyz_hw3, smoothing=5.0, viewpoint=(0.0, 0.0, 0.0)):
    """pcl::IntegralImageNormalEstimation(SIMPLE_3D_GRADIENT, smoothing) on an organized cloud (height, width, 3)."""
    a = np.ascontiguousarray(xyz_hw3, np.float32)
    assert a.ndim == 3 and a.shape[2] == 3
    h, w = a.shape[:2]
    out = np.empty((h * w, 4), np.float32)
    vp = np.asarray(viewpoint, np.float32)
    rc = lib().kplo_normals_integral_image(_p(a, C.c_float), int(w), int(h), C.c_float(smoothing), _p(vp, C.c_float), _p(out, C.c_float))
    assert rc == 0, rc
    return out


def uniform_sample(xyz, leaf, centre=False):
    """pcl::UniformSampling (call site src/main_test_detector.cpp:145-157), restated in FP32 numpy [3P-recalled, PCL 1.8.0
    filters/impl/uniform_sampling.hpp]: ijk = floor(p * (1/leaf)); per voxel keep the point with the smallest
    (p.getVector4fMap() - ijk.cast<float>()).squaredNorm() -- the distance to the voxel INDEX vector taken as a point, over
    (x, y, z, 1) - (i, j, k, 0), SSE2 packet sum (d0 + d2) + (d1 + d3) -- a later point replacing the kept one only when
    strictly closer (ties: lower index).  centre=True: closest to the voxel centre (ijk + 0.5) * leaf instead.
    Ascending indices (PCL's own output order is that of an unordered_map)."""
    x = _xyz(xyz)
    leaf = np.float32(leaf)
    inv = np.float32(1.0) / leaf
    f = np.floor(x * inv)
    mb = np.floor(x.min(axis=0) * inv)
    db = (np.floor(x.max(axis=0) * inv) - mb + 1).astype(np.int64)
    ijk = (f - mb).astype(np.int64)
    key = (ijk[:, 2] * db[1] + ijk[:, 1]) * db[0] + ijk[:, 0]
    if centre:
        e = (f + np.float32(0.5)) * leaf - x
        d = e[:, 0] * e[:, 0] + (e[:, 1] * e[:, 1] + e[:, 2] * e[:, 2])
    else:
        e = x - f
        d = (e[:, 0] * e[:, 0] + e[:, 2] * e[:, 2]) + (e[:, 1] * e[:, 1] + np.float32(1.0))
    order = np.lexsort((np.arange(len(x)), d, key))
    first = np.ones(len(x), bool)
    first[1:] = key[order][1:] != key[order][:-1]
    return np.sort(order[first]).astype(np.int32)


def canon_grid(xyz, r_feat, cpr=4):
    xyz = _xyz(xyz)
    org = np.empty(3, np.float64); cell = C.c_double(); dims = np.empty(3, np.int32)
    rc = lib().kplo_canon_grid(_p(xyz, C.c_float), len(xyz), float(r_feat), int(cpr), _p(org, C.c_double), C.byref(cell), _p(dims, C.c_int32))
    assert rc == 0, rc
    return org, cell.value, dims


def canon_keys(xyz, org, cell, dims):
    xyz = _xyz(xyz)
    org = np.ascontiguousarray(org, np.float64); dims = np.ascontiguousarray(dims, np.int32)
    keys = np.empty(len(xyz), np.int64)
    lib().kplo_canon_keys(_p(xyz, C.c_float), len(xyz), _p(org, C.c_double), float(cell), _p(dims, C.c_int32), _p(keys, C.c_int64))
    return keys


def features(xyz, normals4, r_feat, A=5, B=10, order=1, cpr=4, qidx=None, canon=None):
    """order: 0 ascending index, 1 canonical (cell key, index), 2 traversal (timing only).
    canon = (org, cell, dims) overrides the canonical grid derived from this cloud (multi-GPU slabs)."""
    xyz = _xyz(xyz)
    nrm = np.ascontiguousarray(normals4, np.float32)
    assert nrm.shape == (len(xyz), 4)
    q = None if qidx is None else np.ascontiguousarray(qidx, np.int32)
    m = len(xyz) if q is None else len(q)
    out = np.empty((m, A * B), np.float32)
    if canon is not None:
        org = np.ascontiguousarray(canon[0], np.float64); cell = float(canon[1]); dims = np.ascontiguousarray(canon[2], np.int32)
    else:
        org = dims = None; cell = 0.0
    rc = lib().kplo_features(_p(xyz, C.c_float), _p(nrm, C.c_float), len(xyz), float(r_feat), A, B, order, cpr,
                             _p(org, C.c_double), cell, _p(dims, C.c_int32), _p(q, C.c_int32), m, _p(out, C.c_float))
    assert rc == 0, rc
    return out


def set_normalize_mode(reciprocal):
    """Eigen version of the per-annulus row.normalize(): False = divide (Eigen >= 3.3), True = multiply by 1/norm (3.2.x)."""
    lib().kplo_set_normalize_mode(int(bool(reciprocal)))
    r = ref_lib()
    if r is not None:
        r.kplref_set_normalize_mode(int(bool(reciprocal)))


def forest_sum(forest, feat):
    feat = np.ascontiguousarray(feat, np.float32)
    m, F = feat.shape
    out = np.empty(m, np.float32)
    lib().kplo_forest_sum(_p(forest["roots"], C.c_int32), forest["ntrees"], _p(forest["var"], C.c_int32), _p(forest["thr"], C.c_float),
                          _p(forest["left"], C.c_int32), _p(forest["right"], C.c_int32), _p(forest["value"], C.c_float),
                          _p(feat, C.c_float), m, F, _p(out, C.c_float))
    return out


def forest_fragile(forest, feat, eps=1e-5):
    """Per row: 1 if a split decided on the row's walks had |x[var] - thr| <= eps (the fragile decisions of BASELINE.md s5)."""
    feat = np.ascontiguousarray(feat, np.float32)
    m, F = feat.shape
    out = np.empty(m, np.uint8)
    L = lib()
    L.kplo_forest_fragile.restype = C.c_int64
    L.kplo_forest_fragile(_p(forest["roots"], C.c_int32), forest["ntrees"], _p(forest["var"], C.c_int32), _p(forest["thr"], C.c_float),
                          _p(forest["left"], C.c_int32), _p(forest["right"], C.c_int32), _p(feat, C.c_float), C.c_int64(m), F,
                          C.c_float(eps), _p(out, C.c_uint8))
    return out


def scores_from_sums(sums, ntrees):
    sums = np.ascontiguousarray(sums, np.float32)
    out = np.empty_like(sums)
    lib().kplo_scores(_p(sums, C.c_float), len(sums), int(ntrees), _p(out, C.c_float))
    return out


def scores(forest, feat, normals4=None):
    """forest sums -> 1 - sum/ntrees (hpp:287); points without a finite normal get NaN (hpp:277)."""
    sc = scores_from_sums(forest_sum(forest, feat), forest["ntrees"])
    if normals4 is not None:
        nrm = np.ascontiguousarray(normals4, np.float32)
        lib().kplo_mask_unscored(_p(nrm, C.c_float), len(sc), _p(sc, C.c_float))
    return sc


def nms(xyz, scores, r_nms, th, draws_remove=False, draws_thr=0.0):
    xyz = _xyz(xyz)
    scores = np.ascontiguousarray(scores, np.float32)
    kp = np.empty(len(xyz), np.int32)
    th = float(np.float32(th))  # TestDetector passes a float threshold into a double member
    if draws_remove:
        cnt = lib().kplo_nms_draws(_p(xyz, C.c_float), _p(scores, C.c_float), len(xyz), float(r_nms), th, np.float32(draws_thr), _p(kp, C.c_int32))
    else:
        cnt = lib().kplo_nms(_p(xyz, C.c_float), _p(scores, C.c_float), len(xyz), float(r_nms), th, _p(kp, C.c_int32))
    assert cnt >= 0, cnt
    return kp[:cnt].copy()


def detect(xyz, forest, r_feat=20.0, r_nms=4.0, th=0.85, A=5, B=10, normals4=None, normals_mode=1, k=10,
           viewpoint=(0.0, 0.0, 0.0), flip=False, order=1, cpr=4, threads=None):
    """Whole pipeline in one C call (used for the CPU baseline timing).  Returns dict."""
    xyz = _xyz(xyz)
    n = len(xyz)
    if threads:
        lib().kplo_set_threads(int(threads))
    if normals4 is None:
        nrm = np.empty((n, 4), np.float32)
    else:
        nrm = np.ascontiguousarray(normals4, np.float32).copy(); normals_mode = 0
    vp = np.asarray(viewpoint, np.float32)
    feat = np.empty((n, A * B), np.float32); sc = np.empty(n, np.float32); kp = np.empty(n, np.int32)
    ms = np.zeros(5, np.float64)
    cnt = lib().kplo_detect(_p(xyz, C.c_float), _p(nrm, C.c_float), n, normals_mode, k, _p(vp, C.c_float), int(flip),
                            float(np.float32(r_feat)), float(np.float32(r_nms)), float(np.float32(th)), A, B, order, cpr,
                            _p(forest["roots"], C.c_int32), forest["ntrees"], _p(forest["var"], C.c_int32), _p(forest["thr"], C.c_float),
                            _p(forest["left"], C.c_int32), _p(forest["right"], C.c_int32), _p(forest["value"], C.c_float),
                            _p(feat, C.c_float), _p(sc, C.c_float), _p(kp, C.c_int32), _p(ms, C.c_double))
    assert cnt >= 0, cnt
    return dict(normals=nrm, features=feat, scores=sc, keypoints=kp[:cnt].copy(),
                stage_ms=dict(normals=ms[0], features=ms[1], forest=ms[2], nms=ms[3], total=ms[4]))


def num_threads():
    return lib().kplo_num_threads()


# ---------------------------------------------------------------------------------------------
# opencv_ml_rtrees YAML(.gz) -> flat arrays  (independent Python restatement of OpenCV 3.2
# modules/ml/src/tree.cpp readTree/readNode/readSplit; SURVEY.md s8c schema)
# ---------------------------------------------------------------------------------------------
_TOK = re.compile(r"\b(depth|value|var|le|gt|in|not_in|ntrees|var_count|is_classifier|format)\s*:\s*([-+0-9.eE]+|\[)")


def load_forest_yaml(path):
    opener = gzip.open if open(path, "rb").read(2) == b"\x1f\x8b" else open
    with opener(path, "rt") as f:
        text = f.read()
    if "opencv_ml_rtrees" not in text:
        raise ValueError("not an opencv_ml_rtrees file")
    mt = re.search(r"^\s*trees\s*:", text, re.M)
    if not mt:
        raise ValueError("no trees")
    head, body = text[:mt.start()], text[mt.end():]
    meta = {k: float(v) for k, v in _TOK.findall(head) if v != "["}
    var, thr, left, right, value, roots = [], [], [], [], [], []
    for tree_txt in body.split("nodes:")[1:]:
        base = len(var)
        roots.append(base)
        depths = []
        got_le = set()
        stack = []  # open internal nodes waiting for children: [node, n_children]
        cur = None
        for k, v in _TOK.findall(tree_txt):
            if k == "depth":
                cur = len(var)
                var.append(-1); thr.append(0.0); left.append(-1); right.append(-1); value.append(0.0)
                depths.append(int(float(v)))
                if stack:
                    par = stack[-1]
                    if par[1] == 0:
                        left[par[0]] = cur
                    else:
                        right[par[0]] = cur
                    par[1] += 1
                    if par[1] == 2:
                        stack.pop()
                # a node becomes "open" once we see its split (below)
            elif k == "value":
                value[cur] = float(v)
            elif k == "var":
                if var[cur] == -1:  # primary split only (surrogates ignored)
                    var[cur] = int(float(v))
                    stack.append([cur, 0])
            elif k == "le":
                if cur not in got_le:
                    got_le.add(cur); thr[cur] = float(v)
            elif k in ("gt", "in", "not_in"):
                raise ValueError("unsupported split type %r" % k)
        if stack:
            raise ValueError("truncated tree")
    F = dict(ntrees=len(roots), var_count=int(meta.get("var_count", 0)),
             roots=np.asarray(roots, np.int32), var=np.asarray(var, np.int32), thr=np.asarray(thr, np.float32),
             left=np.asarray(left, np.int32), right=np.asarray(right, np.int32), value=np.asarray(value, np.float32))
    if "ntrees" in meta and int(meta["ntrees"]) != F["ntrees"]:
        raise ValueError("ntrees mismatch")
    return F


# ---------------------------------------------------------------------------------------------
# PCD v0.7 reader (test-side; the product has its own C++ reader)
# ---------------------------------------------------------------------------------------------
def read_pcd_xyz(path):
    with open(path, "rb") as f:
        raw = f.read()
    hdr_end = 0; fields = []; n = 0; data = None; sizes = []
    while True:
        nl = raw.index(b"\n", hdr_end)
        line = raw[hdr_end:nl].decode("ascii", "replace").strip()
        hdr_end = nl + 1
        t = line.split()
        if not t or t[0].startswith("#"):
            continue
        if t[0] == "FIELDS": fields = t[1:]
        elif t[0] == "SIZE": sizes = [int(x) for x in t[1:]]
        elif t[0] == "POINTS": n = int(t[1])
        elif t[0] == "DATA":
            data = t[1]; break
    ix = [fields.index(c) for c in "xyz"]
    if data == "ascii":
        arr = np.loadtxt(raw[hdr_end:].decode("ascii").splitlines(), dtype=np.float32, ndmin=2)
        return np.ascontiguousarray(arr[:n, ix])
    if data == "binary":
        assert all(s == 4 for s in sizes)
        arr = np.frombuffer(raw, np.float32, n * len(fields), hdr_end).reshape(n, len(fields))
        return np.ascontiguousarray(arr[:, ix])
    raise ValueError("unsupported PCD DATA " + str(data))
