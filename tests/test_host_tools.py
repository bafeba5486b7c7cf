"""The C++ host side of the drop-in: PCD v0.7 reader / writers and UniformSampling stand-in (CPU tests,
through the pcd_tool executable), the TestDetector CLI (argument handling on CPU; the full run against
the golden keypoints on the GPU)."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "keypoint_learning_b200")


@pytest.fixture(scope="module")
def tools(kpl):
    from keypoint_learning_b200 import build as B
    B.build_host()
    assert os.path.exists(B.TEST_DETECTOR) and os.path.exists(B.PCD_TOOL)
    return B


def write_pcd(path, xyz, data="ascii", extra_field=False, viewpoint=(0, 0, 0)):
    n = len(xyz)
    fields = "x y z" + (" intensity" if extra_field else "")
    nf = 4 if extra_field else 3
    hdr = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS %s\nSIZE %s\nTYPE %s\nCOUNT %s\nWIDTH %d\nHEIGHT 1\n"
           "VIEWPOINT %g %g %g 1 0 0 0\nPOINTS %d\nDATA %s\n" % (fields, " ".join(["4"] * nf), " ".join(["F"] * nf), " ".join(["1"] * nf), n, *viewpoint, n, data))
    rows = np.zeros((n, nf), np.float32)
    rows[:, :3] = xyz
    if extra_field:
        rows[:, 3] = np.arange(n)
    with open(path, "wb") as f:
        f.write(hdr.encode())
        if data == "ascii":
            for r in rows:
                f.write((" ".join("%.9g" % v for v in r) + "\n").encode())
        elif data == "binary":
            f.write(rows.tobytes())
        else:  # binary_compressed: SoA payload, LZF stream made of literal runs plus one back reference
            soa = np.ascontiguousarray(rows.T).tobytes()
            out = bytearray()
            i = 0
            while i < len(soa):
                # emit a back reference when the next 8 bytes repeat the previous 8 (exercises that branch)
                if i >= 8 and soa[i:i + 8] == soa[i - 8:i] and len(soa) - i >= 8:
                    ln, off = 8 - 2, 8 - 1
                    out += bytes([(ln << 5) | (off >> 8), off & 0xFF])
                    i += 8
                    continue
                run = soa[i:i + 32]
                out += bytes([len(run) - 1]) + run
                i += len(run)
            f.write(struct.pack("<II", len(out), len(soa)))
            f.write(bytes(out))


def dump(tools, path):
    out = subprocess.check_output([tools.PCD_TOOL, "dump", path], text=True).splitlines()
    head = out[0].split()
    pts = np.array([[float(v) for v in l.split()] for l in out[1:]], np.float32).reshape(-1, 3)
    return int(head[0]), [float(v) for v in head[4:7]], pts


@pytest.mark.parametrize("data", ["ascii", "binary", "binary_compressed"])
@pytest.mark.parametrize("extra", [False, True])
def test_pcd_reader_roundtrip(tools, tmp_path, views, data, extra):
    xyz = views["cheff001"][:3000].copy()
    xyz[7] = xyz[6]                      # a repeated point -> repeated bytes for the LZF back-reference branch
    p = str(tmp_path / "c.pcd")
    write_pcd(p, xyz, data, extra, viewpoint=(1.5, -2, 3))
    n, vp, pts = dump(tools, p)
    assert n == len(xyz) and vp == [1.5, -2.0, 3.0]
    assert np.array_equal(pts.view(np.uint32), xyz.view(np.uint32))


def test_pcd_nan_and_writers(tools, tmp_path, views):
    xyz = views["cheff002"][:500].copy()
    xyz[3, 1] = np.nan
    p = str(tmp_path / "n.pcd")
    write_pcd(p, xyz, "ascii")
    out = subprocess.check_output([tools.PCD_TOOL, "dump", p], text=True).splitlines()
    assert out[0].split()[3] == "0"                   # is_dense false
    # ascii -> binary -> ascii conversions keep every bit (%.8g is NOT enough for float32, so compare via binary)
    b, a = str(tmp_path / "b.pcd"), str(tmp_path / "a.pcd")
    good = views["cheff002"][:500]
    write_pcd(p, good, "binary")
    assert subprocess.call([tools.PCD_TOOL, "binary", p, b]) == 0
    _, _, pts = dump(tools, b)
    assert np.array_equal(pts.view(np.uint32), good.view(np.uint32))
    assert subprocess.call([tools.PCD_TOOL, "ascii", p, a]) == 0
    txt = open(a).read()
    assert "FIELDS x y z\n" in txt and "DATA ascii\n" in txt and "POINTS 500\n" in txt
    _, _, pts = dump(tools, a)
    assert np.allclose(pts, good, rtol=1e-7, atol=0)   # precision 8, like pcl::io::savePCDFileASCII
    assert subprocess.call([tools.PCD_TOOL, "dump", str(tmp_path / "missing.pcd")], stderr=subprocess.DEVNULL) == 1


@pytest.mark.gpu
def test_uniform_sampling_through_the_shim(tools, tmp_path, views, oracle):
    """pcl::UniformSampling in the shim forwards to kpl_uniform_sample (device); checked against the FP32
    numpy restatement of PCL's filter."""
    xyz = views["cheff000"][:8000]
    p, o = str(tmp_path / "c.pcd"), str(tmp_path / "s.pcd")
    write_pcd(p, xyz, "binary")
    leaf = 2.0
    assert subprocess.call([tools.PCD_TOOL, "subsample", str(leaf), p, o]) == 0
    _, _, pts = dump(tools, o)
    keep = oracle.uniform_sample(xyz, leaf)
    assert len(pts) == len(keep)
    assert np.array_equal(pts.view(np.uint32), np.ascontiguousarray(xyz[keep]).view(np.uint32))


def test_cli_argument_handling(tools, tmp_path):
    td = tools.TEST_DETECTOR
    r = subprocess.run([td, "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "--radiusFeatures arg (=20)" in r.stdout and "--pathRF" in r.stdout
    r = subprocess.run([td, "--bogus", "1"], capture_output=True, text=True)
    assert r.returncode == 0 and "unrecognised option" in r.stderr           # parse error -> usage, exit 0 (reference :112-113)
    r = subprocess.run([td, "--subSampling"], capture_output=True, text=True)
    assert r.returncode == 0 and "Subsampling needs leaf." in r.stdout
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    r = subprocess.run([td, "--pathRF", str(tmp_path / "nope.yaml.gz")], capture_output=True, text=True)
    assert r.returncode != 0                                                  # forest cannot be loaded -> -1 (reference :136-139)
    if not has_gpu:
        assert "no usable sm_100 device" in r.stderr                          # and never a CPU fallback


@pytest.mark.gpu
@pytest.mark.parametrize("view,data", [("cheff001", "ascii"), ("cheff000", "binary")])
def test_cli_end_to_end_matches_golden(tools, tmp_path, views, golden, view, data):
    """TestDetector on a bundled view with the reference's default arguments == golden keypoints."""
    cloud, kp = str(tmp_path / "cloud.pcd"), str(tmp_path / "kp.pcd")
    write_pcd(cloud, views[view], data)
    forest = os.path.join(ROOT, "tests", "golden", "forests", "synthetic-T100-D15.yaml.gz")
    r = subprocess.run([tools.TEST_DETECTOR, "--pathCloud", cloud, "--pathRF", forest, "--pathKP", kp, "--stats"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for line in ("Detector created.", "Point cloud loaded", "Normals Computed", "Keypoint computed", "DONE"):
        assert line in r.stdout
    g = golden[view]
    lines = open(kp).read().splitlines()
    assert lines[2] == "FIELDS x y z intensity" and lines[9] == "POINTS %d" % len(g["keypoints"])
    rows = np.array([[float(v) for v in l.split()] for l in lines[11:]], np.float32)
    ref = np.concatenate([views[view][g["keypoints"]], g["scores"][g["keypoints"]][:, None]], axis=1)
    assert rows.shape == ref.shape and np.allclose(rows, ref, rtol=1e-7, atol=0)
    # flags: --flipNormals / explicit radii / threshold / annuli / bins are accepted and change the result deterministically
    r2 = subprocess.run([tools.TEST_DETECTOR, "--pathCloud", cloud, "--pathRF", forest, "--pathKP", kp, "--flipNormals", "--radiusNMS=6",
                         "-t", "0.8", "--radiusFeatures", "20", "--annuli", "5", "--bins", "10"], capture_output=True, text=True)
    assert r2.returncode == 0 and "Flipping" in r2.stdout
    r3 = subprocess.run([tools.TEST_DETECTOR, "--pathCloud", cloud, "--pathRF", forest, "--annuli", "4", "--bins", "8"], capture_output=True, text=True)
    assert r3.returncode == 0 and "annuli*bins does not match" in r3.stderr   # var_count mismatch is reported, output empty


@pytest.mark.gpu
def test_cli_gpus_option_matches_single_gpu(tools, tmp_path, views, golden):
    """TestDetector --gpus N (x slabs through kpl_shard_*; with fewer devices than ranks the ranks share the devices as an
    in-process group) writes the keypoints of the single-GPU run."""
    cloud = str(tmp_path / "cloud.pcd")
    xyz = np.ascontiguousarray(views["cheff001"][:, [1, 0, 2]])         # longest axis onto the slab axis
    write_pcd(cloud, xyz, "binary")
    forest = os.path.join(ROOT, "tests", "golden", "forests", "synthetic-T100-D15.yaml.gz")
    out = {}
    for gpus in (1, 3):
        kp = str(tmp_path / ("kp%d.pcd" % gpus))
        r = subprocess.run([tools.TEST_DETECTOR, "--pathCloud", cloud, "--pathRF", forest, "--pathKP", kp, "--stats", "--gpus", str(gpus)],
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "Keypoint computed" in r.stdout and "DONE" in r.stdout
        if gpus > 1:
            assert "rank 2:" in r.stdout and "slabs 3" in r.stdout
        out[gpus] = open(kp).read().splitlines()[11:]
    assert len(out[1]) == len(golden["cheff001"]["keypoints"])
    assert out[1] == out[3]


@pytest.mark.gpu
def test_non_dense_clouds_are_compacted(kpl, views, golden):
    """NaN points (non-dense clouds) are skipped like the reference's kd-tree / runForest do (hpp:277): the facades compact
    the finite points, detect, and map indices and scores back."""
    import keypoint_learning_b200 as K
    xyz = views["cheff001"]
    holes = np.arange(5, len(xyz), 997)
    dirty = np.insert(xyz, holes, np.nan, axis=0).astype(np.float32)
    keep = np.nonzero(np.isfinite(dirty).all(axis=1))[0]
    assert len(keep) == len(xyz) and len(dirty) > len(xyz)
    det = K.KeypointLearningDetector()
    det.setNAnnulus(5); det.setNBins(10); det.setNonMaxima(True); det.setNonMaxRadius(4.0); det.setNonMaximaDrawsRemove(False)
    det.setPredictionThreshold(float(np.float32(0.85))); det.setRadiusSearch(20.0); det.setNormalsMode(1, k=10)
    assert det.loadForest(os.path.join(ROOT, "tests", "golden", "forests", "synthetic-T100-D15.yaml.gz"))
    det.setInputCloud(dirty)
    kp, idx = det.compute()
    g = golden["cheff001"]
    assert np.array_equal(idx, keep[g["keypoints"]])
    sc = det.getResponse()
    assert np.all(np.isnan(sc[~np.isfinite(dirty).all(axis=1)]))
    assert np.array_equal(sc[keep].view(np.uint32), g["scores"].view(np.uint32))
    nrm = det.computeNormals(dirty)
    assert np.all(np.isnan(nrm[~np.isfinite(dirty).all(axis=1)]))
    assert np.array_equal(nrm[keep].view(np.uint32), g["normals"].view(np.uint32)) if "normals" in g.files else True
    det.close()


@pytest.mark.gpu
def test_facade_estimates_missing_normals_like_initcompute(tools, tmp_path, oracle):
    """The C++ detector without setNormals() (pcd_tool detect): an ORGANIZED PCD takes the IntegralImageNormalEstimation
    branch of initCompute (hpp:138-145), an unorganized one the radius-mode NormalEstimation (hpp:130-137)."""
    from keypoint_learning_b200 import synth
    forest_file = os.path.join(ROOT, "tests", "golden", "forests", "synthetic-T100-D15.yaml.gz")
    forest = oracle.load_forest_yaml(forest_file)
    xyz, vp = synth.organized_range_image(240, 180, seed=5)
    h, w = xyz.shape[:2]
    flat = np.ascontiguousarray(xyz.reshape(-1, 3))
    p = str(tmp_path / "organized.pcd")
    hdr = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH %d\nHEIGHT %d\n"
           "VIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA binary\n" % (w, h, w * h))
    with open(p, "wb") as f:
        f.write(hdr.encode()); f.write(flat.tobytes())
    r = subprocess.run([tools.PCD_TOOL, "detect", forest_file, p, "20", "4", "0.5"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.splitlines()
    assert "Computing normals for KPL" in lines[0]
    nkp = int(lines[1])
    got = np.array([[float(v) for v in l.split()] for l in lines[2:2 + nkp]])
    nrm = oracle.normals_integral_image(xyz, 5.0, vp)
    keep = np.nonzero(np.isfinite(flat).all(axis=1))[0]
    ref = oracle.detect(np.ascontiguousarray(flat[keep]), forest, 20.0, 4.0, 0.5, 5, 10, normals4=nrm[keep], order=1)
    assert nkp == len(ref["keypoints"]) and nkp > 0
    assert np.array_equal(got[:, 0].astype(np.int64), keep[ref["keypoints"]])
    assert np.allclose(got[:, 1], ref["scores"][ref["keypoints"]], rtol=1e-7, atol=0)
