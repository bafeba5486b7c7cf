"""N-GPU result == 1-GPU result, bit for bit, over real NCCL halo exchange.
Skipped when the box has a single GPU (the same logic runs on CPU with gloo in test_shard_cpu.py)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_nccl_equals_golden(tmp_path, golden, kpl, world):
    if _ngpus() < world:
        pytest.skip("needs %d GPUs" % world)
    out = str(tmp_path / "kp.npz")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multi_gpu_worker.py"), out, "cheff001"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    # the single-GPU, unsharded run of the same (axis-permuted) cloud is the reference
    xyz = np.load(os.path.join(ROOT, "tests", "golden", "views", "cheff001.npz"))["xyz"]
    xyz = np.ascontiguousarray(xyz[:, [1, 0, 2]])
    det = kpl.KeypointLearningDetector()
    det.setNAnnulus(5); det.setNBins(10); det.setNonMaxima(True); det.setNonMaxRadius(4.0); det.setNonMaximaDrawsRemove(False)
    det.setPredictionThreshold(float(np.float32(0.85))); det.setRadiusSearch(20.0); det.setNormalsMode(1, k=10)
    assert det.loadForest(os.path.join(ROOT, "tests", "golden", "forests", "synthetic-T100-D15.yaml.gz"))
    det.setInputCloud(xyz)
    _, idx_full = det.compute()
    sc_full = det.getResponse().copy()
    det.close()
    assert len(idx_full) == len(golden["cheff001"]["keypoints"]) or len(idx_full) > 0
    assert np.array_equal(np.load(out)["keypoints"], idx_full)
    seen = np.zeros(len(sc_full), bool)
    for rank in range(world):
        d = np.load(out + ".rank%d.npz" % rank)
        assert not seen[d["gidx"]].any()
        seen[d["gidx"]] = True
        assert np.array_equal(d["scores"].view(np.uint32), sc_full[d["gidx"]].view(np.uint32))
        assert int(d["halo_bytes"]) > 0
    assert seen.all()
