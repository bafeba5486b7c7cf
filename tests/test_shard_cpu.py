"""Multi-rank host logic of the slab sharding on CPU: world_size-2/3 `gloo` process groups exchange the
halos for real (torch.distributed all_gather + all_to_all); the CPU oracle stands in for the CUDA detection call, so the
test proves that (a) every rank assembles exactly the slab it needs and (b) the union of the slabs'
owned results equals the unsharded result bit for bit (forced global grid => same canonical order)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R_FEAT, R_NMS, TH = 20.0, 4.0, 0.85


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _cloud():
    xyz = np.load(os.path.join(ROOT, "tests", "golden", "views", "cheff001.npz"))["xyz"]
    # the longest axis of this view is y: make it the slab axis (x); crop to keep the CPU oracle quick
    xyz = np.ascontiguousarray(xyz[:, [1, 0, 2]])
    return np.ascontiguousarray(xyz[(xyz[:, 0] > -60) & (xyz[:, 0] < 60) & (xyz[:, 1] < 0)])


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from keypoint_learning_b200 import shard
    from oracle import oracle as O
    O.lib().kplo_set_threads(2)
    xyz = _cloud()
    forest = O.load_forest_yaml(os.path.join(ROOT, "tests", "golden", "forests", "synthetic-SHOT-like-T50-D10.yaml.gz"))
    job = shard.SlabJob(xyz, R_FEAT, R_NMS, 4, rank, world, "cpu")
    xyz4, role, gidx = job.assemble()
    ref = shard.reference_slab(xyz, job.plan, rank)
    ok_assemble = (np.array_equal(xyz4.numpy(), ref["xyz4"]) and np.array_equal(role.numpy(), ref["role"])
                   and np.array_equal(gidx.numpy(), ref["gidx"]) and np.array_equal(job.local_dims, ref["local_dims"])
                   and np.array_equal(job.offset, ref["offset"]))
    # the oracle in place of kpl_detect_device, with the GLOBAL canonical grid
    p = xyz4.numpy()[:, :3].copy()
    r = role.numpy()
    nrm = O.normals_knn(p, 10)
    q = np.nonzero(r & 1)[0].astype(np.int32)
    feat = O.features(p, nrm, R_FEAT, 5, 10, order=1, qidx=q, canon=(job.plan.origin, job.plan.cell, job.plan.dims))
    sc = np.full(len(p), np.nan, np.float32)
    sc[q] = O.scores(forest, feat, nrm[q])
    kp = O.nms(p, sc, R_NMS, TH)
    kp = kp[r[kp] == 3]
    glob = job.finish(torch.from_numpy(kp.astype(np.int64)))
    owned = r == 3
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), ok_assemble=ok_assemble, gidx=gidx.numpy()[owned], scores=sc[owned],
             keypoints=glob.numpy() if glob is not None else np.zeros(0, np.int64), halo_bytes=getattr(job, "halo_bytes", 0))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_equals_unsharded_gloo(tmp_path, world, oracle):
    xyz = _cloud()
    forest = oracle.load_forest_yaml(os.path.join(ROOT, "tests", "golden", "forests", "synthetic-SHOT-like-T50-D10.yaml.gz"))
    nrm = oracle.normals_knn(xyz, 10)
    feat = oracle.features(xyz, nrm, R_FEAT, 5, 10, order=1)
    sc = oracle.scores(forest, feat, nrm)
    kp = oracle.nms(xyz, sc, R_NMS, TH)
    assert len(kp) > 5

    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    seen = np.zeros(len(xyz), bool)
    for rank in range(world):
        d = np.load(os.path.join(str(tmp_path), "rank%d.npz" % rank))
        assert bool(d["ok_assemble"]), "rank %d assembled a different slab than reference_slab()" % rank
        assert not seen[d["gidx"]].any()
        seen[d["gidx"]] = True
        assert np.array_equal(d["scores"].view(np.uint32), sc[d["gidx"]].view(np.uint32))      # bit-identical scores
        if world > 1:
            assert int(d["halo_bytes"]) > 0
        if rank == 0:
            assert np.array_equal(d["keypoints"], kp)                                          # same keypoint set
    assert seen.all()


def test_plan_is_balanced_and_rejects_thin_slabs():
    sys.path.insert(0, ROOT)
    from keypoint_learning_b200 import shard
    xyz = _cloud()
    plan = shard.plan_slabs(xyz, R_FEAT, R_NMS, 4, 2)
    cx = shard.cell_coords(xyz, plan.origin, plan.cell, 0)
    n0 = int((cx < plan.cuts[1]).sum())
    assert abs(n0 - len(xyz) / 2) < 0.15 * len(xyz)
    assert plan.halo == plan.reach_nms + plan.reach_feat + 1 == 6
    with pytest.raises(ValueError):
        shard.plan_slabs(xyz, R_FEAT, R_NMS, 4, 8)       # 24 cells cannot host 8 slabs of >= 6 cells


def test_plan_properties_on_random_clouds():
    """Cuts are a strictly increasing cover of the cell columns, every slab is at least a halo wide, and the
    modelled per-rank cost is balanced to within one cell column of points."""
    sys.path.insert(0, ROOT)
    from keypoint_learning_b200 import shard
    rng = np.random.default_rng(11)
    for trial in range(12):
        n = int(rng.integers(20_000, 60_000))
        # clumpy along x: a mixture of slabs of different density, like the closed-surface scene
        centres = rng.uniform(-400, 400, 6)
        x = np.concatenate([rng.normal(c, rng.uniform(20, 120), n // 6) for c in centres])
        xyz = np.stack([x, rng.uniform(-50, 50, len(x)), rng.uniform(-20, 20, len(x))], axis=1).astype(np.float32)
        world = int(rng.integers(2, 9))
        try:
            plan = shard.plan_slabs(xyz, R_FEAT, R_NMS, 4, world)
        except ValueError:
            continue                                            # too few columns for that many ranks: rejected, not mis-planned
        cuts = plan.cuts
        assert len(cuts) == world + 1 and cuts[0] == 0 and cuts[-1] == plan.dims[0]
        assert np.all(np.diff(cuts) >= plan.halo)
        cx = shard.cell_coords(xyz, plan.origin, plan.cell, 0)
        hist = np.bincount(cx, minlength=int(plan.dims[0]))
        scored = np.array([hist[max(c0 - plan.reach_nms, 0):c1 + plan.reach_nms].sum() for c0, c1 in zip(cuts[:-1], cuts[1:])])
        assert scored.max() - scored.min() <= 2.5 * hist.max() + 0.15 * scored.mean(), (trial, world, scored.tolist())
        owned = np.array([hist[c0:c1].sum() for c0, c1 in zip(cuts[:-1], cuts[1:])])
        assert owned.sum() == len(xyz)
