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


def _exchange(rank, world, to_left, to_right):
    """Strips to / from the two neighbours over the process group (what ncclSend / ncclRecv do in csrc/shard.cu)."""
    def sizes(t):
        return torch.tensor([0 if t is None else len(t)], dtype=torch.int64)
    reqs, got = [], {}
    for peer, payload in ((rank - 1, to_left), (rank + 1, to_right)):
        if 0 <= peer < world:
            n_in = torch.zeros(1, dtype=torch.int64)
            ops = [dist.P2POp(dist.isend, sizes(payload), peer), dist.P2POp(dist.irecv, n_in, peer)]
            for r in dist.batch_isend_irecv(ops):
                r.wait()
            buf = torch.empty((int(n_in),) + tuple(payload.shape[1:]), dtype=torch.float32)
            ops = [dist.P2POp(dist.isend, torch.from_numpy(np.ascontiguousarray(payload)), peer), dist.P2POp(dist.irecv, buf, peer)]
            for r in dist.batch_isend_irecv(ops):
                r.wait()
            got[peer] = buf.numpy()
    return got.get(rank - 1), got.get(rank + 1)


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
    plan = shard.plan_slabs(xyz, R_FEAT, R_NMS, 4, world)             # the C planner (host only)
    hs = shard.HostSlab(xyz, plan, rank)
    # 1. position strips
    from_l, from_r = _exchange(rank, world, *hs.position_strips())
    p, role = hs.assemble(from_l, from_r)
    # 2. the oracle in place of the device stages, with the GLOBAL canonical grid: normals for everything held,
    #    scores for the owned points only
    nrm = O.normals_knn(p, 10)
    q = np.nonzero(role == 3)[0].astype(np.int32)
    feat = O.features(p, nrm, R_FEAT, 5, 10, order=1, qidx=q, canon=(plan.origin, plan.cell, plan.dims))
    sc = np.full(len(p), np.nan, np.float32)
    sc[q] = O.scores(forest, feat, nrm[q])
    # 3. score strips
    sl, sr = hs.score_strips(sc)
    got_l, got_r = _exchange(rank, world, sl.reshape(-1, 1), sr.reshape(-1, 1))
    sc_all = hs.merge_scores(sc, None if got_l is None else got_l.ravel(), None if got_r is None else got_r.ravel())
    # 4. NMS over everything held, keypoints of the owned points, gathered on rank 0
    kp = hs.owned_keypoints(O.nms(p, sc_all, R_NMS, TH))
    lists = [None] * world if rank == 0 else None
    dist.gather_object(kp, lists, dst=0)
    glob = np.sort(np.concatenate(lists)) if rank == 0 else np.zeros(0, np.int64)
    # expected layout, straight from the full cloud
    cx = shard.cell_coords(xyz, plan.origin, plan.cell, 0)
    c0, c1 = int(plan.cuts[rank]), int(plan.cuts[rank + 1])
    exp_l = np.nonzero((cx >= c0 - plan.halo) & (cx < c0))[0] if rank > 0 else np.zeros(0, np.int64)
    exp_r = np.nonzero((cx >= c1) & (cx < c1 + plan.halo))[0] if rank < world - 1 else np.zeros(0, np.int64)
    exp = np.concatenate([exp_l, np.nonzero((cx >= c0) & (cx < c1))[0], exp_r])
    ok_assemble = np.array_equal(p, xyz[exp]) and hs.n_l == len(exp_l) and hs.n_r == len(exp_r)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), ok_assemble=ok_assemble, gidx=hs.gidx, scores=sc[q], keypoints=glob,
             halo_points=len(hs.sel_l) + len(hs.sel_r), scored=len(q))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_equals_unsharded_gloo(tmp_path, world, oracle, kpl):
    xyz = _cloud()
    forest = oracle.load_forest_yaml(os.path.join(ROOT, "tests", "golden", "forests", "synthetic-SHOT-like-T50-D10.yaml.gz"))
    nrm = oracle.normals_knn(xyz, 10)
    feat = oracle.features(xyz, nrm, R_FEAT, 5, 10, order=1)
    sc = oracle.scores(forest, feat, nrm)
    kp = oracle.nms(xyz, sc, R_NMS, TH)
    assert len(kp) > 5

    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    seen = np.zeros(len(xyz), bool)
    scored = 0
    for rank in range(world):
        d = np.load(os.path.join(str(tmp_path), "rank%d.npz" % rank))
        assert bool(d["ok_assemble"]), "rank %d assembled a different slab than the full cloud dictates" % rank
        assert not seen[d["gidx"]].any()
        seen[d["gidx"]] = True
        assert np.array_equal(d["scores"].view(np.uint32), sc[d["gidx"]].view(np.uint32))      # bit-identical scores
        assert int(d["halo_points"]) > 0
        scored += int(d["scored"])
        if rank == 0:
            assert np.array_equal(d["keypoints"], kp)                                          # same keypoint set
    assert seen.all()
    assert scored == len(xyz)                                                                   # every point scored exactly once


def test_plan_is_balanced_and_rejects_thin_slabs(kpl):
    sys.path.insert(0, ROOT)
    from keypoint_learning_b200 import shard
    xyz = _cloud()
    plan = shard.plan_slabs(xyz, R_FEAT, R_NMS, 4, 2)
    cx = shard.cell_coords(xyz, plan.origin, plan.cell, 0)
    n0 = int((cx < plan.cuts[1]).sum())
    assert abs(n0 - len(xyz) / 2) < 0.15 * len(xyz)
    assert plan.halo == plan.reach_feat + 1 == 5 and plan.reach_nms == 1
    assert abs(plan.cost[0] - plan.cost[1]) < 0.15 * plan.cost.mean()
    with pytest.raises(ValueError):
        shard.plan_slabs(xyz, R_FEAT, R_NMS, 4, 8)       # 24 cells cannot host 8 slabs of >= 5 cells
    # the partition covers the cloud exactly once, in ascending index
    parts = [shard.partition(plan, xyz, r) for r in range(2)]
    assert all(np.all(np.diff(p) > 0) for p in parts)
    assert np.array_equal(np.sort(np.concatenate(parts)), np.arange(len(xyz)))
    # the grid of the plan is the grid the oracle derives for the unsharded cloud
    from oracle import oracle as O
    org, cell, dims = O.canon_grid(xyz, R_FEAT, 4)
    assert np.array_equal(org, plan.origin) and cell == plan.cell and np.array_equal(dims, plan.dims)


def test_plan_properties_on_random_clouds(kpl):
    """Cuts are a strictly increasing cover of the cell columns, every slab is at least a halo wide, and the modelled
    per-rank cost is balanced to within about one cell column."""
    sys.path.insert(0, ROOT)
    from keypoint_learning_b200 import shard
    rng = np.random.default_rng(11)
    planned = 0
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
        planned += 1
        cuts = plan.cuts
        assert len(cuts) == world + 1 and cuts[0] == 0 and cuts[-1] == plan.dims[0]
        assert np.all(np.diff(cuts) >= plan.halo)
        owned = np.array([len(shard.partition(plan, xyz, r)) for r in range(world)])
        assert owned.sum() == len(xyz)
        # more ranks never raise the bottleneck (the minimum slab width can stop it from falling, e.g. a clump
        # narrower than one halo)
        if world > 2:
            fewer = shard.plan_slabs(xyz, R_FEAT, R_NMS, 4, world - 1)
            assert plan.cost.max() <= fewer.cost.max() * 1.0001
    assert planned >= 5
    # where the minimum width does not bind (a ramp of density along x) the modelled cost is balanced
    x = rng.triangular(-600, 600, 600, 120_000)
    xyz = np.stack([x, rng.uniform(-60, 60, len(x)), rng.uniform(-10, 10, len(x))], axis=1).astype(np.float32)
    for world in (2, 4, 8):
        plan = shard.plan_slabs(xyz, R_FEAT, R_NMS, 4, world)
        assert plan.cost.max() <= 1.08 * plan.cost.mean(), (world, plan.cost.tolist())
        owned = np.array([len(shard.partition(plan, xyz, r)) for r in range(world)])
        assert owned.max() > 1.3 * owned.min() or world == 2          # balanced on cost, not on point count
