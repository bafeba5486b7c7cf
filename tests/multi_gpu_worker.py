"""torchrun worker of tests/test_gpu_multi.py: slab-sharded detection over WORLD_SIZE GPUs through the C entry points
(kpl_shard_*: NCCL halo + score exchange inside libkpl_b200.so) on a bundled view; every rank writes its owned scores,
rank 0 the global keypoint list, for the parent test to compare with the single-GPU result.  torch.distributed is
used for ONE thing: handing the 128-byte NCCL id of rank 0 to the other ranks."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    out_path = sys.argv[1]
    view = sys.argv[2] if len(sys.argv) > 2 else "cheff001"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    import keypoint_learning_b200 as K
    from keypoint_learning_b200 import shard

    xyz = np.load(os.path.join(ROOT, "tests", "golden", "views", view + ".npz"))["xyz"]
    xyz = np.ascontiguousarray(xyz[:, [1, 0, 2]])          # longest axis (y) onto the slab axis
    det = K.KeypointLearningDetector(device=local)
    det.setNAnnulus(5); det.setNBins(10); det.setNonMaxima(True); det.setNonMaxRadius(4.0); det.setNonMaximaDrawsRemove(False)
    det.setPredictionThreshold(float(np.float32(0.85))); det.setRadiusSearch(20.0); det.setNormalsMode(1, k=10)
    assert det.loadForest(os.path.join(ROOT, "tests", "golden", "forests", "synthetic-T100-D15.yaml.gz"))
    ids = [shard.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    plan = shard.plan_slabs(xyz, 20.0, 4.0, 4, world)
    job = shard.SlabJob(det, xyz, plan, rank, ids[0])
    sc = np.empty(job.n_owned, np.float32)
    n1, kp1 = shard.detect_widening(job, xyz, 20.0, 4.0, 4, scores_out=sc)
    sc2 = np.empty(job.n_owned, np.float32)
    job.upload()
    n2, kp2 = job.detect(scores_out=sc2)                    # a second step must give the same answer
    assert n1 == n2 and np.array_equal(sc.view(np.uint32), sc2.view(np.uint32))
    info = job.info()
    st = det.stats()
    np.savez(out_path + ".rank%d.npz" % rank, gidx=job.gidx, scores=sc, halo_bytes=info["halo_bytes"], scored=st["n_scored"],
             syncs=st["host_syncs"])
    if rank == 0:
        assert np.array_equal(kp1, kp2)
        np.savez(out_path, keypoints=kp1)
    dist.barrier()
    job.close()
    det.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
