"""torchrun worker of tests/test_gpu_multi.py: slab-sharded detection over WORLD_SIZE GPUs (NCCL halo
exchange) on a bundled view; rank 0 writes the global keypoint list for the parent test to compare with
the golden single-GPU result."""
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
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import keypoint_learning_b200 as K
    from keypoint_learning_b200 import shard

    xyz = np.load(os.path.join(ROOT, "tests", "golden", "views", view + ".npz"))["xyz"]
    xyz = np.ascontiguousarray(xyz[:, [1, 0, 2]])          # longest axis (y) onto the slab axis
    det = K.KeypointLearningDetector(device=local)
    det.setNAnnulus(5); det.setNBins(10); det.setNonMaxima(True); det.setNonMaxRadius(4.0); det.setNonMaximaDrawsRemove(False)
    det.setPredictionThreshold(float(np.float32(0.85))); det.setRadiusSearch(20.0); det.setNormalsMode(1, k=10)
    assert det.loadForest(os.path.join(ROOT, "tests", "golden", "forests", "synthetic-T100-D15.yaml.gz"))
    det.setStream(torch.cuda.current_stream().cuda_stream)
    job = shard.SlabJob(xyz, 20.0, 4.0, 4, rank, world, torch.device("cuda", local))
    n1 = job.step(det)
    n2 = job.step(det)                                      # a second step must give the same answer
    owned = (job._slab[1] == 3)
    sc = job._scores[: job.last_slab_points][owned].cpu().numpy()
    gi = job._slab[2][owned].cpu().numpy()
    np.savez(out_path + ".rank%d.npz" % rank, gidx=gi, scores=sc, halo_bytes=getattr(job, "halo_bytes", 0))
    if rank == 0:
        assert n1 == n2
        np.savez(out_path, keypoints=job.last_global_keypoints.cpu().numpy())
    dist.barrier()
    det.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
