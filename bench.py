#!/usr/bin/env python
"""bench.py -- points scored / second through the detection hot path (search + normals + features +
forest + NMS) on N B200s of one node.  Contract: see the task statement; one JSON line on rank 0.

  python bench.py --gpus 1 --steps 5 --warmup 3                 # our arm (CUDA, C ABI)
  python bench.py --impl reference --steps 1 --warmup 0         # the reference's CPU path (oracle port)
  torchrun --nproc-per-node N bench.py --gpus N ...             # slab-sharded, halo over NCCL

A "step" is one full detection pass over the workload cloud (BASELINE.json configs[3]: the synthetic
10 M-point scene; --workload view1m selects configs[2]).  `value` times steps whose input is already
resident in HBM; `e2e` times kpl_detect() on pinned HOST buffers (H2D + compute + D2H each step).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
FOREST = os.path.join(ROOT, "tests", "golden", "forests", "synthetic-T100-D15.yaml.gz")
R_FEAT, R_NMS, TH, A, B, K_NORMALS = 20.0, 4.0, 0.85, 5, 10, 10


def measured_traffic(workload):
    """DRAM bytes per feature-kernel launch from the committed `ncu --set full` capture of the same workload
    (profiles/traffic.json, written by hand from profiles/*_ncu.txt); None when no capture exists."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get(workload)
    return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


VIEW_W, VIEW_H = 560, 360            # BASELINE.json configs[4]: ~200 k-point 2.5D views


def make_workload(name, n_points):
    from keypoint_learning_b200 import synth
    if name == "views":
        raise ValueError("the views workload is built per rank (make_views)")
    if name == "view1m":
        xyz, vp = synth.view_25d(1250, 800, seed=1234)
        desc = "synthetic 1M-point 2.5D view (1250x800 range image, pitch 0.64 mm)"
    else:
        xyz, vp = synth.scene_closed_surfaces(n_points, seed=4321)
        desc = "synthetic %d-point scene (32 closed surfaces, 2x2x1 m, ~0.64 mm spacing)" % len(xyz)
    return xyz, vp, desc


def make_views(rank, distinct):
    """`distinct` different 560x360 views for this rank (seeds rank*distinct + i), as (n,4) float32."""
    from keypoint_learning_b200 import synth
    out, vp = [], None
    for i in range(distinct):
        xyz, vp = synth.view_25d(VIEW_W, VIEW_H, seed=rank * distinct + i)
        x4 = np.ones((len(xyz), 4), np.float32); x4[:, :3] = xyz
        out.append(x4)
    return out, vp


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([t.strip() for t in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline_run(xyz, vp, sample_points, threads=None):
    """The oracle port of the reference's CPU path on a bounded spatial crop of the workload."""
    from oracle import oracle as O
    forest = O.load_forest_yaml(FOREST)
    # crop: the sample_points points nearest (in x, then y) to the median point -> same density as the workload
    c = np.median(xyz, axis=0)
    d = np.abs(xyz - c).max(axis=1)
    sel = np.argpartition(d, min(sample_points, len(xyz) - 1))[:sample_points]
    crop = np.ascontiguousarray(xyz[np.sort(sel)])
    t0 = time.perf_counter()
    res = O.detect(crop, forest, R_FEAT, R_NMS, TH, A, B, normals_mode=1, k=K_NORMALS, viewpoint=vp, order=2, threads=threads)
    dt = time.perf_counter() - t0
    return dict(value=len(crop) / dt, seconds=dt, cores=O.num_threads(), n=len(crop), stage_ms=res["stage_ms"],
                sample="cube crop of %d points around the median point of the workload, full pipeline, one pass" % len(crop))


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The reference binary
    cannot be built (PCL/FLANN/Eigen/Boost/OpenCV C++ absent), so this is the oracle port, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload == "views":
        hv, vp = make_views(0, 1)
        xyz = hv[0][:, :3].copy()
        desc = "batch of synthetic 2.5D views (%dx%d px, %d points each)" % (VIEW_W, VIEW_H, len(xyz))
    else:
        xyz, vp, desc = make_workload(args.workload, args.points)
    vals = []
    last = None
    for s in range(args.warmup + args.steps):
        last = cpu_baseline_run(xyz, vp, args.cpu_sample)
        if s >= args.warmup:
            vals.append(last)
    tot_pts = sum(v["n"] for v in vals); tot_s = sum(v["seconds"] for v in vals)
    value = tot_pts / tot_s
    line = {"impl": "reference", "metric": "points scored/sec (search+normals+features+RF+NMS)", "value": value, "unit": "points/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / max(1, len(vals)),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "radiusFeatures": R_FEAT, "radiusNMS": R_NMS, "threshold": TH, "annuli": A, "bins": B,
                       "forest": os.path.basename(FOREST), "normals": "kNN-%d" % K_NORMALS},
            "cpu_baseline": {"value": value, "unit": "points/s", "cores": last["cores"], "kind": "port", "sample": last["sample"],
                             "stage_ms": last["stage_ms"]},
            "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_b200(args):
    import torch
    import torch.distributed as dist
    import keypoint_learning_b200 as K

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    peak_gbs, peak_src = measured_peaks()

    views = args.workload == "views"
    if views:
        host_views, vp = make_views(rank, args.distinct_views)
        pts_view = len(host_views[0])
        n_total = pts_view * args.views_per_gpu * world          # points scored per step by the whole job
        desc = ("batch of synthetic 2.5D views (%dx%d px, %d points each): %d views per GPU per step, %d distinct per GPU, "
                "independent units round-robin over the GPUs, no communication (the full config is 4096 views)"
                % (VIEW_W, VIEW_H, pts_view, args.views_per_gpu, args.distinct_views))
        xyz = None
    else:
        xyz, vp, desc = make_workload(args.workload, args.points)
        n_total = len(xyz)

    det = K.KeypointLearningDetector(device=local)
    det.setNAnnulus(A); det.setNBins(B); det.setNonMaxima(True); det.setNonMaxRadius(R_NMS); det.setNonMaximaDrawsRemove(False)
    det.setPredictionThreshold(float(np.float32(TH))); det.setRadiusSearch(R_FEAT)
    det.setNormalsMode(1, k=K_NORMALS, viewpoint=vp)
    det.setCellsPerRadius(args.cpr)
    if not det.loadForest(FOREST):
        raise RuntimeError("forest failed to load")
    stream = torch.cuda.current_stream(dev)
    det.setStream(stream.cuda_stream)

    acc = {"feat_ms": 0.0, "pairs": 0, "scored": 0, "cand": 0, "launches": 0, "stage": None}

    def account(d_=None):
        d_ = d_ or det
        t = d_.timings(); st_ = d_.stats()
        acc["feat_ms"] += t["features_ms"]; acc["pairs"] += st_["feature_pairs"]; acc["scored"] += st_["n_scored"]
        acc["cand"] += st_["candidate_pairs"]; acc["launches"] += st_["kernel_launches"]
        if acc["stage"] is None:
            acc["stage"] = dict(t)
        else:
            for k_ in t:
                acc["stage"][k_] += t[k_]

    if views:
        pinned = [torch.from_numpy(v).pin_memory() for v in host_views]
        d_views = [p_.to(dev, non_blocking=True) for p_ in pinned]
        d_scores = torch.empty(pts_view, dtype=torch.float32, device=dev)
        d_kp = torch.empty(pts_view, dtype=torch.int32, device=dev)
        torch.cuda.synchronize(dev)

        # One 200 k-point view is ~1.5 waves of the feature kernel: a lone view leaves SMs idle in its tail.
        # Views are independent, so S detector contexts on S streams (one host thread each) keep S views in
        # flight; kernels of different views then fill each other's tails.  Context 0 is `det`.
        S = max(1, args.view_streams)
        dets = [det]
        for _ in range(1, S):
            d2_ = K.KeypointLearningDetector(device=local)
            d2_.setNAnnulus(A); d2_.setNBins(B); d2_.setNonMaxima(True); d2_.setNonMaxRadius(R_NMS); d2_.setNonMaximaDrawsRemove(False)
            d2_.setPredictionThreshold(float(np.float32(TH))); d2_.setRadiusSearch(R_FEAT)
            d2_.setNormalsMode(1, k=K_NORMALS, viewpoint=vp); d2_.setCellsPerRadius(args.cpr)
            assert d2_.loadForest(FOREST)
            dets.append(d2_)
        if S > 1:
            for d_ in dets:
                d_.setStream(None)                  # each context on its own non-blocking stream
        outs = [(torch.empty(pts_view, dtype=torch.float32, device=dev), torch.empty(pts_view, dtype=torch.int32, device=dev)) for _ in range(S)]
        torch.cuda.synchronize(dev)
        acc_lock = threading.Lock()

        def run_views(t, res):
            d_, (sc_, kp_) = dets[t], outs[t]
            nk = 0
            for v in range(t, args.views_per_gpu, S):
                nk += d_.detectDevice(d_views[v % len(d_views)].data_ptr(), pts_view, d_scores=sc_.data_ptr(), d_kp_idx=kp_.data_ptr())
                with acc_lock:
                    account(d_)
            res[t] = nk

        def step_fn():
            res = [0] * S
            if S == 1:
                run_views(0, res)
            else:
                th = [threading.Thread(target=run_views, args=(t, res)) for t in range(S)]
                for x in th:
                    x.start()
                for x in th:
                    x.join()
            return sum(res)
        n_local = pts_view * args.views_per_gpu
    elif world > 1:
        from keypoint_learning_b200 import shard
        job = shard.SlabJob(xyz, R_FEAT, R_NMS, args.cpr, rank, world, dev)

        def step_fn():
            nk = job.step(det); account(); return nk
        n_local = job.n_owned
    else:
        xyz4 = np.ones((n_total, 4), np.float32); xyz4[:, :3] = xyz
        host_xyz4 = torch.from_numpy(xyz4).pin_memory()
        d_xyz4 = host_xyz4.to(dev, non_blocking=True)
        d_scores = torch.empty(n_total, dtype=torch.float32, device=dev)
        d_kp = torch.empty(n_total, dtype=torch.int32, device=dev)
        torch.cuda.synchronize(dev)

        def step_fn():
            nk = det.detectDevice(d_xyz4.data_ptr(), n_total, d_scores=d_scores.data_ptr(), d_kp_idx=d_kp.data_ptr()); account(); return nk
        n_local = n_total

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- value: inputs resident in HBM ---------------------------------------------------------
    for _ in range(args.warmup):
        step_fn()
    sampler = ClockSampler(local)
    for k_ in ("feat_ms", "pairs", "scored", "cand", "launches"):
        acc[k_] = 0
    acc["stage"] = None
    barrier()
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    nkp = 0
    for _ in range(args.steps):
        nkp = step_fn()
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        tms = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    ms_per_step = ms / args.steps
    value = n_total / (ms_per_step * 1e-3)
    st = det.stats()
    launches = acc["launches"]
    stage = {k_: v_ / args.steps for k_, v_ in acc["stage"].items()}          # per step (summed over the step's calls)
    snap = dict(acc)                                                          # the e2e leg below keeps accounting
    per_rank = None
    if world > 1 and not views:
        # device time, slab size and pair count of every rank for the last step: shows the load balance
        mine = torch.tensor([stage["total_ms"], float(job.last_slab_points), float(st["feature_pairs"])], device=dev, dtype=torch.float64)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = [{"device_ms": round(float(t[0]), 3), "slab_points": int(t[1]), "pairs": int(t[2])} for t in allr]

    # ---- e2e: host buffers through the public host API (H2D + compute + D2H every step) ----------
    e2e = None
    if views:
        # every view: H2D from pinned host memory, detection, scores + keypoint indices back to the host (kpl_detect)
        steps_e2e = max(1, min(args.steps, 3))
        host_np = [p_.numpy() for p_ in pinned]
        for d_ in dets:
            d_.setInputCloud(host_np[0]); d_.setNormals(None); d_.compute()      # warm the staging buffers
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        nkb = [0] * len(dets)

        def e2e_views(t):
            for v in range(t, args.views_per_gpu, len(dets)):
                dets[t].setInputCloud(host_np[v % len(host_np)])
                _, idx = dets[t].compute()
                nkb[t] += len(idx) * 4

        for _ in range(steps_e2e):
            th = [threading.Thread(target=e2e_views, args=(t,)) for t in range(len(dets))]
            for x in th:
                x.start()
            for x in th:
                x.join()
        nk_bytes = sum(nkb)
        e1.record(stream)
        barrier()
        wall = (time.perf_counter() - t0) / steps_e2e
        e2e_ms = max(e0.elapsed_time(e1) / steps_e2e, wall * 1e3)
        if world > 1:
            tms = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            e2e_ms = float(tms.item())
        e2e = {"value": n_total / (e2e_ms * 1e-3), "unit": "points/s", "h2d_bytes_per_step": int(n_total * 16),
               "d2h_bytes_per_step": int(n_total * 4 + world * nk_bytes // steps_e2e), "ms_per_step": e2e_ms,
               "api": "kpl_detect per view (host xyz, pinned; normals estimated on device; scores + keypoint indices copied back)"}
    elif world == 1:
        sc_host = torch.empty(n_total, dtype=torch.float32).pin_memory().numpy()     # pinned result buffers, as for the input
        kp_host = torch.empty(n_total, dtype=torch.int32).pin_memory().numpy()
        host_np = host_xyz4.numpy()
        det.setInputCloud(host_np); det.setNormals(None)
        det.compute(scores_out=sc_host, kp_out=kp_host)   # warm the staging buffers
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        e0.record(stream)
        steps_e2e = max(1, min(args.steps, 3))
        for _ in range(steps_e2e):
            _, idx = det.compute(scores_out=sc_host, kp_out=kp_host)
        e1.record(stream)
        torch.cuda.synchronize(dev)
        wall = (time.perf_counter() - t0) / steps_e2e
        e2e_ms = max(e0.elapsed_time(e1) / steps_e2e, wall * 1e3)
        e2e = {"value": n_total / (e2e_ms * 1e-3), "unit": "points/s", "h2d_bytes_per_step": int(n_total * 16),
               "d2h_bytes_per_step": int(n_total * 4 + len(idx) * 4 + 64), "ms_per_step": e2e_ms,
               "api": "kpl_detect (host xyz, pinned; normals estimated on device; scores + keypoint indices copied back)"}
        del sc_host, kp_host
    else:
        # every step: owned slab H2D from pinned host memory, halo exchange + detection, keypoints D2H on rank 0
        steps_e2e = max(1, min(args.steps, 3))
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        n_kp_host = 0
        for _ in range(steps_e2e):
            job.xyz4.copy_(job.host_xyz4, non_blocking=True)
            job.step(det)
            if rank == 0:
                n_kp_host = len(job.last_global_keypoints.cpu())
        e1.record(stream)
        barrier()
        tms = torch.tensor([e0.elapsed_time(e1) / steps_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        e2e_ms = float(tms.item())
        e2e = {"value": n_total / (e2e_ms * 1e-3), "unit": "points/s", "h2d_bytes_per_step": int(n_total * 16),
               "d2h_bytes_per_step": int(n_kp_host * 8), "ms_per_step": e2e_ms,
               "api": "SlabJob.step: owned slab H2D from pinned host, NCCL halo exchange, kpl_detect_device, global keypoint list to host"}

    # ---- roofline of the dominant kernel (feature_kernel), algorithmic bytes per SURVEY.md 8d -------
    # per launch = per kpl_detect call; this rank's launches of the timed region
    calls = args.steps * (args.views_per_gpu if views else 1)
    pairs_self = (snap["pairs"] + snap["scored"]) / calls        # K_f with the query itself included, per launch
    feat_bytes = 32.0 * pairs_self
    feat_s = snap["feat_ms"] / calls * 1e-3
    achieved = feat_bytes / feat_s / 1e9
    pts_launch = snap["scored"] / calls
    pipe_bytes = (32.0 * pairs_self + 16.0 * K_NORMALS * pts_launch + 200.0 * pts_launch) * (args.views_per_gpu if views else 1)
    roofline = {"bound": "hbm", "kernel": "feature_kernel", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
                "traffic": (measured_traffic(args.workload) or {}).get("dram_bytes_per_launch") if world == 1 else None,
                "traffic_source": (measured_traffic(args.workload) or {}).get("source") if world == 1 else None,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": feat_bytes,
                "kernel_ms": feat_s * 1e3, "launches_per_step": calls // args.steps, "pairs_per_s": snap["pairs"] / calls / feat_s,
                "candidate_tests_per_s": snap["cand"] / calls / feat_s,
                "acceptance": snap["pairs"] / max(1, snap["cand"]),
                "pipeline_frac_rank0": (pipe_bytes / (ms_per_step * 1e-3) / 1e9) / peak_gbs if world == 1 else None,
                "stage_ms": stage, "fast_math_selftest_passed": bool(st["fast_math"])}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            c = cpu_baseline_run(host_views[0][:, :3].copy() if views else xyz, vp, args.cpu_sample)
            cpu = {"value": c["value"], "unit": "points/s", "cores": c["cores"], "kind": "port", "sample": c["sample"], "stage_ms": c["stage_ms"]}
        line = {"metric": "points scored/sec (search+normals+features+RF+NMS)", "value": value, "unit": "points/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak" if views else "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": desc, "radiusFeatures": R_FEAT, "radiusNMS": R_NMS, "threshold": TH, "annuli": A, "bins": B,
                           "forest": os.path.basename(FOREST), "normals": "kNN-%d on device" % K_NORMALS, "cells_per_radius": args.cpr,
                           "view_streams": args.view_streams if views else None,
                           "parallelism": ("views-dp%d" % world) if views else ("slab%d+halo" % world if world > 1 else "single"),
                           "l2": "inputs and intermediates of a step (%.0f MB) exceed the 126 MB L2; no flush needed" % (n_total / world * (16 + 16 + 16 + 16 + 16) / 1e6)},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
                "keypoints": int(nkp), "near_threshold_points": int(st["n_near_threshold"]), "n_points": n_total, "points_per_rank": n_local, "per_rank": per_rank}
        print(json.dumps(line), flush=True)
    if views:
        for d_ in dets[1:]:
            d_.close()
    det.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="scene10m", choices=["scene10m", "view1m", "views"])
    ap.add_argument("--views-per-gpu", type=int, default=32, help="views workload: detections per GPU per step")
    ap.add_argument("--distinct-views", type=int, default=8, help="views workload: different views resident per GPU")
    ap.add_argument("--view-streams", type=int, default=3, help="views workload: detector contexts / streams per GPU")
    ap.add_argument("--points", type=int, default=10_000_000)
    ap.add_argument("--cpr", type=int, default=4, help="grid cells per radiusFeatures")
    ap.add_argument("--cpu-sample", type=int, default=1_500_000, help="points of the workload crop the CPU legs run on (~10-20 s of host work)")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
