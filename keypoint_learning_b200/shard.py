"""Spatial-slab sharding of ONE large cloud over the GPUs of a node (BASELINE.json configs[3]).

The reference is single process; this is the new build's multi-GPU layer.  The cloud is cut along x
into `world` slabs at CELL boundaries of the global canonical grid (balanced point counts).  Every
step each rank
  1. exchanges halo strips with its left / right neighbour (torch.distributed P2P: NCCL over NVLink
     on GPUs, gloo in the CPU tests) -- the only data-path communication;
  2. runs the whole detection path on [left halo | owned | right halo] with the GLOBAL grid forced
     (origin / cell offset), so cell keys, hence the canonical accumulation order, hence every float,
     are identical to the single-GPU run;
  3. keeps the keypoints of the points it owns; rank 0 gathers the global index lists.
Halo width in cells = reach(radiusNMS) + reach(radiusFeatures) + normal_support_cells: a keypoint
decision of an owned point needs scores within r_nms, which need normals within r_feat, which need
the k nearest points of those (k-NN support, at most `normal_support_cells` cells, checked by the
N-GPU == 1-GPU bit-exactness tests).
Roles (include/kpl.h): owned = 3 (scored + NMS output), halo within reach(r_nms) of the owned range = 1
(scored only), outer halo = 0 (normals / neighbours only).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist

ROLE_HALO, ROLE_SCORE, ROLE_OWNED = 0, 1, 3


@dataclass
class SlabPlan:
    origin: np.ndarray      # 3 doubles, global grid origin
    cell: float
    dims: np.ndarray        # 3 int32, global grid dims
    cuts: np.ndarray        # world+1 cell-x cut positions
    halo: int               # halo width in cells
    reach_nms: int
    reach_feat: int


def canonical_cell(r_feat: float, cpr: int) -> float:
    return float(np.float32(r_feat)) * (1.0 + 2.0 ** -20) / float(cpr)


def reach(radius: float, cell: float) -> int:
    return int(math.floor(float(np.float32(radius)) * (1.0 + 2.0 ** -21) / cell)) + 1


def cell_coords(xyz: np.ndarray, origin: np.ndarray, cell: float, axis: int = 0) -> np.ndarray:
    return np.floor((xyz[:, axis].astype(np.float64) - origin[axis]) / cell).astype(np.int64)


def plan_slabs(xyz: np.ndarray, r_feat: float, r_nms: float, cpr: int, world: int, normal_support_cells: int = 1) -> SlabPlan:
    cell = canonical_cell(r_feat, cpr)
    lo = xyz.min(axis=0).astype(np.float64)
    hi = xyz.max(axis=0).astype(np.float64)
    dims = (np.floor((hi - lo) / cell) + 1).astype(np.int32)
    cx = cell_coords(xyz, lo, cell, 0)
    hist = np.bincount(cx, minlength=int(dims[0])).astype(np.float64)
    rn, rf = reach(r_nms, cell), reach(r_feat, cell)
    halo = rn + rf + normal_support_cells
    nx = int(dims[0])
    cum = np.concatenate([[0.0], np.cumsum(hist)])

    def pts(a, b):                                   # points in columns [a, b), clipped to the grid
        return cum[min(max(b, 0), nx)] - cum[min(max(a, 0), nx)]

    def cost(c0, c1):
        # what a rank owning [c0, c1) computes: features + forest for its columns and the reach_nms margin
        # (cost ~ 1 per point, the neighbour count is set by the sampling density, not by the slab), normals
        # and grid for everything including the halo (~7 % of a scored point each)
        return pts(c0 - rn, c1 + rn) + 0.07 * pts(c0 - halo, c1 + halo)

    def greedy(limit):
        cuts, c0 = [0], 0
        while c0 < nx and len(cuts) <= world:
            c1 = c0 + 1
            while c1 < nx and cost(c0, c1 + 1) <= limit:
                c1 += 1
            cuts.append(c1)
            c0 = c1
        return cuts if cuts[-1] == nx and len(cuts) - 1 <= world else None

    lo_t, hi_t = 0.0, cost(0, nx)
    for _ in range(50):                               # smallest per-rank cost bound that needs <= world slabs
        mid = 0.5 * (lo_t + hi_t)
        if greedy(mid) is None:
            lo_t = mid
        else:
            hi_t = mid
    cuts = greedy(hi_t)
    while len(cuts) - 1 < world:                      # fewer slabs than ranks: split the widest
        w = np.diff(cuts)
        k = int(np.argmax(w))
        if w[k] < 2:
            break
        cuts.insert(k + 1, cuts[k] + int(w[k]) // 2)
    if len(cuts) - 1 != world:
        raise ValueError("cannot cut %d cell columns into %d slabs" % (nx, world))
    plan = SlabPlan(lo, cell, dims, np.asarray(cuts, np.int64), halo, rn, rf)
    widths = np.diff(plan.cuts)
    if world > 1 and widths.min() < plan.halo:
        raise ValueError("slabs (%s cells) are thinner than the halo (%d cells): use fewer ranks" % (widths.tolist(), plan.halo))
    return plan


def reference_slab(xyz: np.ndarray, plan: SlabPlan, rank: int):
    """What rank `rank` must end up with after the halo exchange, computed directly from the full cloud
    (no communication).  Used by the tests as the check of exchange_halo()/assemble() and to emulate a
    sharded run on one GPU.  Returns dict(xyz4, role, gidx, local_dims, offset)."""
    world = len(plan.cuts) - 1
    cx = cell_coords(xyz, plan.origin, plan.cell, 0)
    c0, c1 = int(plan.cuts[rank]), int(plan.cuts[rank + 1])
    pieces = []
    if rank > 0:
        pieces.append(np.nonzero((cx >= c0 - plan.halo) & (cx < c0))[0])
    pieces.append(np.nonzero((cx >= c0) & (cx < c1))[0])
    if rank < world - 1:
        pieces.append(np.nonzero((cx >= c1) & (cx < c1 + plan.halo))[0])
    gidx = np.concatenate(pieces).astype(np.int64)
    xyz4 = np.ones((len(gidx), 4), np.float32)
    xyz4[:, :3] = xyz[gidx]
    c = cx[gidx]
    role = np.zeros(len(gidx), np.uint8)
    role[(c >= c0 - plan.reach_nms) & (c < c1 + plan.reach_nms)] = ROLE_SCORE
    role[(c >= c0) & (c < c1)] = ROLE_OWNED
    x0 = max(c0 - plan.halo, 0)
    x1 = min(c1 + plan.halo, int(plan.dims[0]))
    return dict(xyz4=xyz4, role=role, gidx=gidx, local_dims=np.array([x1 - x0, plan.dims[1], plan.dims[2]], np.int32),
                offset=np.array([x0, 0, 0], np.int32))


class SlabJob:
    """Per-rank state of a sharded detection job.  `device` may be a CUDA device (NCCL) or 'cpu' (gloo)."""

    def __init__(self, xyz: np.ndarray, r_feat: float, r_nms: float, cpr: int, rank: int, world: int, device, plan: SlabPlan | None = None):
        self.rank, self.world, self.device = rank, world, torch.device(device)
        self.plan = plan or plan_slabs(xyz, r_feat, r_nms, cpr, world)
        p = self.plan
        cx = cell_coords(xyz, p.origin, p.cell, 0)
        self.c0, self.c1 = int(p.cuts[rank]), int(p.cuts[rank + 1])
        own = np.nonzero((cx >= self.c0) & (cx < self.c1))[0]           # ascending global index
        self.n_owned = len(own)
        self.n_total = len(xyz)
        xyz4 = np.ones((len(own), 4), np.float32)
        xyz4[:, :3] = xyz[own]
        self.host_xyz4 = torch.from_numpy(xyz4)
        if self.device.type == "cuda":
            self.host_xyz4 = self.host_xyz4.pin_memory()
        self.xyz4 = self.host_xyz4.to(self.device)
        self.gidx = torch.from_numpy(own.astype(np.int64)).to(self.device)
        self.cx = torch.from_numpy(cx[own].astype(np.int32)).to(self.device)
        # the local grid: global origin, x offset so that local keys stay small
        self.x0 = max(self.c0 - p.halo, 0)
        self.x1 = min(self.c1 + p.halo, int(p.dims[0]))
        self.local_dims = np.array([self.x1 - self.x0, p.dims[1], p.dims[2]], np.int32)
        self.offset = np.array([self.x0, 0, 0], np.int32)
        self._scores = None
        self._kp = None
        # which of my points the neighbours need: a property of the resident slab, not of a step
        self._left_sel = torch.nonzero(self.cx < self.c0 + p.halo).flatten() if rank > 0 else None
        self._right_sel = torch.nonzero(self.cx >= self.c1 - p.halo).flatten() if rank < world - 1 else None

    # ---- step pieces (kept separate so the CPU tests can put the oracle in the middle) --------------
    def exchange_halo(self):
        """Send my boundary strips to the two neighbours, receive theirs (NCCL all_to_all over NVLink; gloo on
        CPU).  Returns the (left, right) received buffers, None at the ends."""
        p, r, w = self.plan, self.rank, self.world
        left_sel, right_sel = self._left_sel, self._right_sel

        def pack(sel):
            # one message per neighbour: xyz4 | gidx (as 2 x int32 bit patterns) | cx  -> float32 [m, 7]
            m = len(sel)
            buf = torch.empty((m, 7), dtype=torch.float32, device=self.device)
            buf[:, :4] = self.xyz4[sel]
            buf[:, 4:6] = self.gidx[sel].view(torch.int32).view(m, 2).view(torch.float32)
            buf[:, 6] = self.cx[sel].view(torch.float32)
            return buf

        send_l = pack(left_sel) if left_sel is not None else None
        send_r = pack(right_sel) if right_sel is not None else None
        # strip sizes of every rank (one tiny all_gather), then ONE all_to_all whose only non-empty
        # splits are the two neighbours: the same collective sequence on every rank, whatever its position.
        mine = torch.tensor([len(send_l) if send_l is not None else 0, len(send_r) if send_r is not None else 0],
                            dtype=torch.int64, device=self.device)
        sizes = [torch.zeros(2, dtype=torch.int64, device=self.device) for _ in range(w)]
        dist.all_gather(sizes, mine)
        sizes = torch.stack(sizes).cpu()
        in_split = [0] * w
        out_split = [0] * w
        if r > 0:
            in_split[r - 1] = int(sizes[r, 0]); out_split[r - 1] = int(sizes[r - 1, 1])
        if r < w - 1:
            in_split[r + 1] = int(sizes[r, 1]); out_split[r + 1] = int(sizes[r + 1, 0])
        send = torch.cat([t for t in (send_l, send_r) if t is not None]) if (send_l is not None or send_r is not None) \
            else torch.empty((0, 7), dtype=torch.float32, device=self.device)
        recv = torch.empty((sum(out_split), 7), dtype=torch.float32, device=self.device)
        dist.all_to_all_single(recv, send.contiguous(), output_split_sizes=out_split, input_split_sizes=in_split)
        n_l = out_split[r - 1] if r > 0 else 0
        recv_l = recv[:n_l] if r > 0 else None
        recv_r = recv[n_l:] if r < w - 1 else None
        self.halo_bytes = sum(int(t.numel()) * 4 for t in (send_l, send_r) if t is not None)
        return recv_l, recv_r

    @staticmethod
    def _unpack(buf):
        m = buf.shape[0]
        xyz4 = buf[:, :4].contiguous()
        gidx = buf[:, 4:6].contiguous().view(torch.int32).view(m, 2).view(torch.int64).flatten()
        cx = buf[:, 6].contiguous().view(torch.int32)
        return xyz4, gidx, cx

    def assemble(self):
        """-> (xyz4 [n,4] float32, role [n] uint8, gidx [n] int64).  Pieces are whole cells and each is in
        ascending global index, so inside every cell the order is the global one."""
        recv_l, recv_r = self.exchange_halo() if self.world > 1 else (None, None)
        parts_xyz, parts_g, parts_cx = [], [], []
        for buf in (recv_l,):
            if buf is not None and len(buf):
                a, b, c = self._unpack(buf); parts_xyz.append(a); parts_g.append(b); parts_cx.append(c)
        parts_xyz.append(self.xyz4); parts_g.append(self.gidx); parts_cx.append(self.cx)
        for buf in (recv_r,):
            if buf is not None and len(buf):
                a, b, c = self._unpack(buf); parts_xyz.append(a); parts_g.append(b); parts_cx.append(c)
        xyz4 = torch.cat(parts_xyz) if len(parts_xyz) > 1 else parts_xyz[0]
        gidx = torch.cat(parts_g) if len(parts_g) > 1 else parts_g[0]
        cx = torch.cat(parts_cx) if len(parts_cx) > 1 else parts_cx[0]
        rn = self.plan.reach_nms
        role = torch.zeros(len(cx), dtype=torch.uint8, device=self.device)
        role[(cx >= self.c0 - rn) & (cx < self.c1 + rn)] = ROLE_SCORE
        role[(cx >= self.c0) & (cx < self.c1)] = ROLE_OWNED
        self._slab = (xyz4, role, gidx)
        return xyz4, role, gidx

    def finish(self, kp_local: torch.Tensor):
        """kp_local: local indices (into the assembled slab) of this rank's keypoints.  Rank 0 returns the
        ascending global keypoint index list, the other ranks None."""
        _, _, gidx = self._slab
        mine = gidx[kp_local.long()]
        if self.world == 1:
            return torch.sort(mine).values
        cnt = torch.tensor([len(mine)], dtype=torch.int64, device=self.device)
        cnts = [torch.zeros(1, dtype=torch.int64, device=self.device) for _ in range(self.world)]
        dist.all_gather(cnts, cnt)
        cnts = torch.cat(cnts).cpu().tolist()               # one host sync for all ranks' counts
        mx = max(cnts)
        pad = torch.full((max(mx, 1),), -1, dtype=torch.int64, device=self.device)
        pad[:len(mine)] = mine
        allp = [torch.empty_like(pad) for _ in range(self.world)]
        dist.all_gather(allp, pad)
        if self.rank != 0:
            return None
        out = torch.cat([a[:c] for a, c in zip(allp, cnts)])
        return torch.sort(out).values

    # ---- the GPU step --------------------------------------------------------------------------------
    def step(self, det):
        """One full sharded detection step on the GPU.  Returns the global keypoint count on rank 0."""
        xyz4, role, gidx = self.assemble()
        n = xyz4.shape[0]
        if self._kp is None or self._kp.numel() < n:
            self._kp = torch.empty(n + n // 8, dtype=torch.int32, device=self.device)
            self._scores = torch.empty(n + n // 8, dtype=torch.float32, device=self.device)
        det.setForcedGrid(self.plan.origin, self.local_dims, self.offset)
        if self.device.type == "cuda":
            # the slab was assembled by torch ops on the current stream: run the detection on that stream too
            det.setStream(torch.cuda.current_stream(self.device).cuda_stream)
        nkp = det.detectDevice(xyz4.data_ptr(), n, d_role=role.data_ptr(), d_scores=self._scores.data_ptr(), d_kp_idx=self._kp.data_ptr())
        glob = self.finish(self._kp[:nkp])
        self.last_global_keypoints = glob
        self.last_slab_points = n
        return int(len(glob)) if glob is not None else 0
