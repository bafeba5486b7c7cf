"""Issue-slot model of the feature kernel for different ways of forming its warps (experiments tool, CPU only).

Takes a dense crop of the bench scene, builds the canonical grid, forms warps of 32 queries by
  runs    : 32 consecutive sorted points of one run of a cell row (the shipped work list)
  morton  : 32 consecutive points in Morton order of a sub-cell lattice (cell / SUB), groups = lanes within one cell of the leader
and replays the kernel's tile loop for a sample of warps: candidates tested, votes, vote iterations
(ceil(max popc / 2) per tile, 16 for interior tiles).  Prints lane efficiency of the vote phase, acceptance and a
relative issue-slot estimate.
"""
import sys
import numpy as np

sys.path.insert(0, ".")
from keypoint_learning_b200 import synth  # noqa: E402

R = 20.0
CPR = 4
CELL = R * (1 + 2.0 ** -20) / CPR
R2 = np.float32(R * R)
RCULL2 = R * R * (1 + 1e-5)


def morton3(ix, iy, iz):
    def spread(v):
        v = v.astype(np.uint64)
        out = np.zeros_like(v)
        for b in range(21):
            out |= ((v >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b)
        return out
    return spread(ix) | (spread(iy) << np.uint64(1)) | (spread(iz) << np.uint64(2))


def hilbert3(ix, iy, iz, bits):
    """Hilbert index (Skilling's transform), vectorised."""
    X = [ix.astype(np.uint64).copy(), iy.astype(np.uint64).copy(), iz.astype(np.uint64).copy()]
    M = np.uint64(1) << np.uint64(bits - 1)
    Q = M
    while Q > 1:
        P = Q - np.uint64(1)
        for i in range(3):
            sel = (X[i] & Q) != 0
            X[0] = np.where(sel, X[0] ^ P, X[0])
            t = (X[0] ^ X[i]) & P
            t = np.where(sel, np.uint64(0), t)
            X[0] ^= t
            X[i] ^= t
        Q >>= np.uint64(1)
    for i in range(1, 3):
        X[i] ^= X[i - 1]
    t = np.zeros_like(X[0])
    Q = M
    while Q > 1:
        t = np.where((X[2] & Q) != 0, t ^ (Q - np.uint64(1)), t)
        Q >>= np.uint64(1)
    for i in range(3):
        X[i] ^= t
    out = np.zeros_like(X[0])
    for b in range(bits - 1, -1, -1):
        for i in range(3):
            out = (out << np.uint64(1)) | ((X[i] >> np.uint64(b)) & np.uint64(1))
    return out


def main():
    n_crop = int(sys.argv[1]) if len(sys.argv) > 1 else 600_000
    nsample = int(sys.argv[2]) if len(sys.argv) > 2 else 150
    xyz, _ = synth.scene_closed_surfaces()
    xyz = synth.cube_crop(xyz, n_crop).astype(np.float32)
    n = len(xyz)
    lo = xyz.min(0).astype(np.float64)
    c = np.floor((xyz.astype(np.float64) - lo) / CELL).astype(np.int64)
    dim = c.max(0) + 1
    key = (c[:, 2] * dim[1] + c[:, 1]) * dim[0] + c[:, 0]
    order = np.argsort(key, kind="stable")
    sp = xyz[order]
    sk = key[order]
    sc = c[order]
    ncells = int(dim.prod())
    cell_start = np.searchsorted(sk, np.arange(ncells + 1))
    interior = np.all((sp > xyz.min(0) + 2 * R) & (sp < xyz.max(0) - 2 * R), axis=1)
    print(f"n={n} dim={dim} interior={interior.sum()}")
    rng = np.random.default_rng(0)

    def rows_for_group(qs, mode):
        """list of (start, end) candidate ranges in ascending (z, y) for the group of sorted positions qs"""
        cc = sc[qs]
        x0, y0, z0 = cc.min(0)
        x1, y1, z1 = cc.max(0)
        out = []
        if mode == "cells":
            for zz in range(max(z0 - CPR, 0), min(z1 + CPR, dim[2] - 1) + 1):
                for yy in range(max(y0 - CPR, 0), min(y1 + CPR, dim[1] - 1) + 1):
                    gy = max(max(y0 - yy, yy - y1) - 1, 0)
                    gz = max(max(z0 - zz, zz - z1) - 1, 0)
                    gap2 = (gy * gy + gz * gz) * CELL * CELL
                    if gap2 < RCULL2:
                        rx = min(int(np.sqrt(RCULL2 - gap2) / CELL) + 1, CPR)
                        xa, xb = max(x0 - rx, 0), min(x1 + rx, dim[0] - 1)
                        base = (zz * dim[1] + yy) * dim[0]
                        s, e = cell_start[base + xa], cell_start[base + xb + 1]
                        if e > s:
                            out.append((s, e))
        else:  # exact coordinates of the group's bounding box
            p = sp[qs].astype(np.float64) - lo
            pmin, pmax = p.min(0), p.max(0)
            for zz in range(max(z0 - CPR, 0), min(z1 + CPR, dim[2] - 1) + 1):
                gz = max(0.0, pmin[2] - (zz + 1) * CELL, zz * CELL - pmax[2])
                for yy in range(max(y0 - CPR, 0), min(y1 + CPR, dim[1] - 1) + 1):
                    gy = max(0.0, pmin[1] - (yy + 1) * CELL, yy * CELL - pmax[1])
                    gap2 = gy * gy + gz * gz
                    if gap2 < RCULL2:
                        rx = np.sqrt(RCULL2 - gap2)
                        xa = max(int(np.floor((pmin[0] - rx) / CELL - 1e-3)), 0)
                        xb = min(int(np.floor((pmax[0] + rx) / CELL + 1e-3)), dim[0] - 1)
                        base = (zz * dim[1] + yy) * dim[0]
                        s, e = cell_start[base + xa], cell_start[base + xb + 1]
                        if e > s:
                            out.append((s, e))
        return out

    def replay(groups, mode):
        """groups: list of arrays of sorted positions processed together (one pass each).  Returns counters."""
        cand = pairs = iters = tiles = interior_tiles = 0
        for qs in groups:
            qp = sp[qs]
            for (s, e) in rows_for_group(qs, mode):
                for tb in range(s, e, 32):
                    te = min(tb + 32, e)
                    cp = sp[tb:te]
                    d = qp[:, None, :] - cp[None, :, :]
                    d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
                    m = d2 < R2
                    # self exclusion
                    inside = (qs >= tb) & (qs < te)
                    m[np.nonzero(inside)[0], qs[inside] - tb] = False
                    pc = m.sum(1)
                    tiles += 1
                    cand += te - tb
                    pairs += int(pc.sum())
                    if te - tb == 32 and pc.min() == 32:
                        interior_tiles += 1
                    iters += (int(pc.max()) + 1) // 2
        return np.array([cand, pairs, iters, tiles, interior_tiles], dtype=np.float64)

    def report(name, warps, mode):
        tot = np.zeros(5)
        nq = 0
        ngroups = 0
        for groups in warps:
            tot += replay(groups, mode)
            nq += sum(len(g) for g in groups)
            ngroups += len(groups)
        cand, pairs, iters, tiles, it = tot
        nw = len(warps)
        # issue slots (warp instructions): phase 1 7.75 per candidate (x lanes: one warp instruction serves 32 lanes),
        # votes 160 per iteration (two votes), tile overhead 45
        slots = cand * 7.75 + iters * 160 + tiles * 45
        print(f"{name:28s} warps={nw} fill={nq / (32 * nw):.3f} groups/warp={ngroups / nw:.2f} acceptance={pairs / (cand * nq / nw):.3f} "
              f"vote-lane-eff={pairs / (iters * 2 * 32):.3f} interior={it / tiles:.3f} slots/pair={slots / pairs * 32:.1f} "
              f"(p1 {cand * 7.75 / pairs * 32:.1f} vote {iters * 160 / pairs * 32:.1f} tile {tiles * 45 / pairs * 32:.1f})")
        return slots / pairs

    # ---- shipped work list: runs of a row, span CPR
    work = []
    nrows = int(dim[1] * dim[2])
    for row in range(nrows):
        cs = cell_start[row * dim[0]:(row + 1) * dim[0] + 1]
        run_x = -1
        run_s = run_e = 0
        def close():
            for s in range(run_s, run_e, 32):
                work.append(np.arange(s, min(s + 32, run_e)))
        for x in range(dim[0]):
            if cs[x + 1] > cs[x]:
                if run_x < 0 or x - run_x > CPR:
                    if run_x >= 0:
                        close()
                    run_x = x
                    run_s = cs[x]
                run_e = cs[x + 1]
        if run_x >= 0:
            close()
    ok = [w for w in work if interior[w].all()]
    print(f"runs: {len(work)} warps, fill {n / (32 * len(work)):.3f}; interior warps {len(ok)}")
    pick = rng.choice(len(ok), size=min(nsample, len(ok)), replace=False)
    base = report("runs / cell culling", [[ok[i]] for i in pick], "cells")
    report("runs / exact-box culling", [[ok[i]] for i in pick], "exact")

    # ---- sub-cell curve orders
    for curve in ("morton", "hilbert"):
        for sub in (2, 4, 8):
            f = np.floor((sp.astype(np.float64) - lo) / (CELL / sub)).astype(np.int64)
            bits = int(np.ceil(np.log2(f.max() + 1)))
            code = morton3(f[:, 0], f[:, 1], f[:, 2]) if curve == "morton" else hilbert3(f[:, 0], f[:, 1], f[:, 2], bits)
            perm = np.argsort(code, kind="stable")
            warps_all = [perm[i:i + 32] for i in range(0, n, 32)]
            ok2 = [w for w in warps_all if interior[w].all()]
            pick2 = rng.choice(len(ok2), size=min(nsample, len(ok2)), replace=False)
            for gspan in (1, 2):
                wl = []
                for i in pick2:
                    w = ok2[i]
                    rem = list(range(len(w)))
                    groups = []
                    while rem:
                        l = rem[0]
                        mem = [j for j in rem if np.all(np.abs(sc[w[j]] - sc[w[l]]) <= gspan)]
                        groups.append(np.sort(w[mem]))
                        rem = [j for j in rem if j not in mem]
                    wl.append(groups)
                v = report(f"{curve} sub={sub} gspan={gspan} exact", wl, "exact")
                print(f"    relative to shipped: {v / base:.3f}")


if __name__ == "__main__":
    main()
