"""Executable model of the data-parallel DistributeOctTree formulation used by csrc/octree.cu.

The reference (ORBextractor.cc:539-763) walks a std::list of nodes.  The CUDA kernel instead
  1. gives every candidate a quadtree PATH CODE that depends on its coordinates only (node boundaries are a
     function of the level geometry, never of the data), sorts candidates by that code so that every node of
     every depth is a contiguous range, and
  2. replays the list algorithm in ROUNDS over those ranges, tracking for every node a creation stamp
     (round, 4*rank_of_parent + quadrant).  std::list::push_front makes list order == descending creation
     order, which is all the reference's control flow depends on.
This file is that formulation in numpy/python; tests/test_octree_model.py checks it against the C oracle
(which follows the reference line by line).  Test infrastructure only.
"""
import math

import numpy as np

f32 = np.float32


def _ceil_half(a):
    return int(math.ceil(float(f32(a) / f32(2))))


def path_luts(width, height):
    """per-coordinate path codes.  Returns (n_ini, depth, xcode[width+1], ycode[height+1]) where
    xcode[x] = (ini_bucket, bits...) packed as bucket << depth | bits (MSB = first split)."""
    n_ini = int(np.floor(f32(width) / f32(height) + f32(0.5)))  # roundf for positive values
    if n_ini < 1:
        raise ValueError("nIni == 0")
    hx = f32(width) / f32(n_ini)

    def axis_paths(n, ini_of, bounds_of, depth):
        codes = []
        for v in range(n):
            b = ini_of(v)
            lo, hi = bounds_of(b)
            bits = 0
            for _ in range(depth):
                mid = lo + _ceil_half(hi - lo)
                if f32(v) < f32(mid):
                    bits = bits << 1
                    hi = mid
                else:
                    bits = (bits << 1) | 1
                    lo = mid
            codes.append((b << depth) | bits)
        return codes

    def x_ini(x):
        return min(int(f32(x) / hx), n_ini - 1)

    def x_bounds(b):
        return int(hx * f32(b)), int(hx * f32(b + 1))

    depth = 1
    while True:
        xc = axis_paths(width + 1, x_ini, x_bounds, depth)
        yc = axis_paths(height + 1, lambda y: 0, lambda b: (0, height), depth)
        if len(set(xc)) == len(xc) and len(set(yc)) == len(yc):
            break
        depth += 1
    return n_ini, depth, xc, yc


def _spread(v, depth):
    r = 0
    for i in range(depth):
        r |= ((v >> i) & 1) << (2 * i)
    return r


def distribute(xs, ys, scores, order, width, height, N):
    """xs, ys: int coordinates relative to minX/minY; scores: int; order: original position of every candidate
    (tie-break of the best-response pick).  Returns indices into the input in the reference's output order."""
    n = len(xs)
    n_ini, D, xc, yc = path_luts(width, height)
    mask = (1 << D) - 1
    keys = [((xc[x] >> D) << (2 * D)) | _spread(xc[x] & mask, D) | (_spread(yc[y] & mask, D) << 1) for x, y in zip(xs, ys)]
    perm = sorted(range(n), key=lambda i: keys[i])
    skey = [keys[i] for i in perm]

    def lower_bound(lo, hi, t):
        while lo < hi:
            m = (lo + hi) // 2
            if skey[m] < t:
                lo = m + 1
            else:
                hi = m
        return lo

    def children(node):
        s, e, d, prefix = node["s"], node["e"], node["d"], node["p"]
        sh = 2 * (D - d - 1)
        assert sh >= 0, "node with >1 keys at full depth"
        cuts = [s] + [lower_bound(s, e, ((prefix << 2) | q) << sh) for q in (1, 2, 3)] + [e]
        return [dict(s=cuts[q], e=cuts[q + 1], d=d + 1, p=(prefix << 2) | q) for q in range(4)]

    fin = []  # nodes that stay in the list: (stamp_round, stamp_idx, node)
    act = []  # expandable nodes of the latest round, in PROCESSING order
    rnd = 0
    for b in range(n_ini):
        s = lower_bound(0, n, b << (2 * D))
        e = lower_bound(0, n, (b + 1) << (2 * D))
        nd = dict(s=s, e=e, d=0, p=b, r=0, i=n_ini - 1 - b)
        if e - s == 1:
            fin.append(nd)
        elif e - s > 1:
            act.append(nd)  # list order of the initial nodes is ascending b == descending stamp idx
    size = len(fin) + len(act)

    def expand(procs):
        """divide procs (already in processing order); returns (new_fin, new_act_in_creation_order, growth list)"""
        nf, na, growth = [], [], []
        for rho, nd in enumerate(procs):
            g = -1
            for q, ch in enumerate(children(nd)):
                c = ch["e"] - ch["s"]
                if c == 0:
                    continue
                g += 1
                ch["r"], ch["i"] = rnd, 4 * rho + q
                (nf if c == 1 else na).append(ch)
            growth.append(g)
        return nf, na, growth

    finish = False
    while not finish:
        rnd += 1
        prev = size
        nf, na, growth = expand(act)
        fin += nf
        act = na[::-1]  # next processing order of a BFS pass: list order == descending creation
        size = len(fin) + len(act)
        if size >= N or size == prev:
            finish = True
        elif size + 3 * len(act) > N:
            while not finish:
                rnd += 1
                prev = size
                # sort ascending by (size, creation) and walk from the back == descending (size, stamp idx)
                cand = sorted(act, key=lambda nd: (-(nd["e"] - nd["s"]), -nd["i"]))
                # how many get divided: stop right after the division that makes size >= N
                g = [sum(1 for ch in children(nd) if ch["e"] > ch["s"]) - 1 for nd in cand]
                k, acc = len(cand), size
                for j, gj in enumerate(g):
                    acc += gj
                    if acc >= N:
                        k = j + 1
                        break
                nf, na, growth = expand(cand[:k])
                fin += nf + cand[k:]
                act = na[::-1]
                size = len(fin) + len(act)
                assert size == prev + sum(growth)
                if size >= N or size == prev:
                    finish = True
    final = fin + act
    final.sort(key=lambda nd: (-nd["r"], -nd["i"]))
    out = []
    for nd in final:
        best = None
        for k in range(nd["s"], nd["e"]):
            i = perm[k]
            if best is None or scores[i] > scores[best] or (scores[i] == scores[best] and order[i] < order[best]):
                best = i
        out.append(best)
    return out
