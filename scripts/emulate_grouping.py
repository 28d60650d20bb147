"""Grouping policies of the Barnes-Hut traversal compared on the CPU with the issue-cycle model of DESIGN.md K5
(cycles per group = 160 x walk rounds + 30 x leaf-expansion rounds + 27 x list entries x slots / 64 + 400), calibrated
on the B200 (N = 10M: model 22.6 ms, measured 23.9 ms).  Same sampled segments for every policy.
Usage: python scripts/emulate_grouping.py [N] [theta]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from tests.conftest import plummer_cloud

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
THETA = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
BITS = 21
p = plummer_cloud(N)
t = oracle.Octree(p)
keys = t.keys
pos = p[t.perm, :3]
cm, mass = t.commass[:, :3], t.commass[:, 3]
nchild, first, level, count = t.n_child, t.first_child, t.level, t.count
ext = t.ext


def segments(lo, hi, lvl, seg):
    if hi - lo <= seg or lvl == BITS:
        return [(lo, hi, lvl)]
    shift = 3 * (BITS - lvl - 1)
    d = (keys[lo:hi] >> np.uint64(shift)) & np.uint64(7)
    cuts = lo + np.flatnonzero(np.diff(d.astype(np.int64)) != 0) + 1
    out = []
    for a, b in zip(np.concatenate([[lo], cuts]), np.concatenate([cuts, [hi]])):
        out += segments(int(a), int(b), lvl + 1, seg)
    return out


def walk(a, b):
    lo, hi = pos[a:b].min(0), pos[a:b].max(0)
    c, h = 0.5 * (lo + hi), 0.5 * (hi - lo)
    stack, rounds, leaf_rounds, entries = [0], 0, 0, 0
    while stack:
        k = min(32, len(stack))
        ids = np.array(stack[-k:][::-1])
        del stack[-k:]
        d = np.maximum(np.abs(cm[ids] - c) - h, 0.0)
        opened = THETA * THETA * (d * d).sum(1) < (ext * 0.5 ** level[ids]) ** 2
        for i in ids[opened & (nchild[ids] > 0)]:
            stack.extend(range(first[i], first[i] + nchild[i]))
        leaf = int(count[ids[opened & (nchild[ids] == 0)]].sum())
        rounds += 1
        leaf_rounds += (leaf + 31) // 32
        entries += leaf + int((~opened & (mass[ids] != 0)).sum())
    return rounds, leaf_rounds, entries


def slots(g, cap):
    return cap if g > cap // 2 else max(2, 1 << int(np.ceil(np.log2(g))))


def chunks(a, b, lvl, policy, cap):
    s = b - a
    if policy == "fixed":      # chunks of `cap` from the start, the remainder is one smaller group (the kernel's rule)
        return [(x, min(x + cap, b)) for x in range(a, b, cap)]
    if policy == "balanced":   # equal chunks
        k = -(-s // cap)
        e = np.linspace(a, b, k + 1).astype(int)
        return list(zip(e[:-1], e[1:]))
    if policy == "pow2":       # remainder cut into power-of-two pieces (>= 8) so that no slot is padding
        out = [(x, x + cap) for x in range(a, b - cap + 1, cap)]
        x = a + len(out) * cap
        piece = cap // 2
        while x < b:
            while piece > 8 and piece > b - x:
                piece //= 2
            out.append((x, min(x + piece, b)))
            x += piece
        return out
    if policy == "cells":      # cut at the sub-cell boundaries closest to multiples of `cap`
        if s <= cap:
            return [(a, b)]
        shift = 3 * (BITS - lvl - 2) if lvl + 2 <= BITS else 0
        d = keys[a:b] >> np.uint64(shift)
        cuts = a + np.flatnonzero(d[1:] != d[:-1]) + 1
        out, x = [], a
        while b - x > cap:
            cand = cuts[(cuts > x) & (cuts <= x + cap)]
            y = int(cand[-1]) if len(cand) else x + cap
            out.append((x, y))
            x = y
        out.append((x, b))
        return out
    raise ValueError(policy)


def model(seg, cap, policy, sample_frac=0.012, seed=3):
    segs = segments(0, N, 0, seg)
    rng = np.random.default_rng(seed)
    # sample by TARGETS: the same particle positions decide which segments are looked at under every policy
    marks = np.sort(rng.choice(N, max(int(sample_frac * len(segs)), 150), replace=False))
    starts = np.array([s[0] for s in segs])
    picked = sorted(set(np.searchsorted(starts, marks, side="right") - 1))
    cyc = tgt = inter = grp = 0
    for i in picked:
        a, b, lvl = segs[i]
        for x, y in chunks(a, b, lvl, policy, cap):
            r, lr, e = walk(x, y)
            g = y - x
            cyc += 160 * r + 30 * lr + 27 * e * slots(g, cap) / 64 + 400
            tgt += g
            inter += e * g
            grp += 1
    return cyc / tgt, inter / tgt, tgt / grp


print(f"N = {N} Plummer, theta = {THETA}; modelled issue cycles per target (lower is better)")
base = None
for seg, cap, policy in ((256, 64, "fixed"), (128, 64, "fixed"), (512, 64, "fixed"), (256, 32, "fixed"),
                         (256, 64, "balanced"), (256, 64, "pow2"), (256, 64, "cells"), (512, 64, "cells"),
                         (256, 128, "fixed"), (512, 128, "fixed"), (1024, 128, "fixed")):  # 4 targets per lane
    c, it, gs = model(seg, cap, policy)
    base = base or c
    print(f"segments <= {seg:3d}, groups <= {cap:2d}, {policy:8s}: {c:8.0f} cycles/target ({100 * c / base:5.1f} %), "
          f"{it:6.0f} interactions/target, mean group {gs:4.1f}")
