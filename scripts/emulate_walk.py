"""CPU emulation of the warp-cooperative walk of traverse2_kernel (round structure only, no forces): how full the
32-lane rounds are, how many rounds a group takes, how full the leaf-expansion rounds are.  Uses the CPU statement of
the tree (oracle.Octree) and the same grouping rule (maximal cells of <= 256 targets cut into chunks of 64).
Usage: python scripts/emulate_walk.py [N] [theta]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from tests.conftest import plummer_cloud

N = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
THETA = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
SEG, GROUP, BITS = 256, 64, 21
p = plummer_cloud(N)
t = oracle.Octree(p)
keys = t.keys
pos = p[t.perm, :3]


def segments(lo, hi, level):
    if hi - lo <= SEG or level == BITS:
        return [(lo, hi)]
    shift = 3 * (BITS - level - 1)
    d = (keys[lo:hi] >> np.uint64(shift)) & np.uint64(7)
    cuts = lo + np.flatnonzero(np.diff(d.astype(np.int64)) != 0) + 1
    out = []
    for a, b in zip(np.concatenate([[lo], cuts]), np.concatenate([cuts, [hi]])):
        out += segments(int(a), int(b), level + 1)
    return out


groups = []
for a, b in segments(0, N, 0):
    for s in range(a, b, GROUP):
        groups.append((s, min(s + GROUP, b)))
rng = np.random.default_rng(1)
sample = [groups[i] for i in rng.choice(len(groups), min(400, len(groups)), replace=False)]
cm, mass = t.commass[:, :3], t.commass[:, 3]
nchild, first, level, begin, count = t.n_child, t.first_child, t.level, t.begin, t.count
ext = t.ext
tot = dict(rounds=0, tests=0, leaf_rounds=0, leaf_entries=0, node_entries=0, short_rounds=0, groups=0, slots=0, targets=0)
hist = np.zeros(33, np.int64)
for a, b in sample:
    lo, hi = pos[a:b].min(0), pos[a:b].max(0)
    c, h = 0.5 * (lo + hi), 0.5 * (hi - lo)
    stack = [0]
    while stack:
        k = min(32, len(stack))
        ids = np.array(stack[-k:][::-1])
        del stack[-k:]
        d = np.maximum(np.abs(cm[ids] - c) - h, 0.0)
        d2 = (d * d).sum(1)
        w = ext * 0.5 ** level[ids]
        opened = THETA * THETA * d2 < w * w
        for i in ids[opened & (nchild[ids] > 0)]:
            stack.extend(range(first[i], first[i] + nchild[i]))
        leaf = int(count[ids[opened & (nchild[ids] == 0)]].sum())
        tot["rounds"] += 1
        tot["tests"] += k
        hist[k] += 1
        tot["short_rounds"] += k < 32
        tot["leaf_rounds"] += (leaf + 31) // 32
        tot["leaf_entries"] += leaf
        tot["node_entries"] += int((~opened & (mass[ids] != 0)).sum())
    g = b - a
    tot["groups"] += 1
    tot["targets"] += g
    tot["slots"] += 2 if g <= 2 else 1 << int(np.ceil(np.log2(g)))
r = tot["rounds"]
print(f"N = {N}, theta = {THETA}: {len(groups)} groups (mean {N / len(groups):.1f} targets), {tot['groups']} sampled")
print(f"rounds per group {r / tot['groups']:.1f}, node tests per round {tot['tests'] / r:.2f} of 32 "
      f"({100 * tot['tests'] / (32 * r):.1f} % lane use), rounds with < 32 nodes {100 * tot['short_rounds'] / r:.1f} %")
print(f"leaf expansion: {tot['leaf_rounds'] / tot['groups']:.1f} 32-entry rounds per group, fill "
      f"{100 * tot['leaf_entries'] / max(32 * tot['leaf_rounds'], 1):.1f} %")
print(f"entries per group: {tot['node_entries'] / tot['groups']:.0f} nodes + {tot['leaf_entries'] / tot['groups']:.0f} particles; "
      f"target slots {tot['slots'] / tot['targets']:.3f} per target (padding {100 * (tot['slots'] / tot['targets'] - 1):.1f} %)")
print("round size histogram (nodes popped: rounds):", {int(k): int(v) for k, v in enumerate(hist) if v})
