"""Opening rules for a GROUP of targets compared on the CPU: interactions per target against the error of the
resulting accelerations, with the reference's own per-particle algorithm (oracle.barnes_hut) at the same theta as
the yardstick.  The rule in use opens a node when ANY member of the group could open it (distance from the group's
bounding box to the node's centre of mass): conservative, hence more accurate - and more work - than the reference.
Usage: python scripts/emulate_mac.py [N] [theta] [n_groups] [plummer|uniform] [box factors, comma separated]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from tests.conftest import plummer_cloud, uniform_cloud

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
THETA = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
NG = int(sys.argv[3]) if len(sys.argv) > 3 else 120
CLOUD = sys.argv[4] if len(sys.argv) > 4 else "plummer"
SEG, GROUP, BITS = 256, 64, 21
p = plummer_cloud(N) if CLOUD == "plummer" else uniform_cloud(N)
t = oracle.Octree(p)
keys = t.keys
sp = p[t.perm].astype(np.float64)
pos = sp[:, :3]
cm, mass = t.commass[:, :3].astype(np.float64), t.commass[:, 3].astype(np.float64)
nchild, first, level, begin, count = t.n_child, t.first_child, t.level, t.begin, t.count
ext = float(t.ext)


def segments(lo, hi, lvl):
    if hi - lo <= SEG or lvl == BITS:
        return [(lo, hi)]
    shift = 3 * (BITS - lvl - 1)
    d = (keys[lo:hi] >> np.uint64(shift)) & np.uint64(7)
    cuts = lo + np.flatnonzero(np.diff(d.astype(np.int64)) != 0) + 1
    out = []
    for a, b in zip(np.concatenate([[lo], cuts]), np.concatenate([cuts, [hi]])):
        out += segments(int(a), int(b), lvl + 1)
    return out


groups = [(s, min(s + GROUP, b)) for a, b in segments(0, N, 0) for s in range(a, b, GROUP)]
rng = np.random.default_rng(2)
# sample by target so that dense regions are represented as they are in the cloud
starts = np.array([g[0] for g in groups])
picked = sorted(set(np.searchsorted(starts, rng.choice(N, NG, replace=False), side="right") - 1))
sample = [groups[i] for i in picked]


def evaluate(rule):
    acc, inter = [], 0
    for a, b in sample:
        tg = pos[a:b]
        lo, hi = tg.min(0), tg.max(0)
        c, h = 0.5 * (lo + hi), 0.5 * (hi - lo)
        r = np.sqrt(((tg - c) ** 2).sum(1).max())
        stack, nodes, parts = [0], [], []
        while stack:
            ids = np.array(stack)
            stack = []
            w = ext * 0.5 ** level[ids]
            if rule[0] == "box":        # shrink the box by rule[1] (1 = the rule in use)
                d = np.maximum(np.abs(cm[ids] - c) - rule[1] * h, 0.0)
                d2 = (d * d).sum(1)
            elif rule[0] == "sphere":    # sphere of rule[1] x the group's radius around its centre
                d2 = np.maximum(np.sqrt(((cm[ids] - c) ** 2).sum(1)) - rule[1] * r, 0.0) ** 2
            elif rule[0] == "halves":    # boxes of the first and the second half of the group, the nearer one counts
                m = (b - a + 1) // 2
                d2 = None
                for part in (tg[:m], tg[m:] if b - a > m else tg[:m]):
                    lo2, hi2 = part.min(0), part.max(0)
                    dd = np.maximum(np.abs(cm[ids] - 0.5 * (lo2 + hi2)) - 0.5 * (hi2 - lo2), 0.0)
                    dd = (dd * dd).sum(1)
                    d2 = dd if d2 is None else np.minimum(d2, dd)
            else:                        # "both": the larger of the two lower bounds (still conservative)
                d = np.maximum(np.abs(cm[ids] - c) - h, 0.0)
                d2 = np.maximum((d * d).sum(1),
                                np.maximum(np.sqrt(((cm[ids] - c) ** 2).sum(1)) - r, 0.0) ** 2)
            opened = THETA * THETA * d2 < w * w
            for i in ids[opened & (nchild[ids] > 0)]:
                stack.extend(range(first[i], first[i] + nchild[i]))
            for i in ids[opened & (nchild[ids] == 0)]:
                parts.append(np.arange(begin[i], begin[i] + count[i]))
            nodes.append(ids[~opened & (mass[ids] != 0)])
        nodes = np.concatenate(nodes)
        parts = np.concatenate(parts) if parts else np.zeros(0, np.int64)
        src = np.concatenate([cm[nodes], pos[parts]])
        m = np.concatenate([mass[nodes], sp[parts, 3]])
        d = src[None, :, :] - tg[:, None, :]
        r2 = (d * d).sum(2)
        with np.errstate(divide="ignore", invalid="ignore"):
            s = np.where(r2 > 0, m[None, :] * r2 ** -1.5, 0.0)
        acc.append((d * s[:, :, None]).sum(1))
        inter += len(src) * (b - a)
    return np.concatenate(acc), inter


tg_idx = np.concatenate([np.arange(a, b) for a, b in sample])
tg32 = np.ascontiguousarray(p[t.perm][tg_idx, :3])
exact = oracle.brute_force_exact(tg32, p)
ref = oracle.barnes_hut(tg32, p, THETA, parallel=True)
den = np.linalg.norm(exact, axis=1)


def stats(a):
    e = np.linalg.norm(a - exact, axis=1) / den
    return np.median(e), np.percentile(e, 99), e.max()


rs = stats(ref)
print(f"{CLOUD} N = {N}, theta = {THETA}, {len(sample)} groups / {len(tg_idx)} targets")
print(f"reference algorithm (per-particle rule, oracle): median {rs[0]:.2e}  p99 {rs[1]:.2e}  max {rs[2]:.2e}")
RULES = [(x, 1.0) if x in ("both", "halves") else ("box", float(x)) for x in sys.argv[5].split(",")] if len(sys.argv) > 5 else \
    [("box", 1.0), ("box", 0.75), ("box", 0.5), ("box", 0.25), ("box", 0.0), ("sphere", 1.0), ("sphere", 0.5)]
for rule in RULES:
    a, inter = evaluate(rule)
    s = stats(a)
    print(f"{rule[0]:6s} x {rule[1]:4.2f}: {inter / len(tg_idx):6.0f} interactions/target   median {s[0]:.2e} ({s[0] / rs[0]:4.2f} x ref)  "
          f"p99 {s[1]:.2e} ({s[1] / rs[1]:4.2f} x)  max {s[2]:.2e} ({s[2] / rs[2]:4.2f} x)")
