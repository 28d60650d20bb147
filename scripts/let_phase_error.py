"""Error statistics of the locally essential trees with one walk and with the two-phase walk, against the
single tree on one GPU (ranks as threads on one GPU, tests/local_ranks.py).  GPU box: python scripts/let_phase_error.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
import particular_b200 as pb  # noqa: E402
from particular_b200._ffi import lib  # noqa: E402
from tests.conftest import plummer_cloud, rel_err, uniform_cloud  # noqa: E402
from tests.local_ranks import LocalWorld  # noqa: E402


def stats(e):
    return "median %.4g p99 %.4g max %.4g" % (np.median(e), np.percentile(e, 99), e.max())


for name, p, worlds in (("plummer 40003", plummer_cloud(40003, seed=4), (2, 4)),
                        ("plummer 1M", plummer_cloud(1_000_000, seed=1808), (4, 8)),
                        ("uniform 1M", uniform_cloud(1_000_000, seed=7), (8,))):
    idx = np.sort(np.random.default_rng(1).choice(len(p), min(len(p), 2048), replace=False))
    exact = oracle.brute_force_exact(p[idx, :3], p)
    with pb.CudaContext(0) as c1:
        one = pb.BarnesHut(c1, 0.5, pb.Acceleration.checked()).compute(p)
    print(name, "| one GPU:", stats(rel_err(one[idx], exact)), flush=True)
    for world in worlds:
        for overlap in (0, 1):
            assert lib.pcuda_debug_set(b"bh_forest", 3) == 0 and lib.pcuda_debug_set(b"bh_let_overlap", overlap) == 0
            with LocalWorld(world) as w:
                got = w.barnes_hut(p, 0.5)
            print(f"  {world} ranks, {'two-phase' if overlap else 'one walk '}:", stats(rel_err(got[idx], exact)), flush=True)
