"""Sizes of the locally essential trees at a given rank count, on ONE GPU (ranks as threads over the
in-process communicator; the stage times printed by the trace are not meaningful here, the sizes are).
Usage (GPU box): python scripts/let_sizes.py [N] [world] [theta]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from particular_b200._ffi import lib
from tests.conftest import plummer_cloud
from tests.local_ranks import LocalWorld

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
W = int(sys.argv[2]) if len(sys.argv) > 2 else 8
THETA = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
p = plummer_cloud(N)
assert lib.pcuda_debug_set(b"bh_forest", 3) == 0
assert lib.pcuda_debug_set(b"bh_let_trace", 1) == 0
with LocalWorld(W) as w:
    got = w.barnes_hut(p, THETA)
print("finite", bool(np.isfinite(got).all()))
