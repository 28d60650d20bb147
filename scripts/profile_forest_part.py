"""One partitioned step (8 virtual ranks, N = 10M Plummer) for an ncu launch list of the per-part build."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import particular_b200 as pb
from particular_b200._ffi import check, lib
from tests.conftest import plummer_cloud

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
PARTS = int(sys.argv[2]) if len(sys.argv) > 2 else 8
p = plummer_cloud(N)
d_p = torch.from_numpy(p).cuda()
d_o = torch.empty((N, 3), dtype=torch.float32, device="cuda")
with pb.CudaContext(0) as ctx:
    for _ in range(2):
        check(lib.pcuda_barneshut_f32x3_partitioned_dev(ctx.handle, d_p.data_ptr(), N, PARTS, 0.5, 0.0, 1,
                                                        d_o.data_ptr()), ctx.handle)
        ctx.sync()
    print(ctx.timings())
