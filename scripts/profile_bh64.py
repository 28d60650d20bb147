import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import particular_b200 as pb
from tests.conftest import plummer_cloud
for n in (1_000_000, 4_000_000):
    p = plummer_cloud(n, seed=1808, dtype=np.float64)
    d_p = torch.from_numpy(p).cuda(); d_o = torch.zeros((n,3), dtype=torch.float64, device="cuda")
    with pb.CudaContext(0) as ctx:
        bh = pb.BarnesHut(ctx, 0.5, pb.Acceleration.checked())
        for it in range(3):
            bh.compute_device(None, n, d_p.data_ptr(), n, d_o.data_ptr(), "f64x3"); ctx.sync()
            t = ctx.timings()
            print(f"f64 BH N={n} iter {it}: build {t['build_ms']:.3f} ms traverse {t['compute_ms']:.3f} ms launches {t['kernel_launches']}", flush=True)
