"""FP32-pipe probe modes (csrc/probe.cu) on the GPU box."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import particular_b200 as pb
ctx = pb.CudaContext(0)
names = ["FFMA scalar", "FFMA2 3 regs", "FFMA2 2 regs", "FMUL2", "FADD2 bcast", "1 FFMA2 : 2 FFMA",
         "2 FFMA2 : 2 FFMA", "FFMA2 + MUFU/3", "12 FFMA2 : 2 MUFU", "12 FFMA2 : 2 FMNMX", "12 FFMA2",
         "12 FFMA2:2MUFU:2FMNMX"]
for mode, name in enumerate(names):
    tf, ms = ctx.probe_fp32(mode, 8192, 5)
    print(f"mode {mode} {name:18s}: {tf:7.2f} TFLOP/s-equivalent  ({ms:.3f} ms)  {100*tf/74.45:5.1f}% of 74.45")
