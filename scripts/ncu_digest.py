"""Digest of an ncu report (read on the CPU box): key metrics, opcode mix, stall mix, hottest source lines.
Usage: python scripts/ncu_digest.py gpurun_out/<name>.ncu-rep [n_lines]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
nlines = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 30


def ncu(*args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


rows = list(csv.reader(io.StringIO(ncu("--page", "raw", "--csv"))))
h, v = rows[0], rows[-1]
want = ["gpu__time_duration.sum", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "launch__registers_per_thread",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
units = rows[1]
if "--traffic" in sys.argv:  # python scripts/ncu_digest.py <rep> --traffic <key>: update profiles/ncu_traffic.json
    import json
    import os
    key = sys.argv[sys.argv.index("--traffic") + 1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    val = lambda k: float(v[h.index(k)]) * scale[units[h.index(k)]]  # noqa: E731
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
    data = json.load(open(path)) if os.path.exists(path) else {}
    data[key] = {"dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
                 "kernel": v[h.index("Kernel Name")], "gpu_time_ms_under_ncu": float(v[h.index("gpu__time_duration.sum")]),
                 "source": f"ncu --set full --clock-control none, one launch; report {os.path.basename(rep)}"}
    json.dump(data, open(path, "w"), indent=1, sort_keys=True)
    print(json.dumps(data[key]))
    sys.exit(0)
for k in want:
    if k in h:
        print(f"{k:75s} {v[h.index(k)]} {units[h.index(k)]}")

rows = list(csv.reader(io.StringIO(ncu("--page", "source", "--csv"))))
hdr = rows[1]
iS, iI = hdr.index("# Samples"), hdr.index("Instructions Executed")
ops, samp, stall = collections.Counter(), collections.Counter(), collections.Counter()
cols = [i for i, x in enumerate(hdr) if x.startswith("stall_") and "Not Issued" not in x]
for r in rows[2:]:
    if len(r) < 10:
        continue
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[1].strip())
    op = m.group(2) if m else r[1]
    op = op.split(".")[0] if not op.startswith(("LDS", "STS", "LDG", "STG", "SHFL", "MUFU")) else op
    ops[op] += int(r[iI])
    samp[op] += int(r[iS])
    for i in cols:
        stall[hdr[i]] += int(r[i] or 0)
tot, ts = sum(ops.values()), sum(samp.values())
print(f"\nwarp instructions {tot/1e9:.3f} G")
for op, n in ops.most_common(22):
    print(f"  {op:20s} {n/1e9:7.3f} G {100*n/tot:5.1f} %   samples {100*samp[op]/max(ts,1):5.1f} %")
s = sum(stall.values())
print("stalls: " + ", ".join(f"{k[6:]} {100*x/s:.1f}%" for k, x in stall.most_common(8)))

rows = list(csv.reader(io.StringIO(ncu("--page", "source", "--csv", "--print-source", "cuda,sass"))))
hdr = next(r for r in rows if r and r[0] == "Line No")
iS, iI = hdr.index("# Samples"), hdr.index("Instructions Executed")
lines = []
for r in rows:
    if len(r) > 10 and r[0].isdigit():
        try:
            lines.append((int(r[0]), r[1], int(r[iS]), int(r[iI])))
        except ValueError:
            pass
tI, tS = sum(x[3] for x in lines), sum(x[2] for x in lines)
print("\nhottest source lines (by instructions):")
for ln, src, sm, ins in sorted(lines, key=lambda x: -x[3])[:nlines]:
    print(f"  {ln:5d} {100*ins/tI:5.1f}%I {100*sm/max(tS,1):5.1f}%S  {src.strip()[:95]}")
