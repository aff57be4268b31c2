"""Rebuild profiles/step_kernel_traffic.json (read by bench.py -> roofline.traffic) from the ncu outputs in gpurun_out/:
r02_prof_api_counts.ncu-rep (one isolated launch, --set full) and r02_counts_steady_state{,_noalt}.csv (application
replay inside a stepping loop).  Records the sha256 of the kernel sources the library is built from; bench.py quotes the
numbers only while that still matches."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from dexterous_rl_manipulation_b200.build import build_info  # noqa: E402


def avg(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    mi, vi = hdr.index("Metric Name"), hdr.index("Metric Value")
    agg = {}
    for r in rows[1:]:
        agg.setdefault(r[mi], []).append(float(r[vi]))
    return {k: sum(v) / len(v) for k, v in agg.items()}


out = subprocess.run(["ncu", "-i", os.path.join(ROOT, "gpurun_out", "r02_prof_api_counts.ncu-rep"), "--page", "raw", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
rd = float(rows[2][hdr.index("dram__bytes_read.sum")]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[rows[1][hdr.index("dram__bytes_read.sum")]]
wr = float(rows[2][hdr.index("dram__bytes_write.sum")]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[rows[1][hdr.index("dram__bytes_write.sum")]]
a = avg(os.path.join(ROOT, "gpurun_out", "r02_counts_steady_state.csv"))
b = avg(os.path.join(ROOT, "gpurun_out", "r02_counts_steady_state_noalt.csv"))
bi = build_info()
assert bi["fresh"], "rebuild the library first"
doc = {
    "envs": 1048576, "dram_bytes_per_launch": int(rd + wr), "dram_bytes_read": int(rd), "dram_bytes_write": int(wr),
    "how": "ncu --set full --clock-control none (kernel replay, caches flushed before every pass): one isolated launch",
    "steady_state": {
        "dram_bytes_per_launch": int(a["dram__bytes_read.sum"] + a["dram__bytes_write.sum"]),
        "dram_bytes_read": int(a["dram__bytes_read.sum"]), "dram_bytes_write": int(a["dram__bytes_write.sum"]),
        "lts_sector_hit_rate_pct": round(a["lts__t_sector_hit_rate.pct"], 1),
        "same_direction_every_step_dram_bytes_per_launch": int(b["dram__bytes_read.sum"] + b["dram__bytes_write.sum"]),
        "how": "ncu --replay-mode application --cache-control none: launches 9-12 of a stepping loop, L2 as the previous step left "
               "it, walk direction alternating from step to step (profiles/r02_counts_steady_state.csv); "
               "profiles/r02_counts_steady_state_noalt.csv holds the same with one direction"},
    "kernel": "step_tma_kernel<dense, AoS, counts-only, 2 stages> (the bench's main loop)",
    "report": "profiles/r02_step_tma_tile_width_ncu.txt, last block (launch 9 of tools/profile_step.py 1048576 api_counts)",
    "step_kernel_sha256": bi["step_kernel_sha256"],
    "step_kernel_files": ["csrc/dexsim_step_tma.cuh", "csrc/dexsim_core.cuh"],
}
json.dump(doc, open(os.path.join(ROOT, "profiles", "step_kernel_traffic.json"), "w"), indent=1)
for f in ("r02_counts_steady_state.csv", "r02_counts_steady_state_noalt.csv"):
    open(os.path.join(ROOT, "profiles", f), "w").write(open(os.path.join(ROOT, "gpurun_out", f)).read())
print(json.dumps(doc, indent=1))
