"""Summarise an .ncu-rep (ncu --set full) into a small JSON under profiles/: per captured launch the duration, DRAM
traffic, pipe utilisations, issue activity and the warp-stall mix.  Run here (no GPU needed).

    python tools/ncu_summary.py gpurun_out/r02_x.ncu-rep profiles/r02_x_summary.json
"""
import csv
import io
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "pipe_xu_mufu_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "pipe_tensor_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed": "issue_active_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "launch__registers_per_thread": "registers_per_thread", "launch__grid_size": "grid", "launch__block_size": "block",
    "smsp__inst_executed.sum": "warp_instructions",
}
launches = []
for r in rows[2:]:
    d = {"kernel": r[hdr.index("Kernel Name")]}
    stalls = {}
    for i, h in enumerate(hdr):
        if h in KEYS:
            try:
                d[KEYS[h]] = {"value": float(r[i].replace(",", "")), "unit": units[i]}
            except ValueError:
                pass
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                stalls[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = round(float(r[i]), 3)
            except ValueError:
                pass
    d["warps_stalled_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8])
    launches.append(d)
json.dump({"report": rep, "launches": launches}, open(out, "w"), indent=1)
for d in launches:
    print(d["kernel"][:60], {k: v["value"] for k, v in d.items() if isinstance(v, dict) and "value" in v and k in ("duration", "dram_read", "dram_write", "pipe_xu_mufu_pct", "pipe_tensor_pct", "issue_active_pct")})
