"""profiles/roofline_traffic.json from one `ncu --set full` report of the draw wave (raster_loop_kernel, 512 instances).
usage: make_roofline_json.py <report.ncu-rep> <name of the committed csv> <pixels shaded by the captured launch>"""
import csv
import io
import json
import os
import subprocess
import sys

rep, committed, pixels = sys.argv[1], sys.argv[2], float(sys.argv[3])
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
d = dict(zip(rows[0], rows[2]))
u = dict(zip(rows[0], rows[1]))


def mb(key):
    v = float(d[key])
    return v * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[u[key]]


captured = mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum")
inst = float(d["smsp__inst_executed.sum"])
issue = float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"]) / 100
alu = float(d["sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"]) / 100
lsu = float(d["l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"]) / 100
j = {
    "raster_kernel_dram_bytes_per_launch": int(captured * 8),
    "source": f"profiles/{committed}: dram__bytes_read.sum + dram__bytes_write.sum of the draw-wave launch (raster_loop_kernel) "
              "captured at 512 instances (ncu --set full), x8 for the 4096-instance bench launch",
    "captured_instances": 512,
    "captured_bytes": int(captured),
    "captured_ms": float(d["gpu__time_duration.sum"]),
    "warp_instructions": int(inst),
    "warp_instructions_per_pixel": inst / pixels,
    "issue_slot_frac": issue,
    "alu_pipe_frac": alu,
    "lsu_pipe_frac": lsu,
    "limiter": f"instruction issue and the load/store data pipe, not HBM: {issue * 100:.0f} % of the issue slots busy at "
               f"{float(d['sm__warps_active.avg.per_cycle_active']):.0f} resident warps/SM, the half-rate integer ALU pipe "
               f"{alu * 100:.0f} % busy, the L1/shared-memory data pipe {lsu * 100:.0f} % (texel gathers: ~8 wavefronts per "
               f"load, plus the shared-memory tile and palette traffic), L2 hit rate {float(d['lts__t_sector_hit_rate.pct']):.0f} %, {inst / pixels:.2f} warp "
               f"instructions per shaded pixel (ncu --set full, profiles/{committed})",
}
with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "roofline_traffic.json"), "w") as f:
    json.dump(j, f, indent=1)
print(json.dumps(j, indent=1))
