"""Condense `ncu --set full` captures of the render kernel into profiles/r2_ncu_metrics.json: the hardware-counter figures bench.py
quotes in its roofline object (issue-active %, lanes per instruction, pipe utilisations, cache hit rates).
usage: ncu_metrics_json.py OUT.json scene=REPORT.ncu-rep [scene=REPORT ...]"""
import csv
import json
import subprocess
import sys

WANT = {"issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active", "lanes_per_inst": "smsp__thread_inst_executed_per_inst_executed.ratio",
        "xu_pct": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "alu_pct": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "fma_pct": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "lsu_pct": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "tex_pipe_pct": "sm__inst_executed_pipe_tex.avg.pct_of_peak_sustained_active", "l1tex_hit_pct": "l1tex__t_sector_hit_rate.pct",
        "l2_hit_pct": "lts__t_sector_hit_rate.pct", "dram_read_bytes": "dram__bytes_read.sum", "dram_write_bytes": "dram__bytes_write.sum",
        "duration_ns": "gpu__time_duration.sum", "warp_inst": "smsp__inst_executed.sum", "registers": "launch__registers_per_thread",
        "icc_hit_pct": "sm__icc_request_hit_rate.pct", "gcc_inst_requests": "gcc__cache_requests_type_instruction.sum",
        "gcc_inst_requests_pct_of_peak": "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed",
        "stall_long_scoreboard": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "stall_no_instruction": "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "stall_wait": "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "stall_not_selected": "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"}
SCALE = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "ms": 1e6, "us": 1e3, "ns": 1.0, "s": 1e9}
out = {}
for arg in sys.argv[2:]:
    scene, rep = arg.split("=", 1)
    rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
    hdr, units, r = rows[0], rows[1], rows[2]
    d = {"source": "ncu --set full --clock-control none, " + rep.split("/")[-1] + " (profiles/), kernel " + r[hdr.index("Kernel Name")][:40]}
    for k, m in WANT.items():
        if m in hdr:
            i = hdr.index(m)
            d[k] = float(r[i].replace(",", "")) * SCALE.get(units[i], 1.0)
    out[scene] = d
json.dump(out, open(sys.argv[1], "w"), indent=1, sort_keys=True)
print(json.dumps(out, indent=1, sort_keys=True))
