"""Dev tool: summarise one kernel of an ncu report (`ncu --set full ... -o x`) as JSON for profiles/ and bench.py's
roofline.traffic.   usage: ncu_metrics_json.py <report.ncu-rep> <n_envs or 0> <out.json> [kernel substring]"""
import csv
import json
import subprocess
import sys

rep, n_envs, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
want_k = sys.argv[4] if len(sys.argv) > 4 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
H, U = rows[0], rows[1]
R = next(r for r in rows[2:] if want_k is None or want_k in r[H.index("Kernel Name")])
KEEP = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_active.avg.per_cycle_active", "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor_op_utchmma.avg.pct_of_peak_sustained_active",
        "sm__ops_path_tensor_op_utchmma_src_tf32.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]
d = {"kernel": R[H.index("Kernel Name")]}
for k in KEEP:
    if k in H:
        i = H.index(k)
        try:
            d[k] = {"unit": U[i], "value": float(R[i].replace(",", ""))}
        except ValueError:
            pass
mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
if "dram__bytes_read.sum" in d:
    d["traffic_bytes_per_launch"] = sum(d[k]["value"] * mult[d[k]["unit"]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
if n_envs:
    d["n_envs"] = n_envs
json.dump(d, open(out, "w"), indent=1)
print(json.dumps({k: (v["value"] if isinstance(v, dict) else v) for k, v in d.items()}))
