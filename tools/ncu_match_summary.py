"""Summarise an `ncu --set full` report of tools/prof_match.py into profiles/ (developer tool).

usage: python tools/ncu_match_summary.py gpurun_out/match.ncu-rep profiles/r1k_ncu_match_summary.json
One entry per kernel (its longest launch): time, issue-slot use, ALU / XU (POPC) pipe use, occupancy, DRAM bytes, registers,
warp instructions."""
import csv, io, json, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
M = {"time_us": "gpu__time_duration.sum", "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
     "alu_pipe_pct": "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
     "xu_pipe_pct": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
     "fma_pipe_pct": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
     "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active", "dram_read_MB": "dram__bytes_read.sum",
     "dram_write_MB": "dram__bytes_write.sum", "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
     "stall_math_throttle": "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
     "stall_long_scoreboard": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
     "regs": "launch__registers_per_thread", "inst_executed": "smsp__inst_executed.sum", "grid": "launch__grid_size"}
rows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
best = {}
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("orbb200::", "").split("<")[0]
    e = {"kernel": name}
    for k, m in M.items():
        if m not in ix:
            continue
        v, u = float(r[ix[m]].replace(",", "")), units[ix[m]]
        if k.endswith("_MB"):
            v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}[u]
        if k == "time_us":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}[u]
        e[k] = round(v, 4)
    if name not in best or e["time_us"] > best[name]["time_us"]:
        best[name] = e
json.dump(list(best.values()), open(out, "w"), indent=1)
print("wrote", out, {k: v["time_us"] for k, v in best.items()})
