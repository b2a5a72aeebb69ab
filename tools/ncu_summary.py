"""Summarise an `ncu --set full` report of tools/prof_extract.py into profiles/ (developer tool).

usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1f
writes <prefix>_ncu_full_summary.json (per-launch: time, issue-slot use, occupancy, DRAM bytes, top stalls),
<prefix>_ncu_source_<kernel>.json (instruction share per source line, lines >= 1 %) and profiles/traffic.json
(DRAM read+write bytes per launch and stage, what bench.py reports as roofline.traffic)."""
import csv, io, json, os, subprocess, sys

rep, prefix = sys.argv[1], sys.argv[2]
STAGE = {"k_resize": "pyramid", "k_resize2": "pyramid", "k_resize3": "pyramid", "k_fast": "fast", "k_fast2": "fast", "k_quadtree": "quadtree",
         "k_blur": "blur", "k_describe": "describe", "k_describe2": "describe", "k_describe3": "describe"}
M = {"time_us": "gpu__time_duration.sum", "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
     "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active", "dram_read_MB": "dram__bytes_read.sum",
     "dram_write_MB": "dram__bytes_write.sum", "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
     "stall_long_scoreboard": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
     "stall_short_scoreboard": "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
     "stall_wait": "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
     "stall_barrier": "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
     "stall_math_throttle": "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
     "alu_pipe_pct": "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
     "xu_pipe_pct": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
     "shared_mem_pipe_pct": "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed",
     "lsu_data_pipe_pct": "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
     "fma_pipe_pct": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
     "regs": "launch__registers_per_thread", "smem_per_block": "launch__shared_mem_per_block_dynamic",
     "inst_executed": "smsp__inst_executed.sum"}


def ncu(*a):
    return subprocess.run(["ncu", "-i", rep, *a], capture_output=True, text=True).stdout


rows = list(csv.reader(io.StringIO(ncu("--page", "raw", "--csv"))))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
out, traffic = [], {}
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
        e[k] = "%g %s" % (v, u) if k == "smem_per_block" else round(v, 4)
    out.append(e)
    st = STAGE.get(name)
    if st:
        traffic[st] = traffic.get(st, 0.0) + (e["dram_read_MB"] + e["dram_write_MB"]) * 1e6
json.dump(out, open(prefix + "_ncu_full_summary.json", "w"), indent=1)
json.dump(traffic, open(os.path.join(os.path.dirname(prefix), "traffic.json"), "w"), indent=1)

for kern in ("k_fast2", "k_describe3", "k_blur", "k_resize3", "k_quadtree"):
    rows = list(csv.reader(io.StringIO(ncu("--page", "source", "--print-source", "cuda,sass", "--csv", "-k", "regex:" + kern))))
    acc, tot, smp = {}, 0, 0
    for r in rows:
        if len(r) < 8 or not r[0].strip().isdigit() or not r[7].isdigit():
            continue
        ln = int(r[0]); a = acc.setdefault(ln, [r[1].strip()[:140], 0, 0])
        a[1] += int(r[7]); a[2] += int(r[4]) if r[4].isdigit() else 0
        tot += int(r[7]); smp += int(r[4]) if r[4].isdigit() else 0
    lines = [{"line": ln, "inst_pct": round(100 * a[1] / tot, 2), "stall_samples_pct": round(100 * a[2] / max(smp, 1), 2), "src": a[0]}
             for ln, a in sorted(acc.items()) if a[1] * 100 >= tot or a[2] * 100 >= smp]
    json.dump({"kernel": kern, "launches_summed": "all captured", "instructions": tot, "lines": lines},
              open("%s_ncu_source_%s.json" % (prefix, kern), "w"), indent=1)
print("wrote", prefix + "_ncu_full_summary.json", "traffic:", {k: round(v / 1e6, 1) for k, v in traffic.items()})
