"""Regenerate profiles/README.md (round 2) from the committed measurement files: python tools/profiles_readme_r2.py [bench tag] [ncu tag]"""
import csv, json, os, sys
from collections import defaultdict

btag = sys.argv[1] if len(sys.argv) > 1 else "r2d"
ntag = sys.argv[2] if len(sys.argv) > 2 else "r2d"
P = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")
J = lambda n: json.load(open(os.path.join(P, n)))
l = J(btag + "_bench.json")
ref = J(btag + "_bench_reference_arm.json") if os.path.exists(os.path.join(P, btag + "_bench_reference_arm.json")) else None
d = J(ntag + "_ncu_full_summary.json")
out = []
w = out.append
w("# profiles/ -- measured evidence, named per round and step\n")
w("Everything here was produced on a B200 through `gpurun`; bench numbers are never taken under a profiler. Round 1's files and their")
w("description: `README_r1.md`. Round 2:\n")
w("| file | what |\n|---|---|")
w("| `%s_bench.json`, `%s_bench_{2,4,8}gpu.json` | `python bench.py --steps 20 --warmup 3` (under torchrun for N > 1), no profiler: VGA extraction (headline), the HD config and both matching configs sharded over the ranks, CPU baselines, parity |" % (btag, btag))
w("| `%s_bench_reference_arm.json` | `python bench.py --impl reference`: the reference's own ORBextractor.cc / ORBmatcher.cc on the host threads |" % btag)
w("| `%s_launches.csv` | ncu launch list (`gpu__time_duration.sum`, `--clock-control none`; cold-cache and serialised: compare shares) of `bench.py --steps 2 --warmup 3 --no-matching --no-cpu-baseline` |" % ntag)
w("| `%s_ncu_full_summary.json`, `%s_ncu_source_k_*.json` | `ncu --set full --clock-control none --import-source on`, one batch-256 launch of every extractor kernel (`tools/prof_extract.py`), summarised by `tools/ncu_summary.py`: time, issue slots, resident warps, ALU / FMA / XU pipes, LSU data-pipe wavefronts, DRAM bytes; instruction and stall-sample share per source line |" % (ntag, ntag))
w("| `traffic.json` | DRAM read + write bytes per launch and stage from that capture (`roofline.traffic` in the bench line) |")
w("| `r2_sass_digest.json` | per-kernel instruction-class counts of the shipped `liborb_b200.so` (`tools/sass_digest.py`): `UTMALDG` (TMA tensor copies) in `k_resize3`, `k_fast2`, `k_blur` and `k_describe3`, `LDS.128` / `LDG.E.128` where used |")
w("| `r2a_pipe_peak.json` | issue rates of the instruction classes the kernels are made of, alone and in pairs (`tools/pipe_peak.cu`): which pipe each sits on |")
w("| `r2_pcie_aggregate.json` | pinned host <-> device copy bandwidth with 1, 2, 4, 8 ranks copying at once (`tools/pcie_aggregate.py`): the end-to-end ceiling of the host |")
w("| `r2_tiebreak.json` | the reference under glibc malloc against the canonical quadtree order (`tools/tiebreak_report.py`) |")
w("| `r1k_int_peak.json` | POPC / LOP3 / IADD3 peaks (round 1), the matching roofline's denominator |\n")

ev = l["kernel_ms_per_step"]
w("## %s: one B200, 256 VGA frames per step, 1000 keypoints, 8 levels\n" % btag)
w("* device-resident: **%.0f frames/s** (%.3f ms per step, CUDA events on the launch stream, L2 flushed before every step)" % (l["value"], l["ms_per_step"]))
e = l["e2e"]
w("* end to end, host buffers in and out (`StreamingExtractor`; every step uploads %.1f MB and downloads %.1f MB): **%.0f frames/s** = %.0f %% of what this host's PCIe paths deliver for the same copies without kernels (%.0f frames/s, measured in the same run); blocking call %.0f frames/s; one frame per call %.3f ms"
  % (e["h2d_bytes_per_step"] / 1e6, e["d2h_bytes_per_step"] / 1e6, e["value"], 100 * e["host_copy_ceiling"]["frac"], e["host_copy_ceiling"]["value"], e["blocking_call_value"], e["single_frame_latency_ms"]))
c = l["cpu_baseline"]
w("* CPU baseline (`oracle/_ref` = the reference's own ORBextractor.cc): %.0f frames/s on %d host threads, %.1f ms per frame on one thread" % (c["value"], c["cores"], c["single_thread_ms_per_frame"]))
w("* clocks during the timed region: %s" % json.dumps(l["clocks"]))
w("* parity in the same run (first 8 frames against the reference): %s\n" % json.dumps(l["parity"]))
w("| stage | CUDA-event ms per 256-frame step | share | algorithmic GB/s / measured HBM peak |\n|---|---|---|---|")
tot = sum(ev.values())
for k, v in ev.items():
    w("| %s | %.3f | %.1f %% | %.3f |" % (k, v, 100 * v / tot, l["roofline"]["all"][k]))
w("\nwhole step: %.3f of the HBM peak (the kernels are bound by the SM's integer and shared-memory pipes, DESIGN.md 4)\n" % l["roofline"]["whole_step"])

w("### ncu --set full (%s), one batch-256 launch per kernel\n" % ntag)
w("| kernel | time us | issue slots | resident warps | ALU pipe | FMA pipe | LSU data pipe | regs | warp instructions |\n|---|---|---|---|---|---|---|---|---|")
rz = [x for x in d if x["kernel"].startswith("k_resize")]
if rz:
    w("| `" + rz[0]["kernel"] + "` x%d | %.0f | %.0f-%.0f %% | %.0f-%.0f %% | %.0f-%.0f %% | %.0f-%.0f %% | %.0f-%.0f %% | %d | %.0f M |" % (
        len(rz), sum(x["time_us"] for x in rz), min(x["issue_active_pct"] for x in rz), max(x["issue_active_pct"] for x in rz),
        min(x["warps_active_pct"] for x in rz), max(x["warps_active_pct"] for x in rz), min(x["alu_pipe_pct"] for x in rz), max(x["alu_pipe_pct"] for x in rz),
        min(x["fma_pipe_pct"] for x in rz), max(x["fma_pipe_pct"] for x in rz), min(x["lsu_data_pipe_pct"] for x in rz), max(x["lsu_data_pipe_pct"] for x in rz),
        int(rz[0]["regs"]), sum(x["inst_executed"] for x in rz) / 1e6))
for x in d:
    if not x["kernel"].startswith("k_resize"):
        w("| `%s` | %.0f | %.0f %% | %.0f %% | %.0f %% | %.0f %% | %.0f %% | %d | %.0f M |" % (x["kernel"], x["time_us"], x["issue_active_pct"], x["warps_active_pct"],
          x["alu_pipe_pct"], x.get("fma_pipe_pct", 0), x.get("lsu_data_pipe_pct", 0), int(x["regs"]), x["inst_executed"] / 1e6))

# launch list shares
fn = os.path.join(P, ntag + "_launches.csv")
if os.path.exists(fn):
    rows = [r for r in csv.reader(open(fn)) if len(r) > 5]
    ix = {h: i for i, h in enumerate(rows[0])}
    t, n = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        nm = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("orbb200::", "").split("<")[0]
        try:
            v = float(r[ix["Metric Value"]].replace(",", ""))
        except ValueError:
            continue
        if nm.startswith("k_"):
            t[nm] += v; n[nm] += 1
    tt = sum(t.values())
    w("\n### launch list (%s_launches.csv): share of the summed kernel time\n" % ntag)
    w("| kernel | launches | share |\n|---|---|---|")
    for k in sorted(t, key=lambda k: -t[k]):
        w("| `%s` | %d | %.1f %% |" % (k, n[k], 100 * t[k] / tt))

w("\n## scaling (frames/s, device-resident | end to end | host copy ceiling; HD = 1024 frames 1280x720 / 2000 kp split over the GPUs)\n")
w("| GPUs | VGA device | VGA e2e | copy ceiling | HD device | HD e2e | SearchForInitialization T evals/s | SearchByProjection G map points/s |\n|---|---|---|---|---|---|---|---|")
for n_ in (1, 2, 4, 8):
    f = btag + ("_bench.json" if n_ == 1 else "_bench_%dgpu.json" % n_)
    if not os.path.exists(os.path.join(P, f)):
        continue
    x = J(f); m = x["matching"]
    w("| %d | %.0f | %.0f | %.0f | %.0f | %.0f | %.2f | %.2f |" % (n_, x["value"], x["e2e"]["value"], x["e2e"]["host_copy_ceiling"]["value"], x["hd"]["value"], x["hd"]["e2e"]["value"],
      m["search_for_initialization"]["distance_evals_per_s"] / 1e12, m["search_by_projection"]["map_points_per_s"] / 1e9))
w("\nDevice-resident extraction scales with the GPU count (no data-path collective). End to end every configuration sits at 87-100 % of what")
w("the host delivers when the same buffers are copied with no kernels at all -- the `copy ceiling` column, measured inside each run, because it")
w("depends on the box the run landed on (`r2_pcie_aggregate.json` is one such box: 54 GB/s up for one rank, 99 for two, 103 for four -- its GPUs")
w("0-3 share one path -- and 161 for eight; the 4-GPU box of this table delivered 540 k frames/s, the 8-GPU box 515 k). The matching configs are")
w("strong scaling of millisecond-sized jobs.\n")

m = l["matching"]
w("## matching rows (%s_bench.json, one GPU)\n" % btag)
w("| routine | workload | ms | CPU baseline (reference, host threads) |\n|---|---|---|---|")
for k, v in m.items():
    cb = v.get("cpu_baseline")
    w("| `%s` | %s | %.3f | %s |" % (k, v["workload"], v.get("ms_per_step", v.get("ms_per_call", 0)),
      ("%.3g %s on %d threads" % (cb["value"], cb["unit"], cb["cores"])) if cb else ""))
sfi, sfb = m["search_for_initialization"], m.get("search_for_initialization_extracted_brute")
w("\nSearchForInitialization against the integer-pipe bound (3 POPC per candidate at least): %.0f %% on the unrelated descriptors of configs[2]"
  " (%.1f %% of the candidates survive the 128-bit half-distance test)" % (100 * sfi["roofline"]["frac"], 100 * sfi["half_distance_survivor_fraction"]) +
  ((", %.0f %% on real extracted descriptors (%.1f %% survive)" % (100 * sfb["roofline"]["frac"], 100 * sfb["half_distance_survivor_fraction"])) if sfb else "") + ".\n")
hd = l["hd"]
w("## HD (configs[3]) on one GPU: %.0f frames/s device-resident, %.0f end to end; kernel ms per 1024-frame step: %s; CPU: %s\n" % (
    hd["value"], hd["e2e"]["value"], json.dumps({k: round(v, 3) for k, v in hd["kernel_ms_per_step"].items()}), json.dumps(hd.get("cpu_baseline", {}))))
open(os.path.join(P, "README.md"), "w").write("\n".join(out) + "\n")
print("wrote profiles/README.md")
