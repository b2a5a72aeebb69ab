"""Regenerate profiles/README.md from the committed measurement files (developer tool): python tools/profiles_readme.py r1k"""
import csv, json, os, sys
from collections import defaultdict

tag = sys.argv[1] if len(sys.argv) > 1 else "r1o"
P = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")
J = lambda n: json.load(open(os.path.join(P, n)))
l = J(tag + "_bench.json")
xtag = tag if os.path.exists(os.path.join(P, tag + "_ncu_full_summary.json")) else "r1l"     # the extractor kernels' last full capture
d = J(xtag + "_ncu_full_summary.json")
rows = [r for r in csv.reader(open(os.path.join(P, tag + "_launches.csv"))) if len(r) > 5]
ix = {h: i for i, h in enumerate(rows[0])}
t, n = defaultdict(float), defaultdict(int)
for r in rows[1:]:
    nm = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("orbb200::", "")
    try:
        v = float(r[ix["Metric Value"]].replace(",", ""))
    except ValueError:
        continue
    if nm.startswith("k_"):
        t[nm] += v; n[nm] += 1
tot = sum(t.values())
ev = l["kernel_ms_per_step"]; evtot = sum(ev.values())
stage = {"k_resize": "pyramid", "k_fast<26, 42>": "fast", "k_quadtree": "quadtree", "k_blur": "blur", "k_describe": "describe", "k_pad_level0": None}
tab = "| kernel | launches | total us (ncu) | ncu share | CUDA-event ms / 256-frame step | event share |\n|---|---|---|---|---|---|\n"
for k in stage:
    if k not in t:
        continue
    s = stage[k]
    tab += "| `%s` | %d | %.0f | %.1f %% | %s | %s |\n" % (k, n[k], t[k] / 1e3, 100 * t[k] / tot, ("%.3f" % ev[s]) if s else "(host-call legs only)",
                                                        ("%.1f %%" % (100 * ev[s] / evtot)) if s else "")
rz = [e for e in d if e["kernel"] == "k_resize"]
ktab = "| kernel (one 256-frame launch) | time us | issue slots busy | resident warps | DRAM read MB | DRAM write MB | DRAM % of peak | regs | warp instructions |\n|---|---|---|---|---|---|---|---|---|\n"
ktab += "| `k_resize` x7 (levels 1..7) | %.0f | %.0f-%.0f %% | %.0f-%.0f %% | %.0f | %.0f | %.0f-%.0f | %d | %.0f M |\n" % (
    sum(e["time_us"] for e in rz), min(e["issue_active_pct"] for e in rz), max(e["issue_active_pct"] for e in rz),
    min(e["warps_active_pct"] for e in rz), max(e["warps_active_pct"] for e in rz), sum(e["dram_read_MB"] for e in rz),
    sum(e["dram_write_MB"] for e in rz), min(e["dram_pct"] for e in rz), max(e["dram_pct"] for e in rz), int(rz[0]["regs"]), sum(e["inst_executed"] for e in rz) / 1e6)
for e in d:
    if e["kernel"] != "k_resize":
        ktab += "| `%s` | %.0f | %.0f %% | %.0f %% | %.0f | %.0f | %.1f | %d | %.0f M |\n" % (
            e["kernel"], e["time_us"], e["issue_active_pct"], e["warps_active_pct"], e["dram_read_MB"], e["dram_write_MB"], e["dram_pct"],
            int(e["regs"]), e["inst_executed"] / 1e6)
m, p, c = l["matching"], l["pipeline"], l["cpu_baseline"]
st = l.get("stereo")
mk = {e["kernel"]: e for e in J(tag + "_ncu_match_summary.json")}
ip = J("r1k_int_peak.json")
sfi = m["search_for_initialization"]
mtab = "| kernel (longest launch) | time us | issue slots busy | ALU pipe | resident warps | regs | warp instructions |\n|---|---|---|---|---|---|---|\n"
for k in ("k_build_grid", "k_init_topk", "k_search_init", "k_proj_topk", "k_search_proj", "k_last_topk", "k_search_last"):
    if k in mk:
        e = mk[k]
        mtab += "| `%s` | %.0f | %.0f %% | %.0f %% | %.0f %% | %d | %.0f M |\n" % (k, e["time_us"], e["issue_active_pct"], e.get("alu_pipe_pct", 0), e["warps_active_pct"], int(e["regs"]), e["inst_executed"] / 1e6)
mrows = "\n".join("| `%s` | %s | %.2f |" % (k, v["workload"], v.get("ms_per_step", v.get("ms_per_call"))) for k, v in m.items())
scale = ""
for f, what in (("_bench_2gpu", "2 GPUs, 256 VGA frames per GPU (weak)"), ("_bench_4gpu", "4 GPUs, 256 VGA frames per GPU (weak)"), ("_bench_8gpu", "8 GPUs, 256 VGA frames per GPU (weak)"),
                ("_bench_hd_1gpu", "1 GPU, 1280x720 / 2000 kp, 1024 frames"), ("_bench_hd_4gpu", "4 GPUs, the same 1024 HD frames (strong)")):
    for src in (tag, "r1n", "r1m", "r1l", "r1k", "r1i"):            # the newest run of each configuration (the multi-GPU runs are not repeated for every step)
        fn = os.path.join(P, src + f + ".json")
        if what and os.path.exists(fn):
            x = json.load(open(fn))
            scale += "| %s (`%s`) | %.0f | %.0f | %.0f |\n" % (what, src + f + ".json", x["value"], x["e2e"]["value"], x["e2e"]["blocking_call_value"])
            break
fast = [e for e in d if e["kernel"] == "k_fast"][0]
readme = f"""# profiles/ — measured evidence, named per round and step

Everything here was produced on a B200 through `gpurun`; bench numbers are never taken under a profiler.

| file | what | command |
|---|---|---|
| `{tag}_bench.json` | **current code**: `python bench.py --steps 20 --warmup 3`, no profiler (extractor, every matching row, device pipeline, CPU reference arm) | |
| `{tag}_launches.csv` | ncu launch list of the same code (`gpu__time_duration.sum`, `--clock-control none`, cold-cache and serialised) | `ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-matching` |
| `{xtag}_ncu_full_summary.json` | `ncu --set full --clock-control none --import-source on`, one batch-256 launch of every extractor kernel (unchanged since) | `ncu --set full ... -k regex:"k_fast|k_blur|k_describe|k_resize|k_quadtree" -s 11 -c 11 python tools/prof_extract.py`, summarised by `tools/ncu_summary.py` |
| `{xtag}_ncu_source_k_*.json` | the source page of the same capture: share of executed instructions and of stall samples per source line (lines >= 1 %) | `tools/ncu_summary.py` |
| `traffic.json` | `dram__bytes_read.sum + dram__bytes_write.sum` per launch and stage from that capture; `bench.py` reports it as `roofline.traffic` | |
| `{tag}_bench_4gpu.json`, `{tag}_bench_8gpu.json`, `{tag}_bench_hd_1gpu.json`, `{tag}_bench_hd_4gpu.json` | scaling runs of the current code (below) | `torchrun --nproc-per-node N bench.py --gpus N [--workload hd]` |
| `r1j_bench_8gpu.json`, `r1j_topology_8gpu.txt` | the 8-GPU run repeated with each rank bound to its GPU's CPU set (no change: the box is one NUMA node with 32 virtual CPUs for 8 ranks) | |
| `{tag}_ncu_match_summary.json` | `ncu --set full` of the matcher kernels (longest launch per kernel) | `ncu --set full ... -k regex:"k_init_topk|k_search_init|k_proj_topk|k_search_proj|k_last_topk|k_search_last|k_build_grid" -c 14 python tools/prof_match.py`, summarised by `tools/ncu_match_summary.py` |
| `r1k_int_peak.json` | integer-pipe peaks (POPC, LOP3, IADD3 lanes per clock and SM) and register-only Hamming rates: the matching roofline's denominator | `tools/int_peak.cu` |
| `r1p_bench_extractor_only.json` | `bench.py --no-cpu-baseline --no-matching` after one more `k_fast` change (pass 1 without the divergent region around its loads: 0.578 -> 0.572 ms, 175.7 k frames/s); everything else in `r1o_*` is unchanged by it; `r1p_ncu_full_summary.json`, `r1p_ncu_source_k_fast.json`: `ncu --set full` of that final `k_fast` (one batch-256 launch: 582 us under ncu, 463 M warp instructions, issue slots 71 %, ALU pipe 75 %, 24 resident warps per SM of the 25 that 8 KB + 1 KB of shared memory per one-warp CTA allow) | `ncu --set full ... -k regex:k_fast -s 1 -c 1 python tools/prof_extract.py` |
| `r1n_*` | the step before (one FAST pass at min(iniThFAST, minThFAST) instead of iniThFAST first: `k_fast` 0.716 ms, 160 k frames/s); its 1-GPU HD run and the matcher capture `r1n_ncu_match_summary.json` (matcher kernels unchanged since) still stand | |
| `r1m_*`, `r1m_pipe_utilisation.csv`, `r1m_fast_shared_memory.csv` / `r1n_fast_shared_memory.csv` | two steps before (before the bank-conflict-free FAST pass 1: `k_fast` 0.747 ms, 156 k frames/s); pipe utilisation of every extractor kernel; shared-memory wavefronts and bank conflicts of `k_fast` before / after; `r1n_ncu_match_summary.json` is a copy of `r1m_…` (matcher kernels unchanged) | |
| `r1l_*` | three steps before (before the threshold pruning in `k_init_topk`: SearchForInitialization 5.71 ms) | |
| `r1k_*` | four steps before (before the CSR-ordered candidate records and the CUDA graph of small host calls); its 4- and 8-GPU runs are the current multi-GPU evidence | |
| `r1i_*`, `r1h_ncu_match_summary.json` | an earlier step of this round (before the carry-save distance, the speculative resolve, the row-pair blur and the stereo row): 152 k frames/s, SearchForInitialization 12.3 ms | |
| `r1c_*` ... `r1h_*` | earlier steps of this round, kept for the record (`r1f`: FAST 0.76 ms, describe 0.32 ms; `r1e`: FAST 0.85 ms) | |
| `r1a_*`, `r1b_*` | first bit-exact CUDA path, before any tuning | |
| `tools/profiles_readme.py` | writes this file from the ones above | |

## {tag} (current): one B200, 256 VGA frames per step, 1000 keypoints, 8 levels

* device-resident: **{l['value']:.0f} frames/s** ({l['ms_per_step']:.3f} ms per step, CUDA events, L2 flushed before every step)
* end to end, host buffers in and out (`StreamingExtractor`, three handles in turn; every step uploads {l['e2e']['h2d_bytes_per_step']/1e6:.1f} MB of frames
  and downloads {l['e2e']['d2h_bytes_per_step']/1e6:.1f} MB of keypoints + descriptors): **{l['e2e']['value']:.0f} frames/s**;
  the blocking single call `orbb200_extract_host` reaches {l['e2e']['blocking_call_value']:.0f} frames/s; one frame per call (what `Frame::ExtractORB` sees): {l['e2e']['single_frame_latency_ms']:.3f} ms
* CPU reference arm (`oracle/_ref` = the reference's own `ORBextractor.cc`, {c['cores']} host threads): {c['value']:.0f} frames/s
* SM clock {l['clocks']['sm_mhz']:.0f} MHz of {l['clocks']['sm_max_mhz']:.0f}, throttle reasons {l['clocks']['reasons']}
* device pipeline (128 frame pairs: 2 x extraction + undistort/grid + SearchForInitialization, no host round trip): {p['ms_per_step']:.2f} ms = {p['frames_per_s']:.0f} frames/s
* stereo pipeline (128 rectified pairs: 2 x extraction + keypoint views + ComputeStereoMatches on the extractors' device pyramids): {st['ms_per_step']:.2f} ms = {st['pairs_per_s']:.0f} pairs/s; the stereo matcher alone {st['stereo_matcher_ms']:.3f} ms

### scaling (frames/s: device-resident | end to end streamed | end to end blocking call)

| run | device-resident | streamed e2e | blocking e2e |
|---|---|---|---|
| 1 GPU, 256 VGA frames | {l['value']:.0f} | {l['e2e']['value']:.0f} | {l['e2e']['blocking_call_value']:.0f} |
{scale}
Device-resident throughput scales 4.0x on 4 and 8.0x on 8 GPUs (no data-path collective). The end-to-end figures stop scaling
at 3-4 GPUs and vary from box to box: the ranks share one virtual host (32 vCPUs, one NUMA node, `r1j_topology_8gpu.txt`) and
each asks it for 54 GB/s of page-locked uploads (175 k frames/s x 307 KB), which that host does not deliver eight times over; the
HD workload is upload-bound already on one GPU (0.92 MB per frame against ~55 GB/s of PCIe).

### matching rows (`{tag}_bench.json`)

| routine | workload | ms |
|---|---|---|
{mrows}

The headline matching config (4096 x 1000 x 1000 = 4.1 G candidate evaluations) is bound by the integer pipes, not by bytes.
`tools/int_peak.cu` measured them on this GPU (`r1k_int_peak.json`): {ip['popc_per_clk_per_sm']:.1f} POPC and {ip['lop3_per_clk_per_sm']:.1f} LOP3 lanes per clock and SM; a
register-only loop reaches {ip['hamming256_Geval_s']['popc8']:.0f} G evaluations/s with the plain 8-POPC distance and {ip['hamming256_Geval_s']['csa_popc6']:.0f} / {ip['hamming256_Geval_s']['csa_popc4']:.0f} G/s with 6 / 4 POPC after
carry-save adders.  `k_init_topk` evaluates the 128-bit half distance of every candidate (3 POPC) and finishes only the
candidates that can still matter (5 POPC more for about one in ten; exact, see DESIGN.md).  Against the POPC pipe at 3 POPC per
candidate ({sfi['roofline']['peak']:.0f} G candidates/s) the whole call (grid + top-4 + greedy resolve) reaches **{sfi['roofline']['achieved']:.0f} G candidates/s =
{100*sfi['roofline']['frac']:.0f} %** (`{tag}_bench.json: matching.search_for_initialization.roofline`), the kernel alone {4096e6/mk['k_init_topk']['time_us']/1e3:.0f} G/s with the
ALU pipe {mk['k_init_topk'].get('alu_pipe_pct',0):.0f} % busy (`{tag}_ncu_match_summary.json`).  The greedy, order-dependent halves (`k_search_*`) run one warp per
frame (pair) and decide 32 queries speculatively per round; their time is a dependent chain per frame, not throughput.

{mtab}
### kernel shares: ncu launch list vs CUDA events

The ncu pass covers the whole short bench run (3 warm-up + 2 timed device-resident steps of 256 frames, then the host-call
legs, whose launches are smaller); the event column is the 20-step run without a profiler.

{tab}
The two agree within 2 points per kernel.

### what bounds each kernel (ncu `--set full`, one launch)

{ktab}
No kernel is near the HBM roofline (DRAM throughput 1-20 % of peak): the path is integer/byte arithmetic with a few
operations per byte, and every kernel is bound by instruction issue (50-80 % of issue slots busy) with the remaining slots
lost to dependent-load latency.  `k_fast` (the kernel `bench.py` names in `roofline`) moves {l['roofline']['traffic']/1e6:.0f} MB of DRAM traffic per
launch against {l['roofline']['achieved']*ev['fast']:.0f} MB algorithmic, so there are no wasted re-reads; it executes {fast['inst_executed']/1e6:.0f} M warp instructions per launch
(about {fast['inst_executed']/243200:.0f} per 30 x 30-px cell) at {fast['issue_active_pct']:.0f} % issue utilisation, which is what its {ev['fast']:.2f} ms is made of.  The work this round went into removing
instructions (`{xtag}_ncu_source_k_fast.json` shows where the remaining ones are): 1.80 ms -> {ev['fast']:.2f} ms for `k_fast`,
0.60 -> {ev['describe']:.2f} ms for `k_describe`, 0.91 -> {ev['blur']:.2f} ms for `k_blur`, 3.88 ms -> {l['ms_per_step']:.2f} ms for the step.

## r1a/r1b (first correct path, before tuning) -- kept for the record

| kernel | ncu avg / launch | ncu share | bench event ms / step |
|---|---|---|---|
| `k_resize` x7 | 58.5 us | 10.5 % | 0.405 |
| `k_fast` | 1801 us | 46.0 % | 1.786 |
| `k_quadtree` | 164 us | 4.2 % | 0.166 |
| `k_blur` | 925 us | 23.6 % | 0.912 |
| `k_describe` | 599 us | 15.3 % | 0.594 |

66.1 k frames/s device-resident, 45.4 k frames/s end to end.
"""
open(os.path.join(P, "README.md"), "w").write(readme)
print("wrote profiles/README.md")
