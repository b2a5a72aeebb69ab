#!/usr/bin/env python
"""bench.py -- ORB frames/s on B200 (BASELINE.json: 640x480, 1000 kp, 8 levels; configs[1]:
batched extraction of 256 synthetic frames on a single B200).

  python bench.py --gpus N --steps K --warmup W            our arm (one rank per GPU under torchrun for N>1)
  python bench.py --impl reference --gpus N ...            the reference's own CPU implementation (oracle/_ref)

A step = one pass of ORBextractor::operator() over one batch of 256 frames per GPU.  `value` is
frames/s with the frames already resident in HBM (device time, CUDA events, max over ranks);
`e2e` is the same metric through the host-buffer C-ABI call (H2D + kernels + D2H inside the timed
region).  Frames are independent, so N GPUs = N shards, no collective on the data path (weak scaling);
the only cross-rank traffic is the max-reduction of the elapsed time.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTH, HEIGHT, NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH = 640, 480, 1000, 1.2, 8, 20, 7
BATCH = 256                       # frames per GPU per step (configs[1])
WORKLOAD = "batched ORB extraction, 256 synthetic 640x480 frames per GPU, 1000 kp/8 levels, FAST 20/7 (configs[1])"
METRIC = "ORB frames/s (640x480, 1000 kp, 8 lvl)"
DISTINCT = BATCH                  # distinct synthetic frames generated per rank (every frame of a step is its own image)
HD = dict(width=1280, height=720, nfeatures=2000, total=1024, distinct=64,
          metric="ORB frames/s (1280x720, 2000 kp, 8 lvl)",
          alg_bytes={"pyramid": 2781331 + 1931488, "fast": 2853088 + 160000, "quadtree": 160000 + 32000,
                     "blur": 5706176, "describe": 1498000 + 1024000 + 120000})
MATCH_PAIRS, MATCH_N = 4096, 1000              # configs[2]
PROJ_FRAMES, PROJ_KP, PROJ_MP = 512, 2000, 10000   # configs[4]

# Algorithmic HBM bytes per frame and per kernel (DESIGN.md "Roofline model"):
# every stage reads its input once and writes its output once.
ALG_BYTES = {
    "pyramid": 926546 + 643332,
    "fast": 950532 + 80000,
    "quadtree": 80000 + 16000,
    "blur": 950532 + 950532,
    "describe": 749000 + 512000 + 60000,
}
STAGES = ["pyramid", "fast", "quadtree", "blur", "describe"]


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


def _frames(rank, count, width=None, height=None, distinct=None):
    import numpy as np
    from weiner_slamit_v2_b200.frames import synthetic_frame
    width, height = width or WIDTH, height or HEIGHT
    base = [synthetic_frame(rank * 100000 + i, width, height) for i in range(min(distinct or DISTINCT, count))]
    reps = (count + len(base) - 1) // len(base)
    return np.stack((base * reps)[:count])


def _config(world):
    """The workload description both arms print, key for key (the driver compares the two dicts)."""
    return {"workload": WORKLOAD, "frames_per_step_per_gpu": BATCH, "distinct_frames_per_gpu": min(DISTINCT, BATCH),
            "l2": "256 MiB flush buffer written before every timed step (GPU arm); inputs of a step exceed the CPU caches (reference arm)",
            "parallelism": "frames sharded by batch, no data-path collective",
            "also_in_this_line": "hd = configs[3] (1024 frames 1280x720 / 2000 kp split over the GPUs), matching = configs[2] "
                                 "(4096 pairs) and configs[4] (512 frames) split over the GPUs"}


# --------------------------------------------------------------------------------------------
# reference arm: the reference's own ORBextractor.cc (oracle/_ref) on the host cores
# --------------------------------------------------------------------------------------------
def _cpu_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _load_ref():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ref_lib
    if ref_lib.available():
        return ref_lib, "reference"
    return None, "port"


def cpu_extract_fps(frames, threads, nfeatures=None):
    """frames/s of the CPU implementation over `frames` with `threads` host threads."""
    import numpy as np
    nfeatures = nfeatures or NFEAT
    ref_lib, kind = _load_ref()
    t0 = time.perf_counter()
    if ref_lib is not None:
        counts = ref_lib.extract_batch_mt(frames, threads, nfeatures, SCALE, NLEVELS, INI_TH, MIN_TH)
    else:  # oracle port, single instance per thread via a thread pool (ctypes releases the GIL)
        import oracle_lib
        from concurrent.futures import ThreadPoolExecutor
        def work(chunk):
            o = oracle_lib.OracleExtractor(nfeatures, SCALE, NLEVELS, INI_TH, MIN_TH)
            return [len(o(f)[0]) for f in chunk]
        chunks = [frames[i::threads] for i in range(threads)]
        with ThreadPoolExecutor(threads) as ex:
            counts = np.concatenate([np.asarray(c) for c in ex.map(work, chunks)])
    dt = time.perf_counter() - t0
    return len(frames) / dt, kind, int(np.sum(counts))


def cpu_single_thread(frames):
    """BASELINE configs[0] / SURVEY 8(d) config 1: the CPU implementation on ONE host thread, ms per 640x480 frame."""
    cpu_extract_fps(frames[:2], 1)
    fps, kind, _ = cpu_extract_fps(frames, 1)
    return {"single_thread_ms_per_frame": 1e3 / fps, "single_thread_frames_per_s": fps, "single_thread_sample": "%d frames" % len(frames)}


def cpu_primitive_ratio(frames):
    """How much of the CPU baseline is the stand-in OpenCV: the three image primitives ORBextractor spends its time in (the
    resize chain, FAST per level, GaussianBlur per level) on one thread, once through the scalar mini-cv / oracle primitives the
    reference arm links and once through cv2 (OpenCV's SIMD build, same results bit for bit: tests/test_oracle_primitives.py)."""
    try:
        import cv2
    except Exception:
        return None
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    cv2.setNumThreads(1)
    ex = O.OracleExtractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH)
    ex(frames[0])
    sizes = [ex.level_shape(l) for l in range(NLEVELS)]            # (h, w) per level

    def run(resize, fast, blur, timed=True):
        if timed:
            run(resize, fast, blur, False)             # first calls initialise the library
        t0 = time.perf_counter()
        for img in (frames if timed else frames[:1]):
            lv = [img]
            for l in range(1, NLEVELS):
                lv.append(resize(lv[-1], sizes[l][1], sizes[l][0]))
            for im in lv:
                fast(im); blur(im)
        return (time.perf_counter() - t0) / len(frames) * 1e3
    det = cv2.FastFeatureDetector_create(INI_TH, True)
    t_cv = run(lambda im, w, h: cv2.resize(im, (w, h), interpolation=cv2.INTER_LINEAR), lambda im: det.detect(im, None),
               lambda im: cv2.GaussianBlur(im, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101))
    t_or = run(lambda im, w, h: O.resize_linear(im, w, h), lambda im: O.fast(im, INI_TH, True), lambda im: O.gaussian_blur7(im))
    return {"scalar_primitives_ms_per_frame": t_or, "opencv_simd_primitives_ms_per_frame": t_cv, "ratio": t_or / t_cv,
            "what": "resize chain + FAST(%d) + GaussianBlur 7x7 on all %d levels of %d frames, one host thread; cv2 %s" % (INI_TH, NLEVELS, len(frames), cv2.__version__)}


def cpu_matching_baselines(threads):
    """The reference's own ORBmatcher.cc (oracle/_ref/libref_matcher.so; the oracle port if it was not built) on the host
    threads for configs[2] and configs[4], on a bounded sample of the same synthetic workload: one call per frame pair /
    frame as Tracking issues them, `threads` calls in flight (ORBmatcher is stateless, S/System.cc:156,160).  The time of a
    call includes the harness building the ORB_SLAM2::Frame / MapPoint objects the reference's function takes."""
    from concurrent.futures import ThreadPoolExecutor
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    import ref_lib as R
    from weiner_slamit_v2_b200.workloads import SCALE_FACTORS_8, init_pair, projection_frame
    use_ref = R.matcher_available()
    kind = "reference" if use_ref else "port"
    if use_ref:
        R.mlib()
    out = {}

    def timed(fn, work):
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(fn, work[:threads]))                      # warm-up: first touch of every thread
            t0 = time.perf_counter()
            res = list(ex.map(fn, work))
            return time.perf_counter() - t0, res

    n = MATCH_N
    pairs = [init_pair(900000 + i, n=n, brute_force=True) for i in range(max(4 * threads, 32))]
    f = R.ref_search_for_initialization if use_ref else O.search_for_initialization
    dt, res = timed(lambda p: f(p[0], p[1], p[2], p[3], p[4], (0, 0, 640, 480), 0.9, True, 1000)[0], pairs)
    out["search_for_initialization"] = {
        "value": len(pairs) * n * n / dt, "unit": "distance evaluations/s", "pairs_per_s": len(pairs) / dt,
        "accepted_matches_per_s": float(sum(res)) / dt, "cores": threads, "kind": kind,
        "sample": "%d of the %d frame pairs (1000 x 1000 descriptors, ratio 0.9), %d host threads, one call per pair" % (len(pairs), MATCH_PAIRS, threads)}
    frames = [projection_frame(900000 + i, PROJ_KP, PROJ_MP) for i in range(max(2 * threads, 16))]
    g = R.ref_search_by_projection if use_ref else O.search_by_projection
    dt, res = timed(lambda w: g(w[2], w[0], w[1], SCALE_FACTORS_8, (0, 0, 1280, 720), 0.8, 1.0)[0], frames)
    out["search_by_projection"] = {
        "value": len(frames) * PROJ_MP / dt, "unit": "map points/s", "frames_per_s": len(frames) / dt,
        "accepted_matches_per_s": float(sum(res)) / dt, "cores": threads, "kind": kind,
        "sample": "%d of the %d frames (10000 map points vs 2000 keypoints, th=1, ratio 0.8), %d host threads, one call per frame" % (len(frames), PROJ_FRAMES, threads)}
    return out


def cpu_hd_baseline(threads):
    import numpy as np
    frames = _frames(0, max(2 * threads, 16), HD["width"], HD["height"], HD["distinct"])
    cpu_extract_fps(frames[:threads], threads, HD["nfeatures"])
    fps, kind, _ = cpu_extract_fps(frames, threads, HD["nfeatures"])
    return {"value": fps, "unit": "frames/s", "cores": threads, "kind": kind,
            "sample": "%d of the %d frames (1280x720, 2000 kp), %d host threads, one ORBextractor each" % (len(frames), HD["total"], threads)}


def parity_report(frames, kps, desc, counts):
    """north_star's reporting duty: the device results of the first frames of the step against the CPU implementation
    (oracle/_ref = the reference's own ORBextractor.cc when it was built, else the oracle port) -- checker use only.
    Angles are required within 1e-3 degrees and descriptor bits that flip because of them below 0.1 %; here the
    keypoint records and the descriptors are identical, so both figures are zero."""
    import numpy as np
    ref_lib, kind = _load_ref()
    if ref_lib is not None:
        cpu = ref_lib.RefExtractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH)
    else:
        import oracle_lib
        cpu = oracle_lib.OracleExtractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH)
    from weiner_slamit_v2_b200._lib import KP_DTYPE
    n_kp = n_same = bits = flipped = 0
    max_angle = 0.0
    same_count = True
    for f, img in enumerate(frames):
        ko, do = cpu(img)
        c = int(counts[f])
        same_count &= c == len(ko)
        m = min(c, len(ko))
        kg = np.frombuffer(kps[f, :m].tobytes(), KP_DTYPE)
        n_kp += m
        n_same += int(sum(kg[i].tobytes() == ko[i].tobytes() for i in range(m)))
        if m:
            da = np.abs(kg["angle"].astype(np.float64) - ko["angle"][:m].astype(np.float64))
            max_angle = max(max_angle, float(np.minimum(da, 360.0 - da).max()))
            x = np.bitwise_xor(np.asarray(desc[f, :m], np.uint8), do[:m])
            flipped += int(np.unpackbits(x).sum()); bits += m * 256
    return {"against": kind, "frames_checked": len(frames), "keypoint_counts_equal": bool(same_count), "keypoints": n_kp,
            "keypoint_records_identical": n_same, "max_angle_diff_deg": max_angle, "descriptor_bits_flipped": flipped,
            "descriptor_flip_rate": flipped / bits if bits else 0.0}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = _cpu_threads()
    sample = BATCH                                   # one full step: 256 frames
    frames = _frames(0, sample)
    for _ in range(args.warmup):
        cpu_extract_fps(frames[: max(threads, 8)], threads)
    total_t, kind = 0.0, "port"
    for _ in range(args.steps):
        fps, kind, _ = cpu_extract_fps(frames, threads)
        total_t += sample / fps
    value = sample * args.steps / total_t
    base = {"value": value, "unit": "frames/s", "cores": threads, "kind": kind,
            "sample": "%d of the %d frames of one step, %d host threads, one ORBextractor each" % (sample, BATCH, threads),
            "primitives": "the reference's ORBextractor.cc over the scalar mini-cv stand-in for OpenCV (about 3x slower than OpenCV's SIMD build)"}
    base.update(cpu_single_thread(frames[:8]))
    pr = cpu_primitive_ratio(frames[:4])
    if pr:
        base["primitive_cost"] = pr
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_t / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": _config(args.gpus),
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if args.workload == "vga":
        if not args.no_matching:
            line["matching"] = cpu_matching_baselines(threads)
        line["hd"] = dict(cpu_hd_baseline(threads), metric=HD["metric"])
    _emit(line)


def _bind_to_gpu_numa_node(local):
    """Pin this rank to the CPUs NVML reports as local to its GPU, so that the page-locked frame and result buffers of
    the end-to-end leg are allocated on that socket (with 8 ranks the uploads otherwise cross the socket link)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[local]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else local
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass                                   # affinity is an optimisation of the host side only


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from weiner_slamit_v2_b200 import ORBextractor

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the extractor has no CPU fallback")
    torch.cuda.set_device(local)
    _bind_to_gpu_numa_node(local)            # before any pinned allocation: first touch puts the host buffers next to the GPU
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    ex = ORBextractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, width=WIDTH, height=HEIGHT, max_batch=BATCH, device=local)
    ex.set_profiling(True)
    frames = _frames(rank, BATCH)
    pinned = torch.empty((BATCH, HEIGHT, WIDTH), dtype=torch.uint8, pin_memory=True)
    pinned.numpy()[:] = frames
    d_frames = pinned.cuda()
    stream = torch.cuda.ExternalStream(ex.stream, device=local)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        ex.extract_device(d_frames, BATCH, WIDTH, WIDTH * HEIGHT)

    for _ in range(max(args.warmup, 3)):
        step()
    ex.sync()

    # ---- device-resident throughput: K steps, L2 flushed before each, CUDA events on the launch stream
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    stage_ms = np.zeros(5)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    with torch.cuda.stream(stream):
        for k in range(args.steps):
            flush.fill_(k & 0xff)
            ev[k][0].record()
            step()
            ev[k][1].record()
            stage_ms += ex.stage_ms()          # waits for this step's last stage event
    ex.sync()
    barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    launches = ex.last_launches * args.steps

    # ---- end to end through the host-buffer C-ABI call: H2D + kernels + D2H inside the timed region
    host_in = pinned.numpy()
    cap = ex.max_keypoints
    out_k = torch.empty((BATCH, cap, 28), dtype=torch.uint8, pin_memory=True)
    out_d = torch.empty((BATCH, cap, 32), dtype=torch.uint8, pin_memory=True)
    out_c = torch.empty((BATCH,), dtype=torch.int32, pin_memory=True)
    from weiner_slamit_v2_b200._lib import check

    def e2e_step():
        check(ex._L.orbb200_extract_host(ex._h, host_in.ctypes.data, BATCH, WIDTH, WIDTH * HEIGHT,
                                         out_k.data_ptr(), out_d.data_ptr(), out_c.data_ptr(), cap))
    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()                              # synchronous: returns after the D2H copies
    torch.cuda.synchronize()
    e2e_sync_s = time.perf_counter() - t0
    barrier()
    nkp = int(out_c.sum())
    # the streaming form of the same call (StreamingExtractor: orbb200_extract_host_async on `depth` handles in
    # turn): every step still uploads its frames from pinned host memory and downloads its keypoints,
    # descriptors and counts; the copies of neighbouring steps run beside the kernels.
    from weiner_slamit_v2_b200 import StreamingExtractor
    depth = args.stream_depth
    sx = StreamingExtractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, width=WIDTH, height=HEIGHT, max_batch=BATCH,
                            device=local, depth=depth)
    outs = [(torch.empty((BATCH, cap, 28), dtype=torch.uint8, pin_memory=True),
             torch.empty((BATCH, cap, 32), dtype=torch.uint8, pin_memory=True),
             torch.empty((BATCH,), dtype=torch.int32, pin_memory=True)) for _ in range(depth)]

    def stream_steps(n):
        for k in range(n):
            o = outs[k % depth]
            sx.submit(host_in, BATCH, WIDTH, WIDTH * HEIGHT, o[0], o[1], o[2], cap)
        sx.drain()
    stream_steps(max(3, depth))
    barrier()
    t0 = time.perf_counter()
    stream_steps(args.steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop() if sampler else None
    stream_ok = all(int(o[2].sum()) == nkp for o in outs[:min(depth, args.steps)])
    sx.close()
    if not stream_ok:
        raise SystemExit("bench.py: streamed results differ from the synchronous call")
    # the host's copy ceiling in the same run: every rank uploads its step's frames and downloads its step's results from / to
    # the same page-locked buffers, concurrently, with no kernels in between (profiles/r2_pcie_aggregate.json has the full probe)
    def copy_ceiling(seconds=0.6):
        d_in = torch.empty((BATCH, HEIGHT, WIDTH), dtype=torch.uint8, device="cuda")
        d_k = torch.empty((BATCH, cap, 28), dtype=torch.uint8, device="cuda"); d_d = torch.empty((BATCH, cap, 32), dtype=torch.uint8, device="cuda")
        su, sd = torch.cuda.Stream(), torch.cuda.Stream()
        def once():
            with torch.cuda.stream(su):
                d_in.copy_(pinned, non_blocking=True)
            with torch.cuda.stream(sd):
                out_k.copy_(d_k, non_blocking=True); out_d.copy_(d_d, non_blocking=True)
        for _ in range(2):
            once()
        barrier()
        t0 = time.perf_counter(); n = 0
        while time.perf_counter() - t0 < seconds:
            once(); once()
            su.synchronize(); sd.synchronize()
            n += 2
        dt = time.perf_counter() - t0
        barrier()
        return n * BATCH / dt
    ceiling_fps = copy_ceiling()
    # single-frame latency of the drop-in call (what Frame::ExtractORB sees): batch 1, host buffers
    ex1 = ORBextractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, width=WIDTH, height=HEIGHT, max_batch=1, device=local)
    for _ in range(5):
        check(ex1._L.orbb200_extract_host(ex1._h, host_in.ctypes.data, 1, WIDTH, WIDTH * HEIGHT, out_k.data_ptr(), out_d.data_ptr(), out_c.data_ptr(), cap))
    t0 = time.perf_counter()
    for _ in range(50):
        check(ex1._L.orbb200_extract_host(ex1._h, host_in.ctypes.data, 1, WIDTH, WIDTH * HEIGHT, out_k.data_ptr(), out_d.data_ptr(), out_c.data_ptr(), cap))
    latency_ms = (time.perf_counter() - t0) / 50 * 1e3
    ex1.close()

    t = torch.tensor([dev_ms, e2e_s * 1e3, e2e_sync_s * 1e3], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(nkp), ceiling_fps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)      # max over ranks (device-timed)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)    # the path's only "collective": a count gather
    dev_ms_max, e2e_ms_max, e2e_sync_ms_max = t.tolist()
    nkp_all, ceiling_all = int(tot[0].item()), float(tot[1].item())

    # free the VGA handles before the other workloads of the line allocate theirs
    parity_in = (outs[0][0].numpy()[:8].copy(), outs[0][1].numpy()[:8].copy(), outs[0][2].numpy()[:8].copy())
    ex.close()
    del d_frames, out_k, out_d, outs, pinned
    torch.cuda.empty_cache()

    def reduce_max(x):
        if world == 1:
            return float(x)
        v = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        return float(v.item())

    def reduce_sum(x):
        if world == 1:
            return float(x)
        v = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        dist.all_reduce(v, op=dist.ReduceOp.SUM)
        return float(v.item())

    hd = matching = None
    if args.workload == "vga":
        # every rank takes part: the shards of configs[3], configs[2] and configs[4] are timed like the headline
        # (barrier, CUDA events on the launch stream, MAX over ranks)
        hd = run_hd(local, rank, world, max(3, min(args.steps, 5)), barrier, reduce_max, reduce_sum, flush)
        if not args.no_matching:
            matching = run_matching(local, rank, world, max(3, min(args.steps, 5)), barrier, reduce_max, reduce_sum)

    if rank == 0:
        frames_total = BATCH * world * args.steps
        value = frames_total / (dev_ms_max * 1e-3)
        e2e = frames_total / (e2e_ms_max * 1e-3)
        peak, which = _peaks()
        per_stage = stage_ms / args.steps
        dom = int(np.argmax(per_stage))
        ach = ALG_BYTES[STAGES[dom]] * BATCH / (per_stage[dom] * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.workload == "hd" else "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": _config(world),
            "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": BATCH * WIDTH * HEIGHT,
                    "d2h_bytes_per_step": BATCH * (4 + cap * 60), "keypoints_per_step_per_gpu": nkp_all / world,
                    "keypoints_per_step": nkp_all,
                    "api": "StreamingExtractor (orbb200_extract_host_async, %d handles in turn)" % depth,
                    "blocking_call_value": frames_total / (e2e_sync_ms_max * 1e-3),
                    "host_copy_ceiling": {"value": ceiling_all, "unit": "frames/s", "frac": e2e / ceiling_all,
                                          "how": "the same uploads (pinned frames) and downloads (keypoints + descriptors) by all ranks at once "
                                                 "with no kernels in between, summed over ranks: what this host's PCIe paths deliver"},
                    "single_frame_latency_ms": latency_ms},
            "gpu_launches": launches,
            "kernel_ms_per_step": {s: float(per_stage[i]) for i, s in enumerate(STAGES)},
            "roofline": {"bound": "hbm", "kernel": STAGES[dom], "achieved": ach, "peak": peak, "unit": "GB/s",
                         "frac": ach / peak, "traffic": _traffic(STAGES[dom]), "peak_source": which,
                         "issue": _issue_profile({"pyramid": "k_resize", "fast": "k_fast", "quadtree": "k_quadtree", "blur": "k_blur",
                                                  "describe": "k_describe"}[STAGES[dom]]),
                         "all": {s: ALG_BYTES[s] * BATCH / (per_stage[i] * 1e-3) / 1e9 / peak for i, s in enumerate(STAGES)},
                         "whole_step": sum(ALG_BYTES.values()) * BATCH / (dev_ms_max / args.steps * 1e-3) / 1e9 / peak},
            "clocks": clocks,
        }
        if hd is not None:
            line["hd"] = hd
        if matching is not None:
            line["matching"] = matching
        if world == 1 and not args.no_matching:
            line["pipeline"] = run_pipeline(local, max(2, min(args.steps, 5)))
            line["stereo"] = run_stereo(local, max(2, min(args.steps, 5)))
        if world == 1 and not args.no_cpu_baseline:
            threads = _cpu_threads()
            sample = BATCH
            fps, kind, _ = cpu_extract_fps(frames[:sample], threads)
            line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": threads, "kind": kind,
                                    "sample": "%d of the %d frames of one step, %d host threads" % (sample, BATCH, threads),
                                    "primitives": "the reference's ORBextractor.cc over the scalar mini-cv stand-in for OpenCV (about 3x slower "
                                                  "than OpenCV's SIMD build)"}
            line["cpu_baseline"].update(cpu_single_thread(frames[:8]))
            pr = cpu_primitive_ratio(frames[:4])
            if pr:
                line["cpu_baseline"]["primitive_cost"] = pr
            line["parity"] = parity_report(frames[:8], *parity_in)
            if hd is not None:
                line["hd"]["cpu_baseline"] = cpu_hd_baseline(threads)
            if matching is not None:
                cb = cpu_matching_baselines(threads)
                for k in cb:
                    matching[k]["cpu_baseline"] = cb[k]
        _emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_hd(local, rank, world, steps, barrier, reduce_max, reduce_sum, flush):
    """configs[3]: 1024 frames 1280x720, nFeatures 2000, split contiguously over the GPUs (strong scaling).  Device-resident
    throughput (CUDA events on the launch stream, L2 flushed before every step, MAX over ranks) and the same through the
    blocking host-buffer call (uploads and downloads inside the timed region)."""
    import numpy as np
    import torch
    from weiner_slamit_v2_b200 import ORBextractor
    from weiner_slamit_v2_b200._lib import check
    w, h, n = HD["width"], HD["height"], HD["total"] // world
    ex = ORBextractor(HD["nfeatures"], SCALE, NLEVELS, INI_TH, MIN_TH, width=w, height=h, max_batch=n, device=local)
    ex.set_profiling(True)
    pinned = torch.empty((n, h, w), dtype=torch.uint8, pin_memory=True)
    pinned.numpy()[:] = _frames(1000 + rank, n, w, h, HD["distinct"])
    d_frames = pinned.cuda()
    stream = torch.cuda.ExternalStream(ex.stream, device=local)
    for _ in range(3):
        ex.extract_device(d_frames, n, w, w * h)
    ex.sync()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    stage_ms = np.zeros(5)
    barrier()
    with torch.cuda.stream(stream):
        for k in range(steps):
            flush.fill_(k & 0xff)
            ev[k][0].record()
            ex.extract_device(d_frames, n, w, w * h)
            ev[k][1].record()
            stage_ms += ex.stage_ms()
    ex.sync()
    barrier()
    dev_ms = reduce_max(sum(a.elapsed_time(b) for a, b in ev) / steps)
    cap = ex.max_keypoints
    out_k = torch.empty((n, cap, 28), dtype=torch.uint8, pin_memory=True)
    out_d = torch.empty((n, cap, 32), dtype=torch.uint8, pin_memory=True)
    out_c = torch.empty((n,), dtype=torch.int32, pin_memory=True)
    host_in = pinned.numpy()

    launches = ex.last_launches
    ex.close()
    del d_frames
    torch.cuda.empty_cache()
    # end to end: the step's frames streamed from page-locked memory in sub-batches of up to 128 through three handles in turn
    # (StreamingExtractor, orbb200_extract_host_async): the upload of one sub-batch, the kernels of the previous one and the
    # download of the one before run side by side; every frame is uploaded and every result downloaded inside the timed region
    from weiner_slamit_v2_b200 import StreamingExtractor
    sub = min(128, n)
    sx = StreamingExtractor(HD["nfeatures"], SCALE, NLEVELS, INI_TH, MIN_TH, width=w, height=h, max_batch=sub, device=local, depth=3)

    def e2e_step():
        for f0 in range(0, n, sub):
            m = min(sub, n - f0)
            sx.submit(host_in[f0:f0 + m], m, w, w * h, out_k[f0:f0 + m], out_d[f0:f0 + m], out_c[f0:f0 + m], cap)
        sx.drain()
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_ms = reduce_max((time.perf_counter() - t0) / steps * 1e3)
    sx.close()
    nkp = reduce_sum(int(out_c.sum()))
    del pinned, out_k, out_d
    torch.cuda.empty_cache()
    peak, which = _peaks()
    per_stage = stage_ms / steps
    return {"metric": HD["metric"], "workload": "1280x720 frames, nFeatures=2000, 8 levels, batch 1024 sharded over %d B200 (configs[3])" % world,
            "value": HD["total"] / (dev_ms * 1e-3), "unit": "frames/s", "n_gpus": world, "frames_per_gpu": n, "steps": steps,
            "ms_per_step": dev_ms, "scaling": "strong", "keypoints_per_step": nkp, "gpu_launches_per_step": launches,
            "kernel_ms_per_step": {s: float(per_stage[i]) for i, s in enumerate(STAGES)},
            "roofline_all": {s: HD["alg_bytes"][s] * n / (per_stage[i] * 1e-3) / 1e9 / peak for i, s in enumerate(STAGES)},
            "e2e": {"value": HD["total"] / (e2e_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": n * w * h,
                    "d2h_bytes_per_step": n * (4 + cap * 60),
                    "api": "StreamingExtractor (orbb200_extract_host_async), sub-batches of %d frames through 3 handles in turn" % sub}}


def _int_roofline(evals_per_s):
    """Matching is bound by the integer pipes, not by bytes: the denominator is the POPC issue rate
    measured on this GPU model by tools/int_peak.cu (profiles/r1k_int_peak.json) at the 3 POPC every candidate costs
    at least -- the 128-bit half distance that decides whether the candidate can matter (k_init_topk); the candidates
    that survive it (about one in ten here) cost 5 more, which the numerator does not get credit for."""
    p = os.path.join(ROOT, "profiles", "r1k_int_peak.json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        k = json.load(f)
    peak = k["popc_per_clk_per_sm"] * k["sms"] * k["sm_clock_mhz"] * 1e6 / 3.0
    return {"bound": "integer pipe (POPC)", "achieved": evals_per_s / 1e9, "peak": peak / 1e9, "unit": "G candidate evaluations/s",
            "frac": evals_per_s / peak,
            "peak_source": "profiles/r1k_int_peak.json: %.2f POPC lanes/clk/SM x %d SMs x %.0f MHz / 3 POPC per candidate (half-distance "
                           "lower bound); whole call (grid + top-4 + greedy resolve) in the numerator" % (k["popc_per_clk_per_sm"], k["sms"], k["sm_clock_mhz"])}


def _half_distance_survivors(d1, d2, far=56):
    """Fraction of the (query, candidate) pairs of two descriptor sets whose distance over the first 128 bits is below `far`
    (the candidates k_init_topk finishes to 256 bits; the rest is dropped after 3 POPC).  torch on the GPU: plumbing."""
    import torch
    lut = torch.tensor([bin(i).count("1") for i in range(256)], dtype=torch.int16, device=d1.device)
    x = torch.bitwise_xor(d1[:, None, :16], d2[None, :, :16]).long()
    return float((lut[x].sum(-1) < far).float().mean().item())


def run_matching(local, rank, world, steps, barrier, reduce_max, reduce_sum):
    """configs[2] and configs[4], split contiguously over the ranks (strong scaling): SearchForInitialization on 4096
    brute-force-shaped pairs (1000 x 1000 descriptors, ratio 0.9) and SearchByProjection on 512 frames (10k map points vs
    2000 keypoints); data resident in HBM, CUDA events on the matcher's stream after a barrier, MAX over ranks.  On one GPU
    also config 3a (reference-semantics windowed search on extracted keypoints), the brute-force shape on extracted
    descriptors, and the widened rows."""
    import ctypes as C
    import numpy as np
    import torch
    from weiner_slamit_v2_b200 import _lib
    from weiner_slamit_v2_b200._lib import FrameView, MapPointView, check
    from weiner_slamit_v2_b200.workloads import SCALE_FACTORS_8, init_pair, projection_frame

    L = _lib.load()
    dev = torch.device("cuda", local)
    out = {}

    def up(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    def tile(a, items):
        reps = (items + len(a) - 1) // len(a)
        return np.concatenate([a] * reps)[:items]

    def timed(stream, fn, steps):
        """ms per call, this rank; then MAX over ranks"""
        st = torch.cuda.ExternalStream(stream, device=local)
        for _ in range(2):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(st):
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
        torch.cuda.synchronize()
        return reduce_max(e0.elapsed_time(e1) / steps)

    # ---- configs[2]
    total, n, distinct = MATCH_PAIRS, MATCH_N, 32
    items = total // world
    pairs = [init_pair(rank * 1000 + i, n=n, brute_force=True) for i in range(distinct)]
    def fv(idx_k, idx_d):
        t = dict(n=up(np.full(items, n, np.int32)),
                 x=up(tile(np.stack([p[idx_k]["x"] for p in pairs]), items)), y=up(tile(np.stack([p[idx_k]["y"] for p in pairs]), items)),
                 o=up(tile(np.stack([p[idx_k]["octave"] for p in pairs]), items)), a=up(tile(np.stack([p[idx_k]["angle"] for p in pairs]), items)),
                 d=up(tile(np.stack([p[idx_d] for p in pairs]), items)))
        return t, FrameView(t["n"].data_ptr(), t["x"].data_ptr(), t["y"].data_ptr(), t["o"].data_ptr(), t["a"].data_ptr(), t["d"].data_ptr(), n)
    t1, v1 = fv(0, 1)
    t2, v2 = fv(2, 3)
    prev0 = up(tile(np.stack([p[4] for p in pairs]), items))
    prev = prev0.clone()
    m12 = torch.empty((items, n), dtype=torch.int32, device=dev)
    nm = torch.empty(items, dtype=torch.int32, device=dev)
    h = _lib.vp()
    check(L.orbb200_matcher_create(items, n, local, C.byref(h)))
    bounds = np.array([0, 0, 640, 480], np.float32)
    def init_step():
        prev.copy_(prev0)
        check(L.orbb200_search_for_initialization(h, items, C.byref(v1), C.byref(v2), bounds.ctypes.data, 0.9, 1, 1000,
                                                  prev.data_ptr(), m12.data_ptr(), nm.data_ptr(), 1))
    ms = timed(L.orbb200_matcher_stream(h), init_step, steps)
    acc = reduce_sum(int(nm.sum()))
    surv = _half_distance_survivors(t1["d"][0], t2["d"][0])
    out["search_for_initialization"] = {
        "workload": "4096 frame pairs split over %d GPU(s), 1000 x 1000 descriptors per pair (all octave 0, window > image), ratio 0.9, checkOri (configs[2])" % world,
        "n_gpus": world, "pairs_per_gpu": items, "scaling": "strong",
        "ms_per_step": ms, "pairs_per_s": total / ms * 1e3, "distance_evals_per_s": total * n * n / ms * 1e3,
        "accepted_matches_per_s": acc / ms * 1e3, "accepted_matches": int(acc), "gpu_launches_per_step": L.orbb200_matcher_last_launches(h),
        "half_distance_survivor_fraction": surv,
        "roofline": _int_roofline(total * n * n / ms * 1e3 / world)}
    del t1, t2, prev, prev0, m12
    L.orbb200_matcher_destroy(h)

    # ---- configs[4]
    total, nk, nmp, distinct = PROJ_FRAMES, PROJ_KP, PROJ_MP, 16
    items = total // world
    h = _lib.vp()
    check(L.orbb200_matcher_create(items, nmp, local, C.byref(h)))
    fr = [projection_frame(rank * 1000 + i, nk, nmp) for i in range(distinct)]
    tk = dict(n=up(np.full(items, nk, np.int32)), x=up(tile(np.stack([f[0]["x"] for f in fr]), items)),
              y=up(tile(np.stack([f[0]["y"] for f in fr]), items)), o=up(tile(np.stack([f[0]["octave"] for f in fr]), items)),
              d=up(tile(np.stack([f[1] for f in fr]), items)))
    kv = FrameView(tk["n"].data_ptr(), tk["x"].data_ptr(), tk["y"].data_ptr(), tk["o"].data_ptr(), None, tk["d"].data_ptr(), nk)
    keys = ["in_view", "bad", "x", "y", "xr", "level", "viewcos", "desc", "obs"]
    tm = {k: up(tile(np.stack([f[2][k] for f in fr]), items)) for k in keys}
    tm["n"] = up(np.full(items, nmp, np.int32))
    mv = MapPointView(tm["n"].data_ptr(), *[tm[k].data_ptr() for k in keys], nmp)
    kpmp = torch.empty((items, nk), dtype=torch.int32, device=dev)
    nm2 = torch.empty(items, dtype=torch.int32, device=dev)
    sf = up(SCALE_FACTORS_8)
    b2 = np.array([0, 0, 1280, 720], np.float32)
    def proj_step():
        kpmp.fill_(-1)
        check(L.orbb200_search_by_projection(h, items, C.byref(kv), None, C.byref(mv), kpmp.data_ptr(), None, sf.data_ptr(), 8,
                                             b2.ctypes.data, 0.8, 1.0, nm2.data_ptr(), 1))
    ms = timed(L.orbb200_matcher_stream(h), proj_step, steps)
    acc = reduce_sum(int(nm2.sum()))
    alg = PROJ_MP * (32 + 24) + PROJ_KP * (32 + 20) + 64 * 48 * 4 + PROJ_KP * 4        # bytes per frame
    peak, _ = _peaks()
    out["search_by_projection"] = {
        "workload": "512 frames split over %d GPU(s), 10000 projected map points vs 2000 keypoints per frame, th=1, ratio 0.8 (configs[4])" % world,
        "n_gpus": world, "frames_per_gpu": items, "scaling": "strong",
        "ms_per_step": ms, "frames_per_s": total / ms * 1e3, "map_points_per_s": total * nmp / ms * 1e3,
        "accepted_matches_per_s": acc / ms * 1e3, "accepted_matches": int(acc), "gpu_launches_per_step": L.orbb200_matcher_last_launches(h),
        "algorithmic_GB_per_s_per_gpu": alg * items / (ms * 1e-3) / 1e9, "hbm_frac": alg * items / (ms * 1e-3) / 1e9 / peak}
    if world > 1:
        L.orbb200_matcher_destroy(h)
        return out
    out.update(run_matching_extracted(local, steps, timed))
    # ---- scope row N2: SearchByProjection(CurrentFrame, LastFrame, th, bMono), the per-frame motion-model tracker
    from weiner_slamit_v2_b200._lib import LastFrameView
    from weiner_slamit_v2_b200.workloads import motion_frame
    items, nl, nk, distinct = 512, 1500, 2000, 16
    ws = [motion_frame(i, nl, nk) for i in range(distinct)]
    tk = dict(n=up(np.full(items, nk, np.int32)), x=up(tile(np.stack([w["cur"]["x"] for w in ws]), items)),
              y=up(tile(np.stack([w["cur"]["y"] for w in ws]), items)), o=up(tile(np.stack([w["cur"]["octave"] for w in ws]), items)),
              a=up(tile(np.stack([w["cur"]["angle"] for w in ws]), items)), d=up(tile(np.stack([w["cdesc"] for w in ws]), items)))
    kv = FrameView(tk["n"].data_ptr(), tk["x"].data_ptr(), tk["y"].data_ptr(), tk["o"].data_ptr(), tk["a"].data_ptr(), tk["d"].data_ptr(), nk)
    lk = ["has_mp", "outlier", "wpos", "mp_desc", "mp_obs", "last_octave", "last_angle"]
    tl = {k: up(tile(np.stack([w[k] for w in ws]), items)) for k in lk}
    tl["n"] = up(np.full(items, nl, np.int32))
    lv = LastFrameView(tl["n"].data_ptr(), *[tl[k].data_ptr() for k in lk], nl)
    Rc = up(tile(np.stack([w["Rcw"] for w in ws]), items)); tc = up(tile(np.stack([w["tcw"] for w in ws]), items))
    Kc = np.ascontiguousarray(ws[0]["K"], np.float32)
    b3 = np.array([-13.7, -9.2, 661.3, 492.8], np.float32)
    kpmp3 = torch.empty((items, nk), dtype=torch.int32, device=dev)
    nm3 = torch.empty(items, dtype=torch.int32, device=dev)
    def last_step():
        kpmp3.fill_(-1)
        check(L.orbb200_search_by_projection_last_frame(h, items, C.byref(kv), None, C.byref(lv), Rc.data_ptr(), tc.data_ptr(),
                                                        Kc.ctypes.data, 40.0, kpmp3.data_ptr(), None, sf.data_ptr(), 8,
                                                        b3.ctypes.data, 15.0, 0, 1, nm3.data_ptr(), 1))
    ms = timed(L.orbb200_matcher_stream(h), last_step, steps)
    acc = int(nm3.sum())
    out["search_by_projection_last_frame"] = {
        "workload": "512 frame pairs, 1500 last-frame map points projected into 2000 keypoints, th=15, mono, checkOri",
        "ms_per_step": ms, "frames_per_s": items / ms * 1e3, "map_points_per_s": items * nl / ms * 1e3,
        "accepted_matches_per_s": acc / ms * 1e3, "accepted_matches": acc, "gpu_launches_per_step": L.orbb200_matcher_last_launches(h)}
    # ---- scope row N2, relocalisation: SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist)
    from weiner_slamit_v2_b200._lib import KeyFrameView
    from weiner_slamit_v2_b200.workloads import relocalisation_frame
    ws = [relocalisation_frame(i, nl, nk) for i in range(distinct)]
    tk = dict(n=up(np.full(items, nk, np.int32)), x=up(tile(np.stack([w["cur"]["x"] for w in ws]), items)),
              y=up(tile(np.stack([w["cur"]["y"] for w in ws]), items)), o=up(tile(np.stack([w["cur"]["octave"] for w in ws]), items)),
              a=up(tile(np.stack([w["cur"]["angle"] for w in ws]), items)), d=up(tile(np.stack([w["cdesc"] for w in ws]), items)))
    kv = FrameView(tk["n"].data_ptr(), tk["x"].data_ptr(), tk["y"].data_ptr(), tk["o"].data_ptr(), tk["a"].data_ptr(), tk["d"].data_ptr(), nk)
    tq = {k: up(tile(np.stack([w[k] for w in ws]), items)) for k in ["wpos", "mp_desc", "mf_max", "mf_min", "kf_angle"]}
    tq["valid"] = up(tile(np.stack([(w["valid"] == 1).astype(np.uint8) for w in ws]), items))
    tq["n"] = up(np.full(items, nl, np.int32))
    qv = KeyFrameView(tq["n"].data_ptr(), tq["valid"].data_ptr(), tq["wpos"].data_ptr(), tq["mp_desc"].data_ptr(), tq["mf_max"].data_ptr(),
                      tq["mf_min"].data_ptr(), tq["kf_angle"].data_ptr(), nl)
    Rc = up(tile(np.stack([w["Rcw"] for w in ws]), items)); tc = up(tile(np.stack([w["tcw"] for w in ws]), items))
    Oc = up(tile(np.stack([w["Ow"] for w in ws]), items))
    pre = up(tile(np.stack([w["kp_mp"] for w in ws]), items))
    kpmp4 = torch.empty((items, nk), dtype=torch.int32, device=dev)
    nm4 = torch.empty(items, dtype=torch.int32, device=dev)
    def reloc_step():
        kpmp4.copy_(pre)
        check(L.orbb200_search_by_projection_keyframe(h, items, C.byref(kv), C.byref(qv), Rc.data_ptr(), tc.data_ptr(), Oc.data_ptr(),
                                                      Kc.ctypes.data, kpmp4.data_ptr(), sf.data_ptr(), 8, float(ws[0]["log_scale"]),
                                                      b3.ctypes.data, 10.0, 100, 1, nm4.data_ptr(), 1))
    ms = timed(L.orbb200_matcher_stream(h), reloc_step, steps)
    acc = int(nm4.sum())
    out["search_by_projection_keyframe"] = {
        "workload": "512 (frame, key frame) pairs, 1500 key-frame map points projected into 2000 keypoints, th=10, ORBdist=100, checkOri",
        "ms_per_step": ms, "frames_per_s": items / ms * 1e3, "map_points_per_s": items * nl / ms * 1e3,
        "accepted_matches_per_s": acc / ms * 1e3, "accepted_matches": acc, "gpu_launches_per_step": L.orbb200_matcher_last_launches(h)}
    # ---- scope row N3: SearchByBoW(pKF, F, vpMapPointMatches), tracking against the reference key frame
    from weiner_slamit_v2_b200._lib import BowView
    from weiner_slamit_v2_b200.workloads import bow_pair
    nfe, nnodes = 2000, 100
    ws = [bow_pair(i, nfe, nfe, nnodes) for i in range(distinct)]
    def bow_side(pfx):
        ns = max(len(w[pfx + "_node"]) for w in ws)
        node = np.zeros((distinct, ns), np.uint32); start = np.zeros((distinct, ns + 1), np.int32)
        for i, w in enumerate(ws):
            k = len(w[pfx + "_node"])
            node[i, :k] = w[pfx + "_node"]; start[i, :k + 1] = w[pfx + "_start"]; start[i, k + 1:] = w[pfx + "_start"][-1]
        t = dict(n=up(np.full(items, nfe, np.int32)), nn=up(tile(np.array([len(w[pfx + "_node"]) for w in ws], np.int32), items)),
                 desc=up(tile(np.stack([w[pfx + "_desc"] for w in ws]), items)), ang=up(tile(np.stack([w[pfx + "_angle"] for w in ws]), items)),
                 node=up(tile(node, items)), start=up(tile(start, items)), feat=up(tile(np.stack([w[pfx + "_feat"] for w in ws]), items)))
        if pfx == "kf":
            t["valid"] = up(tile(np.stack([(w["kf_valid"] == 1).astype(np.uint8) for w in ws]), items))
        v = BowView(t["n"].data_ptr(), t["desc"].data_ptr(), t["ang"].data_ptr(), t["valid"].data_ptr() if pfx == "kf" else None,
                    t["nn"].data_ptr(), t["node"].data_ptr(), t["start"].data_ptr(), t["feat"].data_ptr(), nfe, ns)
        return t, v
    tkf, vkf = bow_side("kf")
    tfr, vfr = bow_side("f")
    m5 = torch.empty((items, nfe), dtype=torch.int32, device=dev)
    nm5 = torch.empty(items, dtype=torch.int32, device=dev)
    def bow_step():
        check(L.orbb200_search_by_bow(h, items, C.byref(vkf), C.byref(vfr), 0.7, 1, m5.data_ptr(), nm5.data_ptr(), 1))
    ms = timed(L.orbb200_matcher_stream(h), bow_step, steps)
    acc = int(nm5.sum())
    out["search_by_bow"] = {
        "workload": "512 (key frame, frame) pairs, 2000 x 2000 features in ~100 shared vocabulary nodes, ratio 0.7, checkOri",
        "ms_per_step": ms, "frames_per_s": items / ms * 1e3, "accepted_matches_per_s": acc / ms * 1e3, "accepted_matches": acc,
        "gpu_launches_per_step": L.orbb200_matcher_last_launches(h)}
    # ---- scope row N3: SearchForTriangulation and the search of Fuse; N4: distinctive descriptors (host-buffer calls:
    # these back-end routines are called with host data; the times include the uploads)
    from weiner_slamit_v2_b200.matcher import ORBmatcher
    from weiner_slamit_v2_b200.workloads import fuse_frame, observed_descriptors, triangulation_pair
    mm = ORBmatcher(0.6, True, max_items=64, max_points=3000, device=local)
    tw = [triangulation_pair(i, 2000, 2000, 100) for i in range(64)]
    fw = [fuse_frame(i, 3000, 2000) for i in range(64)]
    obs = observed_descriptors(0, [int(v) for v in np.random.default_rng(0).integers(2, 40, 100000)])

    def host_timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            r = fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e3, r
    ms, r = host_timed(lambda: mm.search_for_triangulation_batch(tw), 3)
    mm.search_for_triangulation_batch(tw, device_reps=steps)
    out["search_for_triangulation"] = {"workload": "64 key-frame pairs, 2000 x 2000 features in ~100 shared nodes",
                                       "ms_per_step": mm.last_device_ms, "pairs_per_s": 64 / mm.last_device_ms * 1e3,
                                       "timing": "inputs and outputs resident in HBM, CUDA events on the matcher's stream",
                                       "host_buffer_call_ms_incl_python_packing": ms, "accepted_matches": int(r[0].sum())}
    ms, r = host_timed(lambda: mm.fuse_search_batch(fw, (0.0, 0.0, 640.0, 480.0), 3.0), 3)
    mm.fuse_search_batch(fw, (0.0, 0.0, 640.0, 480.0), 3.0, device_reps=steps)
    out["fuse_search"] = {"workload": "64 key frames x 3000 candidate map points vs 2000 keypoints, th=3",
                          "ms_per_step": mm.last_device_ms, "map_points_per_s": 64 * 3000 / mm.last_device_ms * 1e3,
                          "timing": "inputs and outputs resident in HBM, CUDA events on the matcher's stream",
                          "host_buffer_call_ms_incl_python_packing": ms, "fused": int(sum((b >= 0).sum() for b, _ in r))}
    off = np.zeros(len(obs) + 1, np.int32); off[1:] = np.cumsum([len(o) for o in obs])
    flat = np.ascontiguousarray(np.concatenate(obs)); bestd = np.zeros(len(obs), np.int32)
    ms, r = host_timed(lambda: check(L.orbb200_distinctive_descriptors(mm._h, len(obs), off.ctypes.data, flat.ctypes.data, int(off[-1]),
                                                                       bestd.ctypes.data, None, 0)), 3)
    d_off, d_flat, d_best = up(off), up(flat), torch.empty(len(obs), dtype=torch.int32, device=dev)
    dms = timed(L.orbb200_matcher_stream(mm._h), lambda: check(L.orbb200_distinctive_descriptors(
        mm._h, len(obs), d_off.data_ptr(), d_flat.data_ptr(), int(off[-1]), d_best.data_ptr(), None, 1)), steps)
    assert np.array_equal(d_best.cpu().numpy(), bestd), "device-resident call differs from the host-buffer call"
    out["distinctive_descriptors"] = {"workload": "100000 map points with 2..39 observations each (%d descriptors)" % int(off[-1]),
                                      "ms_per_step": dms, "map_points_per_s": 100000 / dms * 1e3,
                                      "timing": "inputs and outputs resident in HBM, CUDA events on the matcher's stream",
                                      "host_buffer_call_ms": ms}
    # ---- scope row N4: Frame::ComputeBoW (DBoW2 transform), descriptors and outputs resident in HBM
    from weiner_slamit_v2_b200.matcher import Vocabulary
    from weiner_slamit_v2_b200.workloads import synthetic_vocabulary, vocabulary_features
    voc = synthetic_vocabulary(0, 10, 5)
    V = Vocabulary(voc, device=local)
    nfb = 2000
    dfeat = up(tile(np.stack([vocabulary_features(i, voc, nfb) for i in range(distinct)]), items))
    nb_ = up(np.full(items, nfb, np.int32))
    o_i = [torch.empty(items, dtype=torch.int32, device=dev) for _ in range(2)]
    o_u = [torch.empty((items, nfb), dtype=torch.int32, device=dev) for _ in range(3)]
    o_d = torch.empty((items, nfb), dtype=torch.float64, device=dev)
    o_s = torch.empty((items, nfb + 1), dtype=torch.int32, device=dev)
    def bowt_step():
        check(L.orbb200_bow_transform(h, V._h, items, nb_.data_ptr(), dfeat.data_ptr(), nfb, 4, o_i[0].data_ptr(), o_u[0].data_ptr(),
                                      o_d.data_ptr(), o_i[1].data_ptr(), o_u[1].data_ptr(), o_s.data_ptr(), o_u[2].data_ptr(), 1))
    ms = timed(L.orbb200_matcher_stream(h), bowt_step, steps)
    out["bow_transform"] = {
        "workload": "512 frames x 2000 descriptors through a synthetic vocabulary (%d nodes, k <= 10, L = 5), levelsup 4" % len(voc["parent"]),
        "ms_per_step": ms, "frames_per_s": items / ms * 1e3, "descriptors_per_s": items * nfb / ms * 1e3,
        "words": int(o_i[0].sum()), "gpu_launches_per_step": L.orbb200_matcher_last_launches(h)}
    V.close()
    mm.close()
    L.orbb200_matcher_destroy(h)
    return out


def run_matching_extracted(local, steps, timed):
    """SearchForInitialization on keypoints and descriptors that come out of the extractor (256 frame pairs: frame i and its
    copy shifted by (4, 2) px, extracted, undistorted with the reference's camera and gridded on the device, tiled to 4096
    pairs): (3a) the reference's semantics as Tracking calls it -- octave-0 queries, +-100 px window, greedy steal -- and
    (3b') the brute-force shape of configs[2] on these REAL descriptors (every octave forced to 0, window > image), where
    true matches make far more candidates survive the half-distance test than on the unrelated descriptors of configs[2]."""
    import ctypes as C
    import numpy as np
    import torch
    from weiner_slamit_v2_b200 import _lib
    from weiner_slamit_v2_b200._lib import FrameView, check
    from weiner_slamit_v2_b200.pipeline import InitializationPipeline
    L = _lib.load()
    pairs, items = 256, MATCH_PAIRS
    f1 = _frames(3, pairs)
    f2 = np.stack([np.roll(f, (2, 4), (0, 1)) for f in f1])
    pipe = InitializationPipeline(max_pairs=pairs, device=local)
    pipe.run(torch.from_numpy(f1).cuda(), torch.from_numpy(f2).cuda(), pairs)
    pipe.sync()
    cap, rep = pipe.cap, items // pairs
    side = [{k: v[:pairs].repeat((rep,) + (1,) * (v.dim() - 1)).contiguous() for k, v in sd.items() if k != "kps"} for sd in pipe.side]
    bounds = pipe.bounds.copy()
    pipe.close()
    n1 = side[0]["n"].float().mean().item(); n2 = side[1]["n"].float().mean().item()
    lvl0 = (side[0]["oct"][:pairs] == 0) & (torch.arange(cap, device=side[0]["oct"].device)[None, :] < side[0]["n"][:pairs, None])
    q0 = lvl0.sum(1).float().mean().item()
    h = _lib.vp()
    check(L.orbb200_matcher_create(items, cap, local, C.byref(h)))
    prev0 = torch.stack([side[0]["x"], side[0]["y"]], dim=-1).contiguous()
    prev = prev0.clone()
    m12 = torch.empty((items, cap), dtype=torch.int32, device=prev.device)
    nm = torch.empty(items, dtype=torch.int32, device=prev.device)
    out = {}

    def view(sd, octv):
        return FrameView(sd["n"].data_ptr(), sd["x"].data_ptr(), sd["y"].data_ptr(), octv.data_ptr(), sd["ang"].data_ptr(), sd["desc"].data_ptr(), cap)

    def run(v1, v2, window):
        def step():
            prev.copy_(prev0)
            check(L.orbb200_search_for_initialization(h, items, C.byref(v1), C.byref(v2), bounds.ctypes.data, 0.9, 1, window,
                                                      prev.data_ptr(), m12.data_ptr(), nm.data_ptr(), 1))
        ms = timed(L.orbb200_matcher_stream(h), step, steps)
        return ms, int(nm.sum())
    ms, acc = run(view(side[0], side[0]["oct"]), view(side[1], side[1]["oct"]), 100)
    out["search_for_initialization_extracted_3a"] = {
        "workload": "4096 pairs of extracted frames (~%.0f and ~%.0f keypoints, ~%.0f octave-0 queries per pair), reference semantics: "
                    "octave 0 only, +-100 px window, ratio 0.9, checkOri (SURVEY 8(d) config 3a), matcher only" % (n1, n2, q0),
        "ms_per_step": ms, "pairs_per_s": items / ms * 1e3, "queries_per_s": items * q0 / ms * 1e3,
        "accepted_matches_per_s": acc / ms * 1e3, "accepted_matches": acc, "gpu_launches_per_step": L.orbb200_matcher_last_launches(h)}
    z = [torch.zeros_like(sd["oct"]) for sd in side]
    ms, acc = run(view(side[0], z[0]), view(side[1], z[1]), 1000)
    ev = float((side[0]["n"].double() * side[1]["n"].double()).sum().item())
    n0, m0 = int(side[0]["n"][0]), int(side[1]["n"][0])
    out["search_for_initialization_extracted_brute"] = {
        "workload": "the same 4096 pairs with every octave forced to 0 and window > image: ~%.0f x %.0f REAL descriptors per pair, ratio 0.9" % (n1, n2),
        "ms_per_step": ms, "pairs_per_s": items / ms * 1e3, "distance_evals_per_s": ev / ms * 1e3,
        "accepted_matches_per_s": acc / ms * 1e3, "accepted_matches": acc,
        "half_distance_survivor_fraction": _half_distance_survivors(side[0]["desc"][0, :n0], side[1]["desc"][0, :m0]),
        "roofline": _int_roofline(ev / ms * 1e3)}
    L.orbb200_matcher_destroy(h)
    return out


def run_pipeline(local, steps):
    """Device-resident front end: for 128 frame pairs, extract both frames (256 extractions),
    undistort + SoA + grid, SearchForInitialization -- no host round trip between the stages."""
    import numpy as np
    import torch
    from weiner_slamit_v2_b200.pipeline import InitializationPipeline
    pairs = 128
    base = _frames(7, 16)
    f1 = np.concatenate([base] * (pairs // 16))
    f2 = np.stack([np.roll(f, (2, 4), (0, 1)) for f in f1])
    pipe = InitializationPipeline(max_pairs=pairs, device=local)
    d1, d2 = torch.from_numpy(f1).cuda(), torch.from_numpy(f2).cuda()
    for _ in range(2):
        pipe.run(d1, d2, pairs)
    pipe.sync()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        nm, _, _ = pipe.run(d1, d2, pairs)
    pipe.sync()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    acc = int(nm[:pairs].sum())
    pipe.close()
    return {"workload": "128 frame pairs 640x480 (second frame = first shifted by (4,2) px): 2 x extraction, undistort with the "
                        "reference's camera, grid, SearchForInitialization(window 100, ratio 0.9), all on the device",
            "ms_per_step": dt * 1e3, "pairs_per_s": pairs / dt, "frames_per_s": 2 * pairs / dt, "accepted_matches": acc}


def run_stereo(local, steps):
    """Scope row N4: the stereo Frame constructor for 128 rectified pairs on the device: two extractions per pair
    (256 frames), keypoint views, Frame::ComputeStereoMatches reading both pyramids in place."""
    import numpy as np
    import torch
    from weiner_slamit_v2_b200.frames import stereo_right_frame
    from weiner_slamit_v2_b200.pipeline import StereoPipeline
    pairs = 128
    base = _frames(11, 16)
    rb = np.stack([stereo_right_frame(f, i) for i, f in enumerate(base)])
    left = np.concatenate([base] * (pairs // 16)); right = np.concatenate([rb] * (pairs // 16))
    pipe = StereoPipeline(max_pairs=pairs, device=local)
    dl, dr = torch.from_numpy(left).cuda(), torch.from_numpy(right).cuda()
    for _ in range(2):
        pipe.run(dl, dr, pairs)
    pipe.sync()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        nm, _, _ = pipe.run(dl, dr, pairs)
    pipe.sync()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    # the stereo matcher alone, CUDA events on its stream
    L = pipe.L
    st = torch.cuda.ExternalStream(L.orbb200_matcher_stream(pipe.m), device=local)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import ctypes as C
    from weiner_slamit_v2_b200 import _lib
    from weiner_slamit_v2_b200._lib import FrameView, check
    views = [FrameView(s["n"].data_ptr(), s["x"].data_ptr(), s["y"].data_ptr(), s["oct"].data_ptr(), s["ang"].data_ptr(),
                       s["desc"].data_ptr(), pipe.cap) for s in pipe.side]
    lp, rp = pipe.exl.pyramid_view(), pipe.exr.pyramid_view()
    def stereo_only():
        check(L.orbb200_compute_stereo_matches(pipe.m, pairs, C.byref(views[0]), C.byref(views[1]), C.byref(lp), C.byref(rp),
                                               pipe.scale.ctypes.data, pipe.inv_scale.ctypes.data, len(pipe.scale), pipe.mb, pipe.mbf,
                                               pipe.u_right.data_ptr(), pipe.depth.data_ptr(), pipe.nm.data_ptr(),
                                               _lib.DEVICE_VIEWS | _lib.DEVICE_PYRAMIDS))
    stereo_only()
    torch.cuda.synchronize()
    with torch.cuda.stream(st):
        e0.record()
        for _ in range(steps):
            stereo_only()
        e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    acc = int(nm[:pairs].sum())
    pipe.close()
    return {"workload": "128 rectified stereo pairs 640x480 (right = left shifted by a per-band disparity of 1..48 px): 2 x extraction, "
                        "keypoint views, ComputeStereoMatches (row bands, Hamming, 11x11 SAD over +-5 px, parabola, median cut), all on the device",
            "ms_per_step": dt * 1e3, "pairs_per_s": pairs / dt, "stereo_matcher_ms": ms, "stereo_matcher_pairs_per_s": pairs / ms * 1e3,
            "stereo_matches": acc}


def _issue_profile(kernel):
    """Issue-slot use and instruction count of `kernel` from the committed ncu --set full summary: the path is
    integer/byte work bound by instruction issue, so this is the number that explains the HBM fraction."""
    import glob
    for p in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_full_summary.json")), reverse=True):      # newest round / step first
        with open(p) as f:
            rows = [e for e in json.load(f) if e["kernel"] in (kernel, kernel + "2", kernel + "3")]
        if rows:
            e = max(rows, key=lambda r: r["inst_executed"])       # (FAST is two launches: the one that holds nearly all cells)
            return {"issue_slots_busy_pct": e["issue_active_pct"], "warp_instructions_per_launch": sum(r["inst_executed"] for r in rows),
                    "dram_pct_of_peak": e["dram_pct"], "source": "profiles/" + os.path.basename(p)}
    return None


def _traffic(stage):
    """dram bytes per launch from the committed ncu --set full capture, if one exists."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get(stage)
    return None


_REAL_STDOUT = None


def _quiet_stdout():
    """stdout carries exactly one JSON line: anything a library prints there (NCCL's version banner ...) is sent to
    stderr by pointing fd 1 at fd 2 for the run; _emit() writes the result to the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--stream-depth", type=int, default=3, help="handles the end-to-end leg alternates between")
    ap.add_argument("--no-matching", action="store_true", help="skip the Hamming-matching configs (configs[2], configs[4])")
    ap.add_argument("--workload", default="vga", choices=["vga", "hd"],
                    help="vga = configs[1] (the headline, default); hd = configs[3]: 1280x720, 2000 kp, batch 1024 sharded over the GPUs")
    args = ap.parse_args()
    if args.workload == "hd":
        global WIDTH, HEIGHT, NFEAT, BATCH, WORKLOAD, ALG_BYTES, METRIC
        WIDTH, HEIGHT, NFEAT = 1280, 720, 2000
        METRIC = "ORB frames/s (1280x720, 2000 kp, 8 lvl)"
        BATCH = 1024 // max(1, args.gpus)          # strong scaling: the 1024-frame batch is split over the GPUs
        WORKLOAD = "1280x720 frames, nFeatures=2000, 8 levels, batch 1024 sharded over %d B200 (configs[3])" % args.gpus
        ALG_BYTES = {"pyramid": 2781331 + 1931488, "fast": 2853088 + 160000, "quadtree": 160000 + 32000,
                     "blur": 5706176, "describe": 1498000 + 1024000 + 120000}
        args.no_matching = True
    if not (args.gpus > 1 and "WORLD_SIZE" not in os.environ):
        _quiet_stdout()                           # (the torchrun re-launch below passes the children's stdout through)
    if args.impl == "reference":
        return run_reference(args)
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_ours(args)


if __name__ == "__main__":
    main()
