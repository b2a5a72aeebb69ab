"""Aggregate host<->device copy bandwidth of this box with 1..N ranks copying AT THE SAME TIME from / to page-locked host
memory (no kernels): the ceiling of every end-to-end number bench.py reports at N GPUs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29541 tools/pcie_aggregate.py
    (or plain `python tools/pcie_aggregate.py` for one rank)

Rank 0 prints one JSON line: per-rank and aggregate GB/s for H2D alone, D2H alone and both directions at once, on the
transfer sizes of one bench step (78.6 MB up, 15.7 MB down) and on 256 MiB blocks."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    try:
        from bench import _bind_to_gpu_numa_node
        _bind_to_gpu_numa_node(local)
    except Exception:
        pass
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(up_bytes, down_bytes, seconds=1.5):
        hu = torch.empty(max(up_bytes, 1), dtype=torch.uint8, pin_memory=True); du = torch.empty(max(up_bytes, 1), dtype=torch.uint8, device="cuda")
        hd = torch.empty(max(down_bytes, 1), dtype=torch.uint8, pin_memory=True); dd = torch.empty(max(down_bytes, 1), dtype=torch.uint8, device="cuda")
        su, sd = torch.cuda.Stream(), torch.cuda.Stream()
        def once():
            if up_bytes:
                with torch.cuda.stream(su):
                    du.copy_(hu, non_blocking=True)
            if down_bytes:
                with torch.cuda.stream(sd):
                    hd.copy_(dd, non_blocking=True)
        for _ in range(3):
            once()
        barrier()
        t0 = time.perf_counter(); n = 0
        while time.perf_counter() - t0 < seconds:
            for _ in range(4):
                once()
            su.synchronize(); sd.synchronize()
            n += 4
        dt = time.perf_counter() - t0
        v = torch.tensor([n * up_bytes / dt / 1e9, n * down_bytes / dt / 1e9], dtype=torch.float64, device="cuda")
        allv = [torch.zeros_like(v) for _ in range(world)]
        if world > 1:
            dist.all_gather(allv, v)
        else:
            allv = [v]
        barrier()
        per = [[round(float(x[0]), 2), round(float(x[1]), 2)] for x in allv]
        return {"h2d_GBs_per_rank": [p[0] for p in per], "d2h_GBs_per_rank": [p[1] for p in per],
                "h2d_GBs_total": round(sum(p[0] for p in per), 2), "d2h_GBs_total": round(sum(p[1] for p in per), 2)}

    step_up, step_down = 256 * 640 * 480, 256 * (4 + 1064 * 60)
    out = {"ranks": world, "gpu": torch.cuda.get_device_name(local), "host_cpus": len(os.sched_getaffinity(0)),
           "step_sized": {"h2d_only": run(step_up, 0), "d2h_only": run(0, step_down), "both": run(step_up, step_down)},
           "256MiB": {"h2d_only": run(256 << 20, 0), "d2h_only": run(0, 256 << 20), "both": run(256 << 20, 256 << 20)},
           "how": "pinned host buffers, cudaMemcpyAsync on two streams per rank, all ranks copying concurrently between barriers, wall clock per rank"}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
