"""Profiling target for the matcher configs (developer tool): runs bench.run_matching once on one GPU."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench


def barrier():
    torch.cuda.synchronize()


print(json.dumps(bench.run_matching(0, 0, 1, 2, barrier, float, float)))
