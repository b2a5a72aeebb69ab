"""Profiling target for the matcher configs (developer tool): runs bench.run_matching once."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
print(json.dumps(bench.run_matching(0, 2)))
