"""Small fixed workload for ncu: cfg2 map, one batch of 8 scans, two registrations (first warms up)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from superodom_b200 import api
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
map_xyzi, scans, priors, truths = bench.make_inputs(0, B)
n_points = np.array([len(s) for s in scans], np.uint32)
flat = np.ascontiguousarray(np.concatenate(scans, 0))
ctx = api.Context(max_map_points=len(map_xyzi) + 1024, max_scan_points=int(n_points.max()), max_batch=B, plane_res=0.2)
ctx.map_set_points(map_xyzi)
for _ in range(2):
    res = ctx.register_batch(flat, n_points, priors, 20, 0, skip_map_checks=True)
print("iters", [r.n_iterations for r in res], "err", np.abs(np.array([list(r.pose) for r in res])[:, :3] - truths[:, :3]).max())
