"""Per-kernel-class timing on a B-scan cfg2 batch (profiling mode: CUDA events around every launch with work)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from superodom_b200 import api
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
map_xyzi, scans, priors, truths = bench.make_inputs(0, B)
n_points = np.array([len(s) for s in scans], np.uint32)
flat = np.ascontiguousarray(np.concatenate(scans, 0))
ctx = api.Context(max_map_points=len(map_xyzi) + 1024, max_scan_points=int(n_points.max()), max_batch=B, plane_res=0.2)
ctx.map_set_points(map_xyzi)
for _ in range(3):
    res = ctx.register_batch(flat, n_points, priors, 20, 0, skip_map_checks=True)
t = time.perf_counter()
for _ in range(5):
    res = ctx.register_batch(flat, n_points, priors, 20, 0, skip_map_checks=True)
dt = (time.perf_counter() - t) / 5
ctx.profile_enable(True)
for k in range(6): ctx.profile_get(k, reset=True)
res = ctx.register_batch(flat, n_points, priors, 20, 0, skip_map_checks=True)
ctx.profile_enable(False)
err = np.abs(np.array([list(r.pose) for r in res])[:, :3] - truths[:, :3]).max()
print(f"{os.environ.get('TAG','')} B={B} host-timed step {dt*1e3:.2f} ms -> {B/dt:.0f} scans/s e2e | err {err:.4f} | " +
      " | ".join(f"class{k}: {ctx.profile_get(k)[0]:.3f} ms / {ctx.profile_get(k)[1]} launches" for k in range(6)))
