"""Single-scan registration latency through so_register (host scan in, pose out), beside the CPU oracle on one thread.

    python tests/tools/latency.py            # cfg1 (VLP-16, 28 800 pts, cap 2000, 5 its) and cfg2 (OS1-128, 131 072 pts, 20 its)
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from superodom_b200 import api, synth  # noqa: E402

out = {}
for name in ("cfg1", "cfg1_uncapped", "cfg2"):
    c = synth.make_case(name)
    cfg = c["cfg"]
    ctx = api.Context(max_map_points=len(c["map_xyzi"]) + 1024, max_scan_points=len(c["scan_xyzi"]), plane_res=cfg["plane_res"])
    ctx.map_set_points(c["map_xyzi"])
    for _ in range(3):
        r = ctx.register(c["scan_xyzi"], c["pose_prior"], cfg["max_iterations"], cfg["max_surface_features"])
    wall, dev = [], []
    for _ in range(20):
        t = time.perf_counter()
        r = ctx.register(c["scan_xyzi"], c["pose_prior"], cfg["max_iterations"], cfg["max_surface_features"])
        wall.append((time.perf_counter() - t) * 1e3)
        dev.append(r.time_ms)
    e = {"points": len(c["scan_xyzi"]), "map_points": len(c["map_xyzi"]), "icp_iterations": int(r.n_iterations),
         "gpu_wall_ms_median": float(np.median(wall)), "gpu_device_ms_median": float(np.median(dev))}
    if "--cpu" in sys.argv:
        from oracle import oracle as O
        om = O.OracleMap(c["map_xyzi"], ref_octree=O.has_ref_octree())
        t = time.perf_counter()
        ro = om.register(c["scan_xyzi"], c["pose_prior"], cfg["plane_res"], cfg["max_iterations"], cfg["max_surface_features"],
                         knn_mode=2 if O.has_ref_octree() else 0, n_threads=1)
        e["cpu_oracle_1thread_ms"] = (time.perf_counter() - t) * 1e3
        e["speedup_wall"] = e["cpu_oracle_1thread_ms"] / e["gpu_wall_ms_median"]
        e["pose_dpos_m"] = float(np.abs(np.array(r.pose)[:3] - np.array(ro.pose)[:3]).max())
    out[name] = e
    ctx.close()
print(json.dumps(out))
