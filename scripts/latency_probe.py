"""Where a single so_register call spends its time: the same cfg1 / cfg2 scan registered with different caps on the ICP and LM
trip counts (device time between the library's own events, median of 30), under the WHILE-graph and the unrolled schedule.

    python scripts/latency_probe.py            # prints one JSON line
"""
import json
import os
import subprocess
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def probe():
    from superodom_b200 import api, synth
    out = {"schedule": "unrolled" if os.environ.get("SO_NO_COND_GRAPH") else "while-graph", "fused_lm": not os.environ.get("SO_NO_FUSED_LM")}
    for name, cap in (("cfg1", 2000), ("cfg1", 0), ("cfg2", 0)):
        c = synth.make_case(name)
        ctx = api.Context(max_map_points=len(c["map_xyzi"]) + 1024, max_scan_points=len(c["scan_xyzi"]), plane_res=0.2)
        ctx.map_set_points(c["map_xyzi"])
        for iters, lm in ((1, 1), (1, 4), (2, 4), (5, 4)):
            for _ in range(4):
                r = ctx.register(c["scan_xyzi"], c["pose_prior"], iters, cap, lm_max_iterations=lm)
            dev, wall = [], []
            for _ in range(30):
                t = time.perf_counter()
                r = ctx.register(c["scan_xyzi"], c["pose_prior"], iters, cap, lm_max_iterations=lm)
                wall.append((time.perf_counter() - t) * 1e3)
                dev.append(r.time_ms)
            out[f"{name}_cap{cap}_icp{iters}_lm{lm}"] = {"device_ms": round(float(np.median(dev)), 4), "wall_ms": round(float(np.median(wall)), 4),
                                                      "icp_iterations": int(r.n_iterations), "lm_steps": [int(v) for v in r.iter_lm_steps[:r.n_iterations]]}
        ctx.close()
    return out


if __name__ == "__main__":
    if "--child" in sys.argv:
        print(json.dumps(probe()))
    else:
        res = []
        for env in ({}, {"SO_NO_COND_GRAPH": "1"}, {"SO_NO_FUSED_LM": "1"}):
            p = subprocess.run([sys.executable, __file__, "--child"], env=dict(os.environ, **env), capture_output=True, text=True)
            res.append(json.loads(p.stdout.strip().splitlines()[-1]) if p.returncode == 0 else {"error": p.stderr[-400:]})
        print(json.dumps(res))
