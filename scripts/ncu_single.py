"""Small fixed workload for ncu: one cfg1 scan (cap 2000) registered a few times through so_register."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from superodom_b200 import api, synth
c = synth.make_case("cfg1")
ctx = api.Context(max_map_points=len(c["map_xyzi"]) + 1024, max_scan_points=len(c["scan_xyzi"]), plane_res=0.2)
ctx.map_set_points(c["map_xyzi"])
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    r = ctx.register(c["scan_xyzi"], c["pose_prior"], 5, 2000)
print("iters", r.n_iterations, "lm", list(r.iter_lm_steps[:r.n_iterations]))
