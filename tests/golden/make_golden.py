"""Regenerate tests/golden/tiny_case.npz.  Run HERE (the container that has /root/reference) so that the oracle/_ref
build -- the reference's own octree.h/nanoflann.h compiled verbatim -- is in the loop:
    python tests/golden/make_golden.py
The fixture stores inputs and oracle outputs for one small case (the reference itself ships no golden vectors)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from superodom_b200 import synth  # noqa: E402

O.build()
assert O.has_ref_octree(), "needs the oracle/_ref build"
scene, map_xyzi = synth.make_map(6.0, 0.2)
pose_true = synth.random_sensor_pose(scene, 900, 3.0)
scan = synth.make_scan(scene, "vlp16", pose_true, 1000)[::6]
prior = synth.perturb_pose(pose_true, 2000)
m = O.OracleMap(map_xyzi)
corr, ho, hr = m.correspond(scan, prior, 0.2, 0, 0)
H, g, cost, nok = O.evaluate(corr, prior, 0.2)
r0 = m.register(scan, prior, 0.2, 5, 0, knn_mode=0)
r2 = m.register(scan, prior, 0.2, 5, 0, knn_mode=2)
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tiny_case.npz")
np.savez_compressed(out, map_xyzi=map_xyzi, scan_xyzi=scan, pose_true=pose_true, pose_prior=prior, plane_res=0.2, max_iterations=5,
                    status=corr["status"], nn=corr["nn"], n=corr["n"], d=corr["d"], w=corr["w"], hist_obs=ho, hist_rej=hr,
                    H=H, g=g, cost=cost, pose_exact=np.array(r0.pose), n_iterations_exact=r0.n_iterations,
                    pose_ref_octree=np.array(r2.pose), cov_exact=np.array(r0.cov), pos_err=r0.pos_err, ori_err_deg=r0.ori_err_deg)
print("wrote", out, os.path.getsize(out), "bytes; map", map_xyzi.shape, "scan", scan.shape,
      "| pose delta exact-vs-ref-octree", np.abs(np.array(r0.pose) - np.array(r2.pose)).max())
