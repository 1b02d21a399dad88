"""Live-SLAM loop timing: LidarSLAM::Localization as the node drives it (LidarSlam.cpp:107-171) -- per scan, register against the
CURRENT map (so_register) and then insert the registered scan into it (transformAndAddToMap -> LocalMap::addSurfPointCloud,
LidarSlam.cpp:60-80, LocalMap.h:591-645 = so_map_add_scan).  The map starts as the 1 M-point warehouse map (localization-mode
prior map, laserMapping.cpp:163-173) and keeps growing / being re-filtered by the scans themselves.

    python tests/tools/live_loop.py [--cpu] [--scans 60]

Used by bench.py's `live` block; the CPU comparator (oracle registration + numpy insert restatement, one thread) is the
cpu_baseline leg of that block and runs on a bounded sample of the same scans.
"""
from __future__ import annotations

import json
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def trajectory(synth, scene, n: int, step: float = 0.35, seed: int = 5):
    """n ground-truth poses walking through free space: forward steps, turning away from obstacles."""
    rng = np.random.default_rng(seed)
    p = synth.random_sensor_pose(scene, 900 + seed, 10.0)[:3]
    yaw = rng.uniform(-math.pi, math.pi)
    out = []
    for _ in range(n):
        for _try in range(24):
            q = p + step * np.array([math.cos(yaw), math.sin(yaw), 0.0])
            if scene.free(q):
                p = q
                break
            yaw += math.radians(25.0)
        yaw += rng.normal(0, math.radians(1.5))
        rp = rng.uniform(-0.02, 0.02, size=2)
        qt = synth.quat_mul(synth.quat_from_rotvec(np.array([0, 0, yaw])), synth.quat_from_rotvec(np.array([rp[0], rp[1], 0.0])))
        out.append(np.concatenate([p, qt / np.linalg.norm(qt)]))
    return out


def run(api, synth, device: int = 0, n_scans: int = 60, cpu: bool = False, sensor: str = "os1_128", iters: int = 5, cap: int = 2000,
        plane_res: float = 0.2, prefilter: bool = True):
    scene, map0 = synth.make_map_for("cfg2")
    poses = trajectory(synth, scene, n_scans)
    scans = [synth.make_scan(scene, sensor, T, 8000 + i) for i, T in enumerate(poses)]
    priors = [synth.perturb_pose(T, 9000 + i, dt=0.05, dth_deg=0.5) for i, T in enumerate(poses)]
    ctx = api.Context(device=device, max_map_points=4 << 20, max_scan_points=max(len(s) for s in scans), plane_res=plane_res)
    ctx.map_add_surf(map0)
    lr, pr = 0.1, plane_res
    t_reg, t_ins, t_pre, n_act, errs = [], [], [], [], []
    est = []
    for i in range(n_scans):
        s = scans[i]
        t0 = time.perf_counter()
        if prefilter:                                   # laserMapping::adjustVoxelSize: VoxelGrid(planeRes) on the scan; the result stays on the device
            n_s, lr, pr, _ = ctx.scan_prefilter(s, lr, pr, auto_voxel_size=False, download=False)
            t1 = time.perf_counter()
            r = ctx.register_prefiltered(priors[i], iters, cap)
        else:
            n_s = len(s)
            t1 = time.perf_counter()
            r = ctx.register(s, priors[i], iters, cap)
        t2 = time.perf_counter()
        ctx.map_add_registered_scan(np.array(r.pose))
        t3 = time.perf_counter()
        if i >= 5:                                      # first scans warm caches / graphs
            t_pre.append((t1 - t0) * 1e3); t_reg.append((t2 - t1) * 1e3); t_ins.append((t3 - t2) * 1e3)
        n_act.append(n_s)
        est.append(np.array(r.pose))
        errs.append(float(np.linalg.norm(np.array(r.pose)[:3] - poses[i][:3])))
    out = {"workload": f"live loop: {sensor} scans along a {n_scans}-scan trajectory, scan VoxelGrid({plane_res}) pre-filter, so_register_prefiltered "
                       f"({iters} ICP iterations max, max_surface_features {cap}) + so_map_add_registered_scan per scan (scan resident on the device throughout), map starts at {len(map0)} points",
           "scans": n_scans, "points_per_scan_after_prefilter": float(np.mean(n_act)), "map_points_end": int(ctx.map_size()),
           "ms_prefilter_median": float(np.median(t_pre)), "ms_register_median": float(np.median(t_reg)), "ms_insert_median": float(np.median(t_ins)),
           "ms_per_scan_median": float(np.median(np.array(t_pre) + np.array(t_reg) + np.array(t_ins))),
           "max_pos_err_vs_truth_m": float(max(errs))}
    out["scans_per_s"] = 1e3 / out["ms_per_scan_median"]
    if cpu:
        from oracle import oracle as O
        # CPU path on one thread for a few scans against the same starting map: voxel filter (numpy restatement), oracle registration
        # with the reference octree, numpy insert restatement
        ref = O.has_ref_octree()
        m = map0.copy()
        tc = []
        for i in range(min(3, n_scans)):
            t0 = time.perf_counter()
            s = O.adjust_voxel_size_numpy(scans[i], 0.1, plane_res, auto_voxel_size=False)[0] if prefilter else scans[i]
            om = O.OracleMap(m, ref_octree=ref)             # addSurfPointCloud rebuilds the touched blocks' octrees every scan
            ro = om.register(s, priors[i], plane_res, iters, cap, knn_mode=2 if ref else 0, n_threads=1, skip_map_checks=True)
            m = O.map_insert_numpy(m, O.transform_scan_numpy(s, np.array(ro.pose)), plane_res)
            tc.append((time.perf_counter() - t0) * 1e3)
        out["cpu_1thread_ms_per_scan"] = float(np.median(tc))
        out["cpu_note"] = "oracle registration (reference octree verbatim, rebuilt per scan as addSurfPointCloud does) + numpy restatements of the voxel filters, 1 thread"
        out["speedup"] = out["cpu_1thread_ms_per_scan"] / out["ms_per_scan_median"]
    ctx.close()
    return out


if __name__ == "__main__":
    from superodom_b200 import api, synth
    n = int(sys.argv[sys.argv.index("--scans") + 1]) if "--scans" in sys.argv else 60
    print(json.dumps(run(api, synth, n_scans=n, cpu="--cpu" in sys.argv)))
