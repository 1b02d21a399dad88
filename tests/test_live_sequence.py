"""Live-SLAM sequence: what the node does per scan -- Localization() against the CURRENT map, then transformAndAddToMap() of the
registered scan (LidarSlam.cpp:107-171, LocalMap.h:591-645) -- 200 times in a row, the map growing from the scans themselves.

GPU: so_register + so_map_add_registered_scan (the scan stays on the device between the two).  Oracle: the restated
registration against the numpy map + the numpy restatement of the insert.  Per scan the poses must agree to the north-star
tolerance (1e-4 m / 1e-4 rad; observed ~1e-12), and after the last scan the two maps must be BIT-equal: 199 chained voxel-filter
inserts with no drift of a single float.  (The numpy insert is driven with the GPU pose of each scan: a 1e-15 pose difference
may round one transformed coordinate differently, which is not what this test is about.)
"""
import numpy as np
import pytest

from conftest import quat_angle

pytestmark = pytest.mark.gpu

N_SCANS = 200


def _walk(synth, scene, n, step=0.12, seed=11):
    import math
    rng = np.random.default_rng(seed)
    p = synth.random_sensor_pose(scene, 900 + seed, 6.0)[:3]
    yaw = rng.uniform(-math.pi, math.pi)
    out = []
    for _ in range(n):
        for _try in range(24):
            q = p + step * np.array([math.cos(yaw), math.sin(yaw), 0.0])
            if scene.free(q):
                p = q
                break
            yaw += math.radians(25.0)
        yaw += rng.normal(0, math.radians(2.0))
        rp = rng.uniform(-0.02, 0.02, size=2)
        qt = synth.quat_mul(synth.quat_from_rotvec(np.array([0, 0, yaw])), synth.quat_from_rotvec(np.array([rp[0], rp[1], 0.0])))
        out.append(np.concatenate([p, qt / np.linalg.norm(qt)]))
    return out


def test_live_sequence_200_scans_matches_the_oracle_sequence(gpu_api, oracle_mod):
    from superodom_b200 import synth
    cfg = synth.CONFIGS["cfg1"]
    scene = synth.make_scene(cfg["half_extent"], seed=77)
    poses = _walk(synth, scene, N_SCANS)
    scans = [synth.make_scan(scene, "vlp16", T, 6000 + i) for i, T in enumerate(poses)]
    priors = [synth.perturb_pose(T, 6500 + i, dt=0.05, dth_deg=0.5) for i, T in enumerate(poses)]
    ctx = gpu_api.Context(max_map_points=1 << 21, max_scan_points=1 << 16, plane_res=0.2)
    ctx.map_add_scan(scans[0], poses[0])                              # initializeMapping (LidarSlam.cpp:83-94) at the first pose
    m_np = oracle_mod.map_insert_numpy(np.zeros((0, 4), np.float32), oracle_mod.transform_scan_numpy(scans[0], poses[0]), 0.2)
    worst_dp, worst_dr, iters = 0.0, 0.0, []
    launches0 = ctx.kernel_launches()
    for i in range(1, N_SCANS):
        r = ctx.register(scans[i], priors[i], 5, 2000)                # shipped VLP-16 options (vlp_16.yaml:24-28)
        om = oracle_mod.OracleMap(m_np, ref_octree=False)
        ro = om.register(scans[i], priors[i], 0.2, 5, 2000, knn_mode=0)
        assert r.status == 0 and ro.status == 0, (i, r.status, ro.status)
        pg, po = np.array(r.pose), np.array(ro.pose)
        dp, dr = float(np.abs(pg[:3] - po[:3]).max()), float(quat_angle(pg[3:], po[3:]))
        assert dp <= 1e-4 and dr <= 1e-4, (i, dp, dr)
        assert r.n_iterations == ro.n_iterations and list(r.hist_reject_plane) == list(ro.hist_reject_plane), i
        worst_dp, worst_dr = max(worst_dp, dp), max(worst_dr, dr)
        iters.append(r.n_iterations)
        ctx.map_add_registered_scan(pg)                               # transformAndAddToMap of the scan that is still on the device
        m_np = oracle_mod.map_insert_numpy(m_np, oracle_mod.transform_scan_numpy(scans[i], pg), 0.2)
        assert ctx.map_size() == len(m_np), (i, ctx.map_size(), len(m_np))
        # and the trajectory is actually tracked (x, y: a map grown from sparse 16-beam scans observes the floor, hence z, weakly)
        assert np.linalg.norm(pg[:2] - poses[i][:2]) < 0.05, i
    got = ctx.map_download(0)
    assert got.shape == m_np.shape and np.array_equal(got, m_np[oracle_mod.cube_order(m_np)])      # bit-equal after 199 chained inserts
    print(f"[live-sequence] {N_SCANS} scans: worst |dpos| {worst_dp:.2e} m, worst |drot| {worst_dr:.2e} rad, map {len(m_np)} points, "
          f"mean ICP iterations {np.mean(iters):.2f}, {ctx.kernel_launches() - launches0} kernel launches")
    ctx.close()


def test_registered_scan_insert_equals_host_scan_insert(gpu_api):
    """so_map_add_registered_scan (device-resident scan of the last so_register) == so_map_add_scan of the same host cloud."""
    from conftest import get_case
    c = get_case("tiny")
    res = []
    for mode in (0, 1):
        ctx = gpu_api.Context(max_map_points=1 << 20, max_scan_points=65536, plane_res=0.2)
        ctx.map_add_surf(c["map_xyzi"])
        r = ctx.register(c["scan_xyzi"], c["pose_prior"], 5, 0)
        if mode == 0:
            ctx.map_add_scan(c["scan_xyzi"], np.array(r.pose))
        else:
            ctx.map_add_registered_scan(np.array(r.pose))
        res.append((np.array(r.pose), ctx.map_download(0)))
        # the index over the grown map is live: registering again works and lands on the same pose to a few millimetres
        r2 = ctx.register(c["scan_xyzi"], c["pose_prior"], 5, 0)
        assert r2.status == 0 and np.linalg.norm(np.array(r2.pose)[:3] - np.array(r.pose)[:3]) < 5e-3
        ctx.close()
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
