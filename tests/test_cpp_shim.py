"""The C++ shim (include/superodom_b200/LidarSlam.hpp) keeps the reference's LidarSLAM/LocalMap member names; a small
C++ program drives it the way laserMapping does.  CPU: it must compile and link against the in-tree library.
GPU: its poses must equal the same sequence driven through the C ABI from Python, and the oracle's."""
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import get_case, quat_angle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path, name="shim_test"):
    from superodom_b200 import build
    lib = build.build()
    exe = str(tmp_path / name)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", name + ".cpp"), "-o", exe, lib, "-Wl,-rpath," + os.path.dirname(lib)])
    return exe


def test_shim_compiles_and_links(tmp_path):
    assert os.path.exists(_build(tmp_path))
    assert os.path.exists(_build(tmp_path, "feature_test"))


def _sequence(n_scans=4):
    from superodom_b200 import synth
    c0 = get_case("tiny", 0)
    scans, priors = [], []
    for i in range(n_scans):
        c = get_case("tiny", i)
        scans.append(c["scan_xyzi"][::3].copy())
        priors.append(c["pose_true"] if i == 0 else c["pose_prior"])     # first scan: pose taken as is (map init)
    return c0, scans, priors


@pytest.mark.gpu
def test_shim_sequence_matches_abi_and_oracle(tmp_path, gpu_api, oracle_mod):
    exe = _build(tmp_path)
    c0, scans, priors = _sequence()
    path = tmp_path / "case.bin"
    with open(path, "wb") as f:
        f.write(struct.pack("<iiif", len(scans), 5, 0, 0.2))
        for s, p in zip(scans, priors):
            f.write(struct.pack("<i", len(s)))
            f.write(np.asarray(p, np.float64).tobytes())
            f.write(np.ascontiguousarray(s, np.float32).tobytes())
    out = subprocess.run([exe, str(path)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().splitlines()
    rows = [list(map(float, ln.split())) for ln in lines[:-1]]
    # same sequence through the C ABI from Python: setOrigin -> add_scan -> (register -> add_scan)*
    ctx = gpu_api.Context(max_map_points=1 << 21, max_scan_points=1 << 17, plane_res=0.2, line_res=0.1)
    ctx.map_set_origin(priors[0][:3])
    ctx.map_add_scan(scans[0], priors[0])
    om_points = oracle_mod.map_insert_numpy(np.zeros((0, 4), np.float32), oracle_mod.transform_scan_numpy(scans[0], priors[0]), 0.2,
                                            origin=tuple(ctx.map_origin()))
    assert np.array_equal(ctx.map_download(0), om_points[oracle_mod.cube_order(om_points, tuple(ctx.map_origin()))])
    assert np.allclose(rows[0][1:8], priors[0])
    prev_hist = np.zeros(9, np.int32)                   # PlaneFeatureHistogramObs before the first registration
    for i in range(1, len(scans)):
        r = ctx.register(scans[i], priors[i], 5, 0)
        assert np.array_equal(np.array(r.pose), np.array(rows[i][1:8]))              # shim == ABI, bit for bit
        # a17 EstimateLidarUncertainty (LidarSlam.cpp:915-986): stats.uncertainty_* of scan i come from scan i-1's observability histogram
        assert np.array_equal(np.array(rows[i][16:22]), oracle_mod.lidar_uncertainty(prev_hist)), (rows[i][16:22], prev_hist)
        assert rows[i][12] == rows[i][16]
        if i > 1:
            assert 0 < min(rows[i][16:22]) and max(rows[i][16:22]) <= 1.0
        prev_hist = np.array(list(r.hist_obs), np.int32)
        assert int(rows[i][8]) == r.n_iterations and int(rows[i][9]) == r.map_surf_5x5
        # oracle on the same (device-built) map
        om = oracle_mod.OracleMap()
        om.set_origin(ctx.map_origin())             # the origin shiftMap left behind inside so_register
        om.set_points(om_points)
        ro = om.register(scans[i], priors[i], 0.2, 5, 0, knn_mode=0, skip_map_checks=True)
        assert np.abs(np.array(r.pose)[:3] - np.array(ro.pose)[:3]).max() < 1e-4 and quat_angle(np.array(r.pose)[3:], np.array(ro.pose)[3:]) < 1e-4
        ctx.map_add_scan(scans[i], np.array(r.pose))
        om_points = oracle_mod.map_insert_numpy(om_points, oracle_mod.transform_scan_numpy(scans[i], np.array(r.pose)), 0.2,
                                                origin=tuple(ctx.map_origin()))
        assert np.array_equal(ctx.map_download(0), om_points[oracle_mod.cube_order(om_points, tuple(ctx.map_origin()))])   # voxel-filter insert: bit-exact
        assert int(rows[i][10]) == len(om_points)
        # the map here is whatever the previous sparse VLP-16 scans inserted (weak floor coverage): only x, y are well observed
        assert np.linalg.norm(np.array(r.pose)[:2] - get_case("tiny", i)["pose_true"][:2]) < 0.03
    n_all, n_near = map(int, lines[-1].split()[1:])
    assert n_all == len(om_points) and 0 < n_near <= n_all
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("imu_only", [0, 1])
def test_feature_extraction_shim_matches_abi(tmp_path, gpu_api, imu_only):
    """FeatureExtraction.hpp (removePointDistortion + uniformFeatureExtraction on PointcloudXYZITR records and a std::map pose
    buffer) == the same calls through the C ABI from Python, bit for bit."""
    from superodom_b200 import synth
    exe = _build(tmp_path, "feature_test")
    d = synth.make_raw_sweep(30_000, seed=4200)
    pts = d["points"].copy()
    pts[~np.isfinite(pts)] = 0.0                 # the record reader of the test program casts the ring column to uint16
    skip, block_range = 3, 0.2
    path, outp = tmp_path / "raw.bin", tmp_path / "out.bin"
    with open(path, "wb") as f:
        f.write(struct.pack("<iiiifd", len(pts), len(d["sample_times"]), imu_only, skip, block_range, d["start_time"]))
        f.write(np.asarray(d["T_i_l"], np.float64).tobytes())
        f.write(pts.tobytes())
        for t, p in zip(d["sample_times"], d["sample_poses"]):
            f.write(struct.pack("<d", t) + np.asarray(p, np.float64).tobytes())
    out = subprocess.run([exe, str(path), str(outp)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    raw = open(outp, "rb").read()
    start = np.frombuffer(raw[:56], np.float64)
    xyz = np.frombuffer(raw[56:56 + 12 * len(pts)], np.float32).reshape(-1, 3)
    k = struct.unpack("<i", raw[56 + 12 * len(pts):60 + 12 * len(pts)])[0]
    feat = np.frombuffer(raw[60 + 12 * len(pts):], np.float32).reshape(-1, 4)
    assert len(feat) == k
    ctx = gpu_api.Context(max_map_points=1024, max_scan_points=1 << 18, plane_res=0.2)
    got = pts.copy()
    st, past = ctx.scan_deskew(got, 5, d["start_time"], d["sample_times"], d["sample_poses"], imu_only=bool(imu_only), T_i_l=d["T_i_l"])
    assert past == 0 and np.array_equal(start, st)
    assert np.array_equal(xyz, got[:, :3]) and np.abs(xyz - pts[:, :3]).max() > 0.05
    exp_feat = ctx.scan_extract_uniform(got, 5, skip, block_range)
    assert np.array_equal(feat, exp_feat) and 0 < k < len(pts)
    ctx.close()
