import numpy as np

from conftest import get_case
from superodom_b200 import synth


def test_generator_is_deterministic():
    a = synth.make_case("tiny")
    b = synth.make_case("tiny")
    for k in ("map_xyzi", "scan_xyzi", "pose_true", "pose_prior"):
        assert np.array_equal(a[k], b[k])


def test_no_plane_through_origin_and_block_layout():
    c = get_case("cfg1")
    b = c["scene"].boxes
    for ax in range(3):
        assert (np.abs(b[:, 2 * ax:2 * ax + 2]) >= 1.0 - 1e-9).all()
    lin = synth.block_linear(synth.block_of(c["map_xyzi"][:, :3]))
    assert (lin == 10 + 21 * 10 + 21 * 21 * 5).all()          # one 50 m block, the constructor-origin centre block
    assert 90_000 < len(c["map_xyzi"]) < 110_000
    assert len(c["scan_xyzi"]) == 28_800


def test_voxel_filter_semantics():
    rng = np.random.default_rng(3)
    pts = rng.uniform(-3, 3, size=(5000, 4)).astype(np.float32)
    out = synth.voxel_filter_blocks(pts, 0.5)
    inv = np.float32(1.0) / np.float32(0.5)
    vox = np.floor(pts[:, :3] * inv).astype(np.int64)
    # one output per occupied voxel, each output inside its voxel, ordered by (k, j, i)
    uniq = np.unique(vox, axis=0)
    assert len(out) == len(uniq)
    ov = np.floor(out[:, :3] * inv).astype(np.int64)
    assert len(np.unique(ov, axis=0)) == len(out)
    key = (ov[:, 2] * 1000 + ov[:, 1]) * 1000 + ov[:, 0]
    assert (np.diff(key) > 0).all()
    # centroid of a known voxel, float32 sequential accumulation
    v0 = ov[0]
    m = (vox == v0).all(1)
    acc = np.zeros(4, np.float32)
    for p in pts[m]:
        acc = acc + p
    assert np.array_equal(out[0], acc / np.float32(m.sum()))
    # idempotent on its own output (one point per voxel stays put)
    assert np.array_equal(synth.voxel_filter_blocks(out, 0.5), out)


def test_block_of_matches_reference_truncation_quirk():
    # int((x+25)/50) then -1 if negative: x+25 == -50 exactly lands in block -2 (not -1)
    x = np.array([[-75.0, 0, 0], [-74.999, 0, 0], [24.999, 0, 0], [25.0, 0, 0], [-25.001, 0, 0]], np.float32)
    c = synth.block_of(x, origin=(0, 0, 0))
    assert list(c[:, 0]) == [-2, -1, 0, 1, -1]


def test_scan_ranges_and_sampling_rule():
    c = get_case("cfg1")
    r = np.linalg.norm(c["scan_xyzi"][:, :3], axis=1)
    assert r.min() > 0.2 and r.max() < 130.0
