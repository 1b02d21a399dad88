"""include/superodom_b200/PcdIO.hpp: the prior-map reader of localization mode (utils::readPointCloud -> pcl::PCDReader upstream,
superodom_utils.cpp:16-33).  PCD files in the three DATA encodings are written here from the published container layout and must
come back bit-exact; malformed files return false, as upstream."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    d = tmp_path_factory.mktemp("pcd")
    out = str(d / "pcd_test")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "pcd_test.cpp"), "-o", out])
    return out


def lzf_compress(data: bytes) -> bytes:
    """Small greedy LZF encoder (literal runs + back-references), enough to exercise both paths of the decoder."""
    out, lit, i, n, table = bytearray(), bytearray(), 0, len(data), {}

    def flush():
        nonlocal lit
        while lit:
            run = lit[:32]
            out.append(len(run) - 1)
            out.extend(run)
            lit = lit[32:]
    while i < n:
        key = data[i:i + 3]
        j = table.get(key, -1) if len(key) == 3 else -1
        if len(key) == 3:
            table[key] = i
        if j >= 0 and 0 < i - j <= 8192:
            ln = 3
            while i + ln < n and ln < 264 and data[j + ln] == data[i + ln]:
                ln += 1
            flush()
            dist, l2 = i - j - 1, ln - 2
            if l2 < 7:
                out.append((l2 << 5) | (dist >> 8))
            else:
                out.append((7 << 5) | (dist >> 8))
                out.append(l2 - 7)
            out.append(dist & 0xFF)
            i += ln
        else:
            lit.append(data[i])
            i += 1
    flush()
    return bytes(out)


def _cloud(n=5000, seed=7):
    rng = np.random.default_rng(seed)
    xyz = rng.normal(0, 20, (n, 3)).astype(np.float32)
    xyz[::97] = np.round(xyz[::97])                      # repeated byte patterns => LZF back-references
    inten = rng.uniform(0, 255, n).astype(np.float32)
    ring = rng.integers(0, 128, n).astype(np.uint16)
    t = np.linspace(0, 0.1, n)
    xyz[5, 0] = np.nan
    return xyz, inten, ring, t


HEADER = "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity ring time\nSIZE 4 4 4 4 2 8\nTYPE F F F F U F\n" \
         "COUNT 1 1 1 1 1 1\nWIDTH {n}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {n}\nDATA {mode}\n"


def _read(exe, path, tmp_path):
    out = tmp_path / "out.bin"
    r = subprocess.run([exe, str(path), str(out)], capture_output=True, text=True)
    return r.returncode, (np.fromfile(out, np.float32).reshape(-1, 4) if r.returncode == 0 else None), r.stderr


def test_binary_ascii_and_compressed_round_trip(exe, tmp_path):
    xyz, inten, ring, t = _cloud()
    n = len(xyz)
    exp = np.concatenate([xyz, inten[:, None]], 1)
    rec = np.zeros(n, dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("i", "<f4"), ("r", "<u2"), ("t", "<f8")])
    rec["x"], rec["y"], rec["z"], rec["i"], rec["r"], rec["t"] = xyz[:, 0], xyz[:, 1], xyz[:, 2], inten, ring, t
    p = tmp_path / "b.pcd"
    p.write_bytes(HEADER.format(n=n, mode="binary").encode() + rec.tobytes())
    rc, got, _ = _read(exe, p, tmp_path)
    assert rc == 0 and np.array_equal(got.view(np.uint32), exp.view(np.uint32))
    # binary_compressed: field-by-field layout, LZF
    soa = b"".join(rec[k].tobytes() for k in ("x", "y", "z", "i", "r", "t"))
    comp = lzf_compress(soa)
    assert len(comp) < len(soa)                                        # the fixture really contains back-references
    p = tmp_path / "c.pcd"
    p.write_bytes(HEADER.format(n=n, mode="binary_compressed").encode() + struct.pack("<II", len(comp), len(soa)) + comp)
    rc, got, err = _read(exe, p, tmp_path)
    assert rc == 0, err
    assert np.array_equal(got.view(np.uint32), exp.view(np.uint32))
    # ascii (repr round-trips float32 exactly through 9 significant digits), CRLF line ends tolerated
    p = tmp_path / "a.pcd"
    rows = "\r\n".join(f"{x:.9g} {y:.9g} {z:.9g} {i:.9g} {r} {tt:.17g}" for (x, y, z), i, r, tt in zip(xyz, inten, ring, t))
    p.write_bytes((HEADER.format(n=n, mode="ascii").replace("\n", "\r\n") + rows + "\r\n").encode())
    rc, got, err = _read(exe, p, tmp_path)
    assert rc == 0, err
    assert np.array_equal(got.view(np.uint32)[np.isfinite(exp).all(1)], exp.view(np.uint32)[np.isfinite(exp).all(1)]) and np.isnan(got[5, 0])


def test_missing_intensity_other_types_and_errors(exe, tmp_path):
    n = 100
    rng = np.random.default_rng(3)
    xyz = rng.normal(0, 5, (n, 3))
    # x y z as doubles, no intensity field, an ignored multi-count field in front
    hdr = f"VERSION .7\nFIELDS normal x y z\nSIZE 4 8 8 8\nTYPE F F F F\nCOUNT 3 1 1 1\nWIDTH {n // 4}\nHEIGHT 4\nDATA binary\n"
    rec = np.zeros(n, dtype=[("nrm", "<f4", 3), ("x", "<f8"), ("y", "<f8"), ("z", "<f8")])
    rec["x"], rec["y"], rec["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    p = tmp_path / "d.pcd"
    p.write_bytes(hdr.encode() + rec.tobytes())
    rc, got, err = _read(exe, p, tmp_path)
    assert rc == 0, err
    assert np.array_equal(got[:, :3], xyz.astype(np.float32)) and (got[:, 3] == 0).all()      # organised 25 x 4 -> 100 points
    # errors: missing file, truncated payload, no xyz
    assert _read(exe, tmp_path / "nope.pcd", tmp_path)[0] == 1
    p.write_bytes(hdr.encode() + rec.tobytes()[:-7])
    assert _read(exe, p, tmp_path)[0] == 1
    p.write_bytes(b"VERSION .7\nFIELDS a b\nSIZE 4 4\nTYPE F F\nCOUNT 1 1\nWIDTH 1\nHEIGHT 1\nDATA ascii\n1 2\n")
    assert _read(exe, p, tmp_path)[0] == 1


def test_hostile_headers_are_rejected_without_reading_out_of_bounds(exe, tmp_path):
    """Untrusted files: COUNT 0 fields, odd SIZEs, record / point counts that would wrap or exhaust memory, and a compressed
    size larger than the file all return false (exit code 1) instead of reading past a buffer or throwing."""
    p = tmp_path / "h.pcd"
    body = np.zeros(8, np.float32).tobytes()
    cases = [
        "VERSION .7\nFIELDS x y z i\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 0\nWIDTH 2\nHEIGHT 1\nDATA binary\n",          # last field COUNT 0
        "VERSION .7\nFIELDS x y z\nSIZE 4 4 3\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 2\nHEIGHT 1\nDATA binary\n",                  # SIZE 3
        "VERSION .7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1000000\nWIDTH 2\nHEIGHT 1\nDATA binary\n",            # huge COUNT
        "VERSION .7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 2000000000\nHEIGHT 1\nDATA binary\n",          # points * rec >> file
        "VERSION .7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 2000000000\nHEIGHT 1\nDATA ascii\n",           # ascii rows >> file
    ]
    for hdr in cases:
        p.write_bytes(hdr.encode() + body)
        assert _read(exe, p, tmp_path)[0] == 1, hdr
    hdr = "VERSION .7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 2\nHEIGHT 1\nDATA binary_compressed\n"
    p.write_bytes(hdr.encode() + struct.pack("<II", 0xFFFFFFF0, 24) + body)                                                 # csize >> file
    assert _read(exe, p, tmp_path)[0] == 1
