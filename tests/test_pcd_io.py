"""ag_load_pcd (the pcl::io::loadPCDFile call of localization.cpp:184,198): PCD v0.7 files written here in the
three DATA encodings by an independent writer (including a small LZF compressor that emits literal runs and
back references, overlapping ones too) must come back as the same pcl::PointXYZRGBA records."""
import struct

import numpy as np
import pytest

from agile_grasp_b200 import api


def lzf_compress(data: bytes) -> bytes:
    """greedy LZF (liblzf stream format): literal runs (ctrl < 32) and back references (len 3..264, dist <= 8192)"""
    out = bytearray()
    lit = bytearray()
    table = {}
    i, n = 0, len(data)

    def flush():
        k = 0
        while k < len(lit):
            run = lit[k:k + 32]
            out.append(len(run) - 1)
            out.extend(run)
            k += 32
        lit.clear()

    while i < n:
        key = data[i:i + 3]
        j = table.get(key, -1) if len(key) == 3 else -1
        if len(key) == 3:
            table[key] = i
        if j >= 0 and i - j <= 8192:
            length = 3
            while i + length < n and length < 264 and data[j + length] == data[i + length]:
                length += 1
            flush()
            dist = i - j - 1
            l2 = length - 2
            if l2 < 7:
                out.append((l2 << 5) | (dist >> 8))
            else:
                out.append((7 << 5) | (dist >> 8))
                out.append(l2 - 7)
            out.append(dist & 0xFF)
            i += length
        else:
            lit.append(data[i])
            i += 1
    flush()
    return bytes(out)


def make_cloud(n=1500, seed=0):
    rng = np.random.default_rng(seed)
    xyz = rng.normal(size=(n, 3)).astype(np.float32)
    xyz[::17] = np.nan  # invalid pixels of an organised cloud
    xyz[5::40, 2] = xyz[4::40, 2][: len(xyz[5::40])]  # repeated values -> back references
    rgba = rng.integers(0, 2**32, size=n, dtype=np.uint32)
    rgba[::3] = 0xFF102030
    return xyz, rgba


def header(fields, sizes, types, n, w, h, kind):
    return ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS %s\nSIZE %s\nTYPE %s\nCOUNT %s\n"
            "WIDTH %d\nHEIGHT %d\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA %s\n" % (
                " ".join(fields), " ".join(map(str, sizes)), " ".join(types), " ".join("1" for _ in fields), w, h, n, kind)).encode()


def check(path, xyz, rgba, w, h):
    pts, ww, hh = api.load_pcd(path)
    assert pts.shape == (len(xyz), 8) and (ww, hh) == (w, h)
    got = pts[:, :3]
    assert np.array_equal(np.isnan(got), np.isnan(xyz))
    assert np.array_equal(got[~np.isnan(xyz)].view(np.uint32), xyz[~np.isnan(xyz)].view(np.uint32))
    if rgba is not None:
        assert np.array_equal(pts[:, 4].view(np.uint32), rgba)


def test_binary(tmp_path):
    xyz, rgba = make_cloud()
    rec = np.zeros(len(xyz), dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("rgba", "<u4")])
    rec["x"], rec["y"], rec["z"], rec["rgba"] = xyz[:, 0], xyz[:, 1], xyz[:, 2], rgba
    p = tmp_path / "b.pcd"
    p.write_bytes(header(["x", "y", "z", "rgba"], [4, 4, 4, 4], ["F", "F", "F", "U"], len(xyz), 50, 30, "binary") + rec.tobytes())
    check(p, xyz, rgba, 50, 30)


def test_binary_field_order_and_extra_fields(tmp_path):
    """fields in another order, a float64 coordinate and an unrelated field: offsets come from the header"""
    xyz, rgba = make_cloud(400, 3)
    rec = np.zeros(len(xyz), dtype=[("rgb", "<f4"), ("z", "<f8"), ("intensity", "<f4"), ("y", "<f4"), ("x", "<f4")])
    rec["x"], rec["y"], rec["z"], rec["rgb"] = xyz[:, 0], xyz[:, 1], xyz[:, 2].astype(np.float64), rgba.view(np.float32)
    p = tmp_path / "o.pcd"
    p.write_bytes(header(["rgb", "z", "intensity", "y", "x"], [4, 8, 4, 4, 4], ["F", "F", "F", "F", "F"], len(xyz), len(xyz), 1,
                         "binary") + rec.tobytes())
    check(p, xyz, rgba, len(xyz), 1)


def test_binary_compressed(tmp_path):
    xyz, rgba = make_cloud(3000, 1)
    soa = xyz[:, 0].tobytes() + xyz[:, 1].tobytes() + xyz[:, 2].tobytes() + rgba.tobytes()
    comp = lzf_compress(soa)
    assert len(comp) < len(soa)  # the stream really contains back references
    p = tmp_path / "c.pcd"
    p.write_bytes(header(["x", "y", "z", "rgba"], [4, 4, 4, 4], ["F", "F", "F", "U"], len(xyz), 100, 30, "binary_compressed") +
                  struct.pack("<II", len(comp), len(soa)) + comp)
    check(p, xyz, rgba, 100, 30)


def test_lzf_overlapping_reference(tmp_path):
    """run-length style data: a back reference whose source overlaps its destination"""
    n = 64
    xyz = np.zeros((n, 3), np.float32)
    xyz[:, 0] = 1.5
    soa = xyz[:, 0].tobytes() + xyz[:, 1].tobytes() + xyz[:, 2].tobytes()
    comp = lzf_compress(soa)
    assert len(comp) < 40
    p = tmp_path / "r.pcd"
    p.write_bytes(header(["x", "y", "z"], [4, 4, 4], ["F", "F", "F"], n, n, 1, "binary_compressed") +
                  struct.pack("<II", len(comp), len(soa)) + comp)
    check(p, xyz, None, n, 1)


def test_ascii(tmp_path):
    xyz, rgba = make_cloud(300, 2)
    lines = []
    for (x, y, z), c in zip(xyz, rgba):
        lines.append("%s %s %s %d" % tuple(["nan" if np.isnan(v) else repr(float(v)) for v in (x, y, z)] + [int(c)]))
    p = tmp_path / "a.pcd"
    p.write_bytes(header(["x", "y", "z", "rgba"], [4, 4, 4, 4], ["F", "F", "F", "U"], len(xyz), len(xyz), 1, "ascii") +
                  ("\n".join(lines) + "\n").encode())
    check(p, xyz, rgba, len(xyz), 1)


def test_errors(tmp_path):
    with pytest.raises(RuntimeError):
        api.load_pcd(tmp_path / "missing.pcd")
    p = tmp_path / "bad.pcd"
    p.write_bytes(header(["a", "b"], [4, 4], ["F", "F"], 1, 1, 1, "binary") + b"\0" * 8)
    with pytest.raises(RuntimeError):
        api.load_pcd(p)
    p.write_bytes(header(["x", "y", "z"], [4, 4, 4], ["F", "F", "F"], 10, 10, 1, "binary") + b"\0" * 8)
    with pytest.raises(RuntimeError):
        api.load_pcd(p)
