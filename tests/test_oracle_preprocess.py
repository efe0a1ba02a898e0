"""Oracle preprocess against an independent numpy statement of
localization.cpp:17-45,216-355."""
import numpy as np

from agile_grasp_b200 import scenes
from agile_grasp_b200.ctypes_defs import default_params


def numpy_preprocess(pts, size_left, P):
    xyz = pts[:, :3]
    fin = np.isfinite(xyz).all(1)
    x = xyz[fin]
    rank = np.arange(len(x))
    cam = (rank >= size_left).astype(np.int32) if not P.fix_cam_source else (np.nonzero(fin)[0] >= size_left).astype(np.int32)
    w = list(P.workspace)
    xd = x.astype(np.float64)
    keep = (xd[:, 0] >= w[0]) & (xd[:, 0] <= w[1]) & (xd[:, 1] >= w[2]) & (xd[:, 1] <= w[3]) & (xd[:, 2] >= w[4]) & (xd[:, 2] <= w[5])
    xd, cam = xd[keep], cam[keep]
    out_xyz, out_cam = [], []
    for c in (0, 1):
        p = xd[cam == c]
        if len(p) == 0:
            continue
        mn = np.minimum(p.min(0), 10000.0)
        k = np.floor((p - mn) / P.voxel_size).astype(np.int64)
        k = np.unique(k, axis=0)  # lexicographic sorted unique rows
        out_xyz.append((k.astype(np.float64) * P.voxel_size + mn).astype(np.float32))
        out_cam.append(np.full(len(k), c, np.int32))
    return np.concatenate(out_xyz), np.concatenate(out_cam)


def test_single_camera_matches_numpy(oracle, small_scene):
    s = small_scene
    xyz, cam = numpy_preprocess(s["pts"], s["size_left"], s["P"])
    assert xyz.shape == s["xyz"].shape
    assert (xyz.view(np.uint32) == s["xyz"].view(np.uint32)).all()
    assert (cam == s["cam"]).all() and (cam == 0).all()


def test_two_cameras_and_label_shift_quirk(oracle, two_view_scene):
    s = two_view_scene
    xyz, cam = numpy_preprocess(s["pts"], s["size_left"], s["P"])
    assert (xyz.view(np.uint32) == s["xyz"].view(np.uint32)).all() and (cam == s["cam"]).all()
    assert set(np.unique(cam)) == {0, 1}
    # camera-0 block first, each block sorted lexicographically by voxel key
    assert (np.diff(cam) >= 0).all()
    # the quirk: labels shift by the number of NaNs; fix_cam_source gives a different split
    P2 = default_params()
    for f, _ in P2._fields_:
        setattr(P2, f, getattr(s["P"], f))
    P2.fix_cam_source = 1
    xyz2, cam2 = oracle.preprocess(s["pts"], s["size_left"], P2, False)
    assert (cam2 == 0).sum() != (cam == 0).sum()
    xyz3, cam3 = numpy_preprocess(s["pts"], s["size_left"], P2)
    assert (xyz3.view(np.uint32) == xyz2.view(np.uint32)).all() and (cam3 == cam2).all()


def test_std_set_and_sort_paths_agree(oracle, small_scene):
    s = small_scene
    a = oracle.preprocess(s["pts"], s["size_left"], s["P"], True)
    assert (a[0].view(np.uint32) == s["xyz"].view(np.uint32)).all() and (a[1] == s["cam"]).all()


def test_workspace_filter_and_empty(oracle):
    pts, size_left, P, _ = scenes.config_cloud(2, small=(160, 120, 10))
    P.workspace[:] = [0.6, 0.9, -0.2, 0.2, -10, 10]
    xyz, cam = oracle.preprocess(pts, size_left, P, False)
    ref, _ = numpy_preprocess(pts, size_left, P)
    assert (xyz.view(np.uint32) == ref.view(np.uint32)).all()
    assert len(xyz) > 0 and xyz[:, 0].min() >= 0.6 - 0.003 and xyz[:, 0].max() <= 0.9
    # everything filtered out -> empty cloud, not an error
    P.workspace[:] = [100, 101, 100, 101, 100, 101]
    xyz, cam = oracle.preprocess(pts, size_left, P, False)
    assert len(xyz) == 0
    # size_left == 0 is the reference's "Input cloud is empty!" case (localization.cpp:9-15)
    try:
        oracle.preprocess(pts, 0, P, False)
        assert False
    except RuntimeError as e:
        assert "empty" in str(e)


def test_plane_removal_restatement(oracle, small_scene):
    """uses_clustering (localization.cpp:51-98): the RANSAC plane of the tabletop scene is the table; every
    iteration's inlier count and the final inlier set agree with an independent numpy statement of the same
    planes; the refitted plane is the least-squares plane of the winner's inliers."""
    xyz = small_scene["xyz"]
    res = oracle.remove_plane(xyz, seed=20150320)
    assert res is not None
    keep, counts, plane = res
    X = xyz.astype(np.float64)
    d = (plane[0] * X[:, 0] + plane[1] * X[:, 1]) + (plane[2] * X[:, 2] + plane[3])
    assert np.array_equal(keep, ~(np.abs(d) < 0.01))
    assert 0.3 < (~keep).mean() < 0.95 and counts.max() >= 0.3 * len(xyz)  # the table dominates the scene
    from agile_grasp_b200 import scenes
    n_table = scenes.table_frame()[3]
    assert abs(abs(plane[:3] @ n_table) - 1.0) < 1e-3 and abs(np.linalg.norm(plane[:3]) - 1.0) < 1e-12
    # least squares: the normal is the smallest principal direction of the inliers of the best candidate (a superset
    # check: refitting the final inliers again moves the plane by less than a tenth of a millimetre)
    P = X[~keep]
    c = P.mean(0)
    w, V = np.linalg.eigh((P - c).T @ (P - c))
    assert abs(abs(V[:, 0] @ plane[:3]) - 1.0) < 1e-4 and abs(plane[:3] @ c + plane[3]) < 1e-4
    # deterministic, and sensitive to the seed only through which triples are tried
    keep2, counts2, plane2 = oracle.remove_plane(xyz, seed=20150320)
    assert np.array_equal(keep, keep2) and np.array_equal(counts, counts2)
    keep3, counts3, _ = oracle.remove_plane(xyz, seed=7)
    assert not np.array_equal(counts, counts3) and (keep3 == keep).mean() > 0.99
    assert oracle.remove_plane(xyz[:2], seed=1) is None
