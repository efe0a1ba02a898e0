"""Neighbour search semantics: kd-tree == brute force == numpy float32 definition."""
import numpy as np


def np_search(xyz, q, r):
    d = xyz - q  # float32
    d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
    idx = np.nonzero(d2 < np.float32(r * r))[0]
    order = np.lexsort((idx, d2[idx]))
    return idx[order].astype(np.int32), d2[idx][order]


def test_tree_equals_brute_equals_numpy(oracle, small_scene):
    s = small_scene
    rng = np.random.default_rng(0)
    for i in rng.choice(len(s["xyz"]), 25, replace=False):
        for r in (0.01, 0.03, 0.08):
            q = s["xyz"][i]
            a_i, a_d = s["tree"].radius_search(q, r, 0)
            b_i, b_d = s["tree"].radius_search(q, r, 1)
            c_i, c_d = np_search(s["xyz"], q, r)
            assert np.array_equal(a_i, b_i) and np.array_equal(a_i, c_i)
            assert np.array_equal(a_d, b_d) and np.array_equal(a_d, c_d)
            assert a_i[0] == i and a_d[0] == 0.0  # the query point itself comes first


def test_boundary_points_decided_by_float_rounding(oracle):
    # lattice vectors of squared length exactly 100 cells sit on the r=0.03 sphere: membership is
    # decided by binary32 rounding, and all three implementations must agree
    k = np.array([[10, 0, 0], [6, 8, 0], [0, 6, 8], [8, 0, 6], [0, 0, 0], [9, 4, 2], [7, 7, 1]], np.float64)
    mn = np.array([0.4123, -0.2177, 0.7311])
    xyz = (k * 0.003 + mn).astype(np.float32)
    tree = oracle.Tree(xyz)
    for r in (0.03, 0.0300001, 0.0299999):
        a_i, a_d = tree.radius_search(xyz[4], r, 0)
        b_i, b_d = tree.radius_search(xyz[4], r, 1)
        c_i, c_d = np_search(xyz, xyz[4], r)
        assert np.array_equal(a_i, c_i) and np.array_equal(b_i, c_i)


def test_sample_draw_sorted_distinct(oracle):
    for n, S in ((1000, 100), (50, 80), (12345, 2000), (7, 7)):
        idx = oracle.draw_samples(n, S, 42)
        assert len(idx) == min(n, S)
        assert (np.diff(idx) > 0).all() and idx.min() >= 0 and idx.max() < n
        assert np.array_equal(idx, oracle.draw_samples(n, S, 42))
