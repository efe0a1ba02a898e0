"""HandleSearch::findHandles + Handle (handle_search.cpp:4-118, handle.cpp:3-73): the oracle against an
independent numpy statement of the reference's loops, including the behaviour of shortenHandle as it
actually executes (the out-of-range read `inliers[i](2)` makes the lower-part branch the only one taken)."""
import numpy as np
import pytest

from agile_grasp_b200.ctypes_defs import GRASP_DTYPE


def synthetic_grasps(seed=0, clutter=40):
    """two bars (collinear grasps every 5 mm, the second one with a 4 cm hole) + random clutter"""
    rng = np.random.default_rng(seed)
    rows = []

    def bar(origin, u, approach, ts, jitter):
        u = u / np.linalg.norm(u)
        approach = approach - u * (approach @ u)
        approach /= np.linalg.norm(approach)
        for t in ts:
            a = u + rng.normal(scale=0.02, size=3)
            a /= np.linalg.norm(a)
            ap = approach + rng.normal(scale=0.03, size=3)
            ap /= np.linalg.norm(ap)
            p = origin + t * u + rng.normal(scale=jitter, size=3)
            rows.append((a * rng.choice([-1, 1]), ap, p, p - 0.02 * ap, 0.03 + 0.01 * rng.random()))

    bar(np.array([0.1, 0.0, 0.5]), np.array([1.0, 0.2, 0.0]), np.array([0.0, 0.0, -1.0]), np.arange(0, 0.12, 0.005), 0.001)
    ts = np.concatenate([np.arange(0, 0.05, 0.005), np.arange(0.09, 0.16, 0.005)])
    bar(np.array([-0.2, 0.1, 0.6]), np.array([0.0, 1.0, 0.3]), np.array([1.0, 0.0, 0.0]), ts, 0.0015)
    for _ in range(clutter):
        a = rng.normal(size=3); a /= np.linalg.norm(a)
        ap = rng.normal(size=3); ap /= np.linalg.norm(ap)
        p = rng.uniform(-0.3, 0.3, size=3) + np.array([0, 0, 0.5])
        rows.append((a, ap, p, p - 0.02 * ap, 0.04))
    order = rng.permutation(len(rows))
    g = np.zeros(len(rows), GRASP_DTYPE)
    for k, r in enumerate(order):
        a, ap, p, s, w = rows[r]
        g[k]["axis"], g[k]["approach"], g[k]["bottom"], g[k]["surface"], g[k]["width"] = a, ap, p, s, w
        g[k]["binormal"] = np.cross(ap, a)
    return g


def numpy_find_handles(g, min_inliers, min_length):
    """literal restatement of handle_search.cpp:4-118 / handle.cpp:3-73 in numpy (vectorised inner loop)"""
    n = len(g)
    width = g["width"].copy()
    handles, inliers = [], []
    for i in range(n):
        if width[i] == -1:
            continue
        live = np.nonzero(width != -1)[0]
        ia, ip, inrm = g["axis"][i], g["bottom"][i], g["approach"][i]
        d = g["bottom"][live] - ip
        proj = d - np.outer(d @ ia, ia)
        dist_from_line = np.linalg.norm(proj, axis=1)
        along = d @ ia
        aa = np.arccos(np.clip(g["axis"][live] @ ia, -1, 1))
        ang = np.minimum(aa, np.pi - aa)
        an = np.arccos(np.clip(g["approach"][live] @ inrm, -1, 1))
        ok = (dist_from_line < 0.01) & (ang < 0.34) & (an < 0.34)
        js, ds = live[ok], along[ok]
        if len(js) < min_inliers:
            continue
        o = np.lexsort((js, ds))
        js, ds = js[o], ds[o]
        gaps = np.nonzero(np.diff(ds) > 0.02)[0]
        if len(gaps):
            js, ds = js[:gaps[0]], ds[:gaps[0]]  # elements strictly before position i (the quirk)
        if len(js) < min_inliers:
            continue
        if not (ds.max() - ds.min() > min_length):
            continue
        A = g["axis"][js]
        w, V = np.linalg.eigh(A.T @ A)
        axis = V[:, np.argmax(w)]
        if axis @ g["axis"][js[0]] < 0:
            axis = -axis
        al = g["bottom"][js] @ axis
        k = int(np.argmin(np.abs(al - (al.max() + al.min()) / 2.0)))
        handles.append(dict(axis=axis, center=g["bottom"][js[k]], approach=g["approach"][js[k]],
                            hands_center=g["surface"][js[k]], binormal=np.cross(g["approach"][js[k]], axis),
                            width=g["width"][js].mean()))
        inliers.append(js)
        width[js] = -1
    return handles, inliers


@pytest.mark.parametrize("seed,min_inliers,min_length", [(0, 3, 0.005), (1, 5, 0.02), (2, 2, 0.0)])
def test_oracle_matches_numpy_statement(oracle, seed, min_inliers, min_length):
    g = synthetic_grasps(seed)
    H, inl = oracle.find_handles(g, min_inliers, min_length)
    Hn, inln = numpy_find_handles(g, min_inliers, min_length)
    assert len(H) == len(Hn) and len(H) >= 2
    for k in range(len(H)):
        assert np.array_equal(inl[k], inln[k])
        for nm in ("center", "approach", "hands_center"):
            assert np.array_equal(H[k][nm], Hn[k][nm]), nm
        assert np.allclose(H[k]["axis"], Hn[k]["axis"], atol=1e-12)
        assert np.allclose(H[k]["binormal"], Hn[k]["binormal"], atol=1e-12)
        assert abs(H[k]["width"] - Hn[k]["width"]) <= 1e-15


def test_shorten_handle_keeps_only_the_part_before_the_first_gap(oracle):
    """handle_search.cpp:103 as executed: a 4 cm hole in a bar cuts the inlier list to the grasps strictly
    before the gap position (the grasp at the gap's lower edge is dropped too)."""
    g = synthetic_grasps(3, clutter=0)
    H, inl = oracle.find_handles(g, 3, 0.005)
    for k in range(len(H)):
        al = (g["bottom"][inl[k]] - g["bottom"][inl[k][0]]) @ H[k]["axis"]
        assert np.all(np.diff(np.sort(al)) <= 0.02 + 1e-9)
    # eliminated grasps never appear twice
    flat = np.concatenate(inl)
    assert len(flat) == len(set(flat.tolist()))


def test_width_sentinel_hides_inputs(oracle):
    g = synthetic_grasps(4)
    g["width"][::2] = -1  # the reference's own "eliminated" marker (handle_search.cpp:13,23)
    H, inl = oracle.find_handles(g, 2, 0.0)
    for k in range(len(H)):
        assert np.all(inl[k] % 2 == 1)
