"""Seeded synthetic tabletop scenes rendered as organised PointXYZRGBA clouds.

The reference ships no .pcd file, so every BASELINE.json configuration runs on
clouds from this generator: a pinhole depth camera (fx=fy=525, 640x480, the Kinect model the
reference's launch files assume) ray-casts a slightly tilted table carrying cylinders, boxes,
spheres and one bar-shaped "handle"; depth noise is N(0, (1 mm * z^2)^2); a few percent of the
pixels are dropped to NaN like a real structured-light sensor.  Clouds are expressed in the robot
base frame with the camera at the pose the reference executables use (src/nodes/test.cpp:47-74),
so the camera origins in ag_params are meaningful.  The table is tilted ~8 degrees against the
base axes so that voxelised neighbourhoods are never exactly planar / axis aligned (an exactly
planar neighbourhood makes the reference's Taubin eigenproblem degenerate and its output
arbitrary — SURVEY.md §7 hard part 2).
"""
import numpy as np

from .ctypes_defs import BASE_TF, SQRT_TF, default_params

FX = FY = 525.0


def _unit(v):
    v = np.asarray(v, dtype=np.float64)
    return v / np.linalg.norm(v)


def table_frame():
    n = _unit([0.12, -0.08, 1.0])
    u = _unit(np.cross([0.0, 1.0, 0.0], n))
    v = np.cross(n, u)
    origin = np.array([0.95, 0.0, -0.22])
    return origin, u, v, n


def make_scene(seed, n_objects=None):
    """Returns a list of primitives placed on the table (all in the base frame)."""
    rng = np.random.default_rng(seed)
    origin, u, v, n = table_frame()
    prims = [dict(kind="rect", p0=origin, u=u, v=v, n=n, hu=0.55, hv=0.65)]
    k = int(rng.integers(8, 16)) if n_objects is None else n_objects
    placed = []
    tries = 0
    while len(placed) < k and tries < 2000:
        tries += 1
        a = rng.uniform(-0.38, 0.38)
        b = rng.uniform(-0.5, 0.5)
        rad = rng.uniform(0.05, 0.09)
        if any((a - pa) ** 2 + (b - pb) ** 2 < (rad + pr) ** 2 for pa, pb, pr in placed):
            continue
        placed.append((a, b, rad))
    kinds = ["cyl_up", "cyl_lying", "box", "sphere"]
    for i, (a, b, rad) in enumerate(placed):
        base = origin + a * u + b * v
        kind = "bar" if i == 0 else kinds[int(rng.integers(0, len(kinds)))]
        yaw = rng.uniform(0, np.pi)
        d1 = np.cos(yaw) * u + np.sin(yaw) * v
        d2 = np.cross(n, d1)
        if kind == "cyl_up":
            r = rng.uniform(0.02, 0.045)
            h = rng.uniform(0.06, 0.16)
            prims.append(dict(kind="cyl", c=base + n * (h / 2), a=n, hl=h / 2, r=r))
        elif kind == "cyl_lying":
            r = rng.uniform(0.02, 0.04)
            hl = rng.uniform(0.04, min(0.09, rad + 0.02))
            prims.append(dict(kind="cyl", c=base + n * r, a=d1, hl=hl, r=r))
        elif kind == "box":
            hx = rng.uniform(0.015, 0.05)
            hy = rng.uniform(0.015, 0.05)
            hz = rng.uniform(0.02, 0.06)
            prims.append(dict(kind="box", c=base + n * hz, R=np.stack([d1, d2, n], 1), h=np.array([hx, hy, hz])))
        elif kind == "sphere":
            r = rng.uniform(0.03, 0.055)
            prims.append(dict(kind="sphere", c=base + n * r, r=r))
        else:  # the 15 cm bar "handle", raised on two posts
            prims.append(dict(kind="cyl", c=base + n * 0.07, a=d1, hl=0.075, r=0.012))
            for sgn in (-1, 1):
                prims.append(dict(kind="cyl", c=base + sgn * 0.07 * d1 + n * 0.035, a=n, hl=0.035, r=0.01))
    return prims


def _hit(prim, o, d):
    """ray parameter t (depth along unit-z camera rays) per ray; inf where missed."""
    inf = np.full(d.shape[0], np.inf)
    kind = prim["kind"]
    if kind == "rect":
        den = d @ prim["n"]
        with np.errstate(divide="ignore", invalid="ignore"):
            t = ((prim["p0"] - o) @ prim["n"]) / den
        p = o + t[:, None] * d - prim["p0"]
        ok = (t > 0) & (np.abs(p @ prim["u"]) <= prim["hu"]) & (np.abs(p @ prim["v"]) <= prim["hv"])
        return np.where(ok, t, inf)
    if kind == "sphere":
        oc = o - prim["c"]
        a = np.einsum("ij,ij->i", d, d)
        b = 2 * (d @ oc)
        c = oc @ oc - prim["r"] ** 2
        disc = b * b - 4 * a * c
        with np.errstate(invalid="ignore"):
            t = (-b - np.sqrt(disc)) / (2 * a)
        return np.where((disc > 0) & (t > 0), t, inf)
    if kind == "cyl":
        ax = prim["a"]
        oc = o - prim["c"]
        dpar = d @ ax
        dperp = d - dpar[:, None] * ax
        ocpar = oc @ ax
        ocperp = oc - ocpar * ax
        a = np.einsum("ij,ij->i", dperp, dperp)
        b = 2 * (dperp @ ocperp)
        c = ocperp @ ocperp - prim["r"] ** 2
        disc = b * b - 4 * a * c
        with np.errstate(invalid="ignore", divide="ignore"):
            t = (-b - np.sqrt(disc)) / (2 * a)
            s = ocpar + t * dpar
            side = np.where((disc > 0) & (t > 0) & (np.abs(s) <= prim["hl"]), t, inf)
            best = side
            for sgn in (-1.0, 1.0):  # caps
                tc = (sgn * prim["hl"] - ocpar) / dpar
                pc = ocperp + tc[:, None] * dperp
                okc = (tc > 0) & (np.einsum("ij,ij->i", pc, pc) <= prim["r"] ** 2)
                best = np.minimum(best, np.where(okc, tc, inf))
        return best
    if kind == "box":
        R = prim["R"]
        ol = (o - prim["c"]) @ R
        dl = d @ R
        with np.errstate(divide="ignore", invalid="ignore"):
            t1 = (-prim["h"] - ol) / dl
            t2 = (prim["h"] - ol) / dl
        tmin = np.nanmax(np.minimum(t1, t2), axis=1)
        tmax = np.nanmin(np.maximum(t1, t2), axis=1)
        return np.where((tmax >= tmin) & (tmin > 0), tmin, inf)
    raise ValueError(kind)


def render(prims, cam_tf, seed, width=640, height=480, dropout=0.04, noise=0.001):
    """Organised cloud (height*width, 8) float32 = PointXYZRGBA records, base frame, NaN = invalid."""
    rng = np.random.default_rng(seed)
    cam_tf = np.asarray(cam_tf, dtype=np.float64)
    cx, cy = (width - 1) / 2.0, (height - 1) / 2.0
    fx = FX * width / 640.0
    fy = FY * height / 480.0
    uu, vv = np.meshgrid(np.arange(width), np.arange(height))
    dc = np.stack([(uu.ravel() - cx) / fx, (vv.ravel() - cy) / fy, np.ones(width * height)], 1)
    R, o = cam_tf[:3, :3], cam_tf[:3, 3]
    d = dc @ R.T
    depth = np.full(width * height, np.inf)
    for prim in prims:
        depth = np.minimum(depth, _hit(prim, o, d))
    valid = np.isfinite(depth) & (depth < 4.0)
    depth = np.where(valid, depth, 1.0)
    depth = depth + noise * depth * depth * rng.standard_normal(depth.shape)
    valid &= rng.random(depth.shape) >= dropout
    pts = o + depth[:, None] * d
    out = np.zeros((width * height, 8), np.float32)
    out[:, :3] = np.where(valid[:, None], pts, np.nan).astype(np.float32)
    out[:, 3] = 1.0
    rgba = rng.integers(0, 2 ** 32, size=width * height, dtype=np.uint64).astype(np.uint32)
    out[:, 4] = rgba.view(np.float32)
    return out


CONFIGS = {
    1: dict(name="tabletop_50k_400", width=320, height=240, samples=400, views=1),
    2: dict(name="vga_307k_2000", width=640, height=480, samples=2000, views=1),
    3: dict(name="two_view_614k_4000", width=640, height=480, samples=4000, views=2),
    4: dict(name="batch32_vga_2000", width=640, height=480, samples=2000, views=1, batch=32),
    5: dict(name="fused_2m_20000", width=640, height=480, samples=20000, views=7),
}


def view_poses(nviews):
    """camera poses: single camera = launch-file camera_pose (launch/single_camera_grasps.launch:11-14);
    two views = base_tf*sqrt_tf^-1 / base_tf*sqrt_tf (find_grasps.cpp:36-45); more views fan around."""
    if nviews == 1:
        return [BASE_TF.copy()]
    if nviews == 2:
        return [BASE_TF @ np.linalg.inv(SQRT_TF), BASE_TF @ SQRT_TF]
    poses = []
    for k in range(nviews):
        ang = (k - (nviews - 1) / 2.0) * 0.22
        c, s = np.cos(ang), np.sin(ang)
        Rz = np.array([[c, -s, 0, 0], [s, c, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]])
        T = np.eye(4)
        T[:3, 3] = [0.95, 0.0, 0.0]
        Ti = np.eye(4)
        Ti[:3, 3] = [-0.95, 0.0, 0.0]
        poses.append(T @ Rz @ Ti @ BASE_TF)
    return poses


def config_cloud(config_id, scene_offset=0, small=None):
    """Returns (points32 (n,8) f32, size_left, AgParams, num_samples) for a BASELINE.json config.
    `small` = (width, height, samples) override for quick tests."""
    cfg = dict(CONFIGS[config_id])
    if small is not None:
        cfg["width"], cfg["height"], cfg["samples"] = small
    seed = 20150320 + config_id + 1000 * scene_offset
    prims = make_scene(seed)
    poses = view_poses(cfg["views"])
    clouds = [render(prims, T, seed * 7 + i, cfg["width"], cfg["height"]) for i, T in enumerate(poses)]
    pts = np.concatenate(clouds, 0)
    if cfg["views"] == 1:
        size_left = pts.shape[0]
        p = default_params(cam_tf_left=poses[0], cam_tf_right=poses[0])
    elif cfg["views"] == 2:
        size_left = clouds[0].shape[0]
        p = default_params(cam_tf_left=poses[0], cam_tf_right=poses[1])
    else:  # fused scene: treated as one registered cloud seen from the central pose
        size_left = pts.shape[0]
        mid = poses[len(poses) // 2]
        p = default_params(cam_tf_left=mid, cam_tf_right=mid)
    p.num_samples = cfg["samples"]
    p.workspace[:] = [-10, 10, -10, 10, -10, 10] if config_id != 1 else [-10, 10, -10, 10, -10, 1]
    p.filters_boundaries = 1 if config_id == 1 else 0  # test_svm builds Localization(.., true, ..) (test.cpp:72)
    return pts, size_left, p, cfg["samples"]
