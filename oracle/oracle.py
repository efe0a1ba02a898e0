"""ctypes binding of the CPU oracle (oracle/libag_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Never imported by the agile_grasp_b200 package.
"""
import ctypes as C
import glob
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from agile_grasp_b200.ctypes_defs import (AG_HOG_DIM, AG_IMAGE_COLS, AG_IMAGE_ROWS, FRAME_DTYPE,  # noqa: E402
                                          GRASP_DTYPE, AgFrame, AgGrasp, AgParams)

_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libag_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle_core.cpp", "oracle_learn.cpp", "oracle_internal.h",
                                             "ag_oracle.h", "Makefile")]
    srcs.append(os.path.join(_HERE, "..", "include", "ag_b200.h"))
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs
                                            if os.path.exists(s))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"], env={k: v for k, v in os.environ.items()
                                                               if k not in ("CXX", "CC")})
    return so


def _find_lapack():
    """dggev_ providers present in this image: cv2's OpenBLAS (dggev_), scipy's (scipy_dggev_)."""
    env = os.environ.get("AG_ORACLE_LAPACK")
    if env:
        path, _, sym = env.partition(":")
        return [(path, sym or "dggev_")]
    first, second = [], []
    for sp in sys.path:
        for f in glob.glob(os.path.join(sp, "opencv_python_headless.libs", "libopenblas*.so*")):
            first.append((f, "dggev_"))
        for f in glob.glob(os.path.join(sp, "scipy.libs", "libscipy_openblas*.so*")):
            second.append((f, "scipy_dggev_"))
    return first + second


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    so = build()
    L = C.CDLL(so)
    L.ago_last_error.restype = C.c_char_p
    L.ago_tree_build.restype = C.c_void_p
    L.ago_find_hands.restype = C.c_void_p
    L.ago_hands_grasps.restype = C.POINTER(AgGrasp)
    L.ago_svm_load.restype = C.c_void_p
    L.ago_svm_sv.restype = C.POINTER(C.c_float)
    L.ago_svm_alpha.restype = C.POINTER(C.c_double)
    L.ago_svm_decision.restype = C.c_float
    L.ago_localize.restype = C.c_void_p
    for path, sym in _find_lapack():
        if set_lapack(path, sym, L):
            break
    _LIB = L
    return L


def set_lapack(path, sym, L=None):
    """Select the dggev_ provider.  The wheel-bundled OpenBLAS needs its sibling libgfortran /
    libquadmath, which are not on the loader path: preload them globally first."""
    L = L or lib()
    d = os.path.dirname(path)
    for pat in ("libquadmath*.so*", "libgfortran*.so*"):
        for dep in sorted(glob.glob(os.path.join(d, pat))):
            try:
                C.CDLL(dep, mode=C.RTLD_GLOBAL)
            except OSError:
                pass
    if L.ago_set_lapack(path.encode(), sym.encode()) == 0:
        L._lapack = (path, sym)
        return True
    return False


def lapack_providers():
    return _find_lapack()


def _err():
    return lib().ago_last_error().decode()


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _take(ptr, n, dtype):
    """copy n items out of a malloc'ed buffer and free it"""
    if n == 0:
        out = np.zeros(0, dtype)
    else:
        out = np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)
    lib().ago_free(ptr)
    return out


def preprocess(points32, size_left, params: AgParams, use_std_set=False):
    """points32: (n, 8) float32 view of PointXYZRGBA records (32-byte stride)."""
    pts = np.ascontiguousarray(points32)
    stride = pts.strides[0]
    xyz = C.POINTER(C.c_float)()
    cam = C.POINTER(C.c_int32)()
    n = C.c_int()
    rc = lib().ago_preprocess(pts.ctypes.data_as(C.c_void_p), stride, pts.shape[0], int(size_left), C.byref(params),
                              int(use_std_set), C.byref(xyz), C.byref(cam), C.byref(n))
    if rc != 0:
        raise RuntimeError(_err())
    return _take(xyz, n.value * 3, np.float32).reshape(-1, 3), _take(cam, n.value, np.int32)


class Tree:
    def __init__(self, xyz):
        self.xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        self.h = C.c_void_p(lib().ago_tree_build(_p(self.xyz, C.c_float), self.xyz.shape[0]))

    def __del__(self):
        if getattr(self, "h", None):
            lib().ago_tree_free(self.h)
            self.h = None

    def radius_search(self, q, radius, method=1):
        q = np.ascontiguousarray(q, dtype=np.float32)
        idx = C.POINTER(C.c_int32)()
        dist = C.POINTER(C.c_float)()
        n = C.c_int()
        lib().ago_radius_search(self.h, _p(self.xyz, C.c_float), self.xyz.shape[0], _p(q, C.c_float),
                                C.c_double(radius), method, C.byref(idx), C.byref(dist), C.byref(n))
        return _take(idx, n.value, np.int32), _take(dist, n.value, np.float32)


def fit_quadrics(tree: Tree, cam, indices, radius, params: AgParams, sum_perm=0, want_params=False, want_mn=False):
    cam = np.ascontiguousarray(cam, dtype=np.int32)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    S = indices.shape[0]
    frames = np.zeros(S, FRAME_DTYPE)
    par = np.zeros((S, 10)) if want_params else None
    mn = np.zeros((S, 2, 10, 10)) if want_mn else None
    eig = np.zeros((S, 10)) if want_params else None
    rc = lib().ago_fit_quadrics(_p(tree.xyz, C.c_float), _p(cam, C.c_int32), tree.xyz.shape[0], tree.h,
                                _p(indices, C.c_int), S, C.c_double(radius), C.byref(params), int(sum_perm),
                                frames.ctypes.data_as(C.POINTER(AgFrame)),
                                _p(par, C.c_double) if want_params else None,
                                _p(mn, C.c_double) if want_mn else None,
                                _p(eig, C.c_double) if want_params else None)
    if rc != 0:
        raise RuntimeError(_err())
    out = {"frames": frames}
    if want_params:
        out["params"] = par
        out["eigvals"] = eig
    if want_mn:
        out["MN"] = mn
    return out


class Hands:
    """Result handle: grasp records + points_for_learning + per-(sample,rotation) debug info."""

    def __init__(self, handle):
        self.h = C.c_void_p(handle)

    def __del__(self):
        if getattr(self, "h", None):
            lib().ago_hands_free(self.h)
            self.h = None

    def __len__(self):
        return lib().ago_hands_count(self.h)

    @property
    def grasps(self):
        n = len(self)
        if n == 0:
            return np.zeros(0, GRASP_DTYPE)
        ptr = lib().ago_hands_grasps(self.h)
        buf = C.string_at(ptr, n * C.sizeof(AgGrasp))
        return np.frombuffer(buf, dtype=GRASP_DTYPE).copy()

    def points(self, k):
        pts = C.POINTER(C.c_double)()
        cam = C.POINTER(C.c_int32)()
        m = C.c_int()
        if lib().ago_hands_points(self.h, k, C.byref(pts), C.byref(cam), C.byref(m)) != 0:
            raise RuntimeError(_err())
        if m.value == 0:
            return np.zeros((3, 0)), np.zeros(0, np.int32)
        P = np.ctypeslib.as_array(pts, shape=(m.value, 3)).T.copy()
        Cc = np.ctypeslib.as_array(cam, shape=(m.value,)).copy()
        return P, Cc

    def debug(self, n_samples):
        ptrs = [C.POINTER(C.c_int32)() for _ in range(5)]
        lib().ago_hands_debug(self.h, *[C.byref(p) for p in ptrs])
        names = ("status", "hand_idx", "depth_steps", "finger_mask")
        out = {nm: np.ctypeslib.as_array(p, shape=(n_samples * 8,)).reshape(n_samples, 8).copy()
               for nm, p in zip(names, ptrs[:4])} if n_samples else {nm: np.zeros((0, 8), np.int32) for nm in names}
        out["num_slab"] = np.ctypeslib.as_array(ptrs[4], shape=(n_samples,)).copy() if n_samples else np.zeros(0)
        return out

    def image(self, k, params: AgParams):
        img = np.zeros((AG_IMAGE_ROWS, AG_IMAGE_COLS), np.uint8)
        if lib().ago_grasp_image(self.h, k, C.byref(params), _p(img, C.c_uint8)) != 0:
            raise RuntimeError(_err())
        return img

    def image_cam(self, k, cam, params: AgParams):
        """training instance image (learning.cpp:375-400): cam = -1 all box points, 0 / 1 camera 1 / 2 only"""
        img = np.zeros((AG_IMAGE_ROWS, AG_IMAGE_COLS), np.uint8)
        if lib().ago_grasp_image_cam(self.h, k, int(cam), C.byref(params), _p(img, C.c_uint8)) != 0:
            raise RuntimeError(_err())
        return img

    def classify(self, svm, params: AgParams):
        n = len(self)
        keep = np.zeros(n, np.uint8)
        if lib().ago_classify(self.h, svm.h, C.byref(params), _p(keep, C.c_uint8)) != 0:
            raise RuntimeError(_err())
        return keep


def find_hands(tree: Tree, cam, indices, frames, sample_cam, cloud_normals, params: AgParams) -> Hands:
    cam = np.ascontiguousarray(cam, dtype=np.int32)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    frames = np.ascontiguousarray(frames)
    sample_cam = np.ascontiguousarray(sample_cam, dtype=np.int32)
    cn = None if cloud_normals is None else np.ascontiguousarray(cloud_normals, dtype=np.float64)
    h = lib().ago_find_hands(_p(tree.xyz, C.c_float), _p(cam, C.c_int32), tree.xyz.shape[0], tree.h,
                             _p(indices, C.c_int), indices.shape[0], frames.ctypes.data_as(C.POINTER(AgFrame)),
                             _p(sample_cam, C.c_int32), None if cn is None else _p(cn, C.c_double),
                             C.byref(params))
    return Hands(h)


def filter_hands(grasps, params: AgParams):
    g = np.ascontiguousarray(grasps)
    keep = np.zeros(g.shape[0], np.uint8)
    lib().ago_filter_hands(g.ctypes.data_as(C.POINTER(AgGrasp)), g.shape[0], C.byref(params), _p(keep, C.c_uint8))
    return keep


def points_image(pts3xm, binormal, surface, cam_pos):
    P = np.ascontiguousarray(np.asarray(pts3xm, dtype=np.float64).T)  # (m,3) row-major == 3xm column-major
    img = np.zeros((AG_IMAGE_ROWS, AG_IMAGE_COLS), np.uint8)
    b = np.ascontiguousarray(binormal, dtype=np.float64)
    s = np.ascontiguousarray(surface, dtype=np.float64)
    c = np.ascontiguousarray(cam_pos, dtype=np.float64)
    lib().ago_points_image(_p(P, C.c_double), P.shape[0], _p(b, C.c_double), _p(s, C.c_double), _p(c, C.c_double),
                           _p(img, C.c_uint8))
    return img


def hog(image):
    img = np.ascontiguousarray(image, dtype=np.uint8)
    assert img.shape == (AG_IMAGE_ROWS, AG_IMAGE_COLS)
    d = np.zeros(AG_HOG_DIM, np.float32)
    lib().ago_hog(_p(img, C.c_uint8), _p(d, C.c_float))
    return d


class Svm:
    def __init__(self, path):
        h = lib().ago_svm_load(str(path).encode())
        if not h:
            raise RuntimeError(_err())
        self.h = C.c_void_p(h)
        kt, vc, st, dg = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        rho, gm, c0 = C.c_double(), C.c_double(), C.c_double()
        lib().ago_svm_info(self.h, C.byref(kt), C.byref(vc), C.byref(st), C.byref(rho), C.byref(dg), C.byref(gm),
                           C.byref(c0))
        self.kernel, self.var_count, self.sv_total = kt.value, vc.value, st.value
        self.rho, self.degree, self.gamma, self.coef0 = rho.value, dg.value, gm.value, c0.value

    def __del__(self):
        if getattr(self, "h", None):
            lib().ago_svm_free(self.h)
            self.h = None

    @property
    def support_vectors(self):
        p = lib().ago_svm_sv(self.h)
        return np.ctypeslib.as_array(p, shape=(self.sv_total, self.var_count)).copy()

    @property
    def alpha(self):
        p = lib().ago_svm_alpha(self.h)
        return np.ctypeslib.as_array(p, shape=(self.sv_total,)).copy()

    def decision(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        return float(lib().ago_svm_decision(self.h, _p(x, C.c_float)))


def localize(points32, size_left, params: AgParams, indices=None, flags=0, svm=None, use_std_set=True):
    """Full CPU path.  Returns (Hands, times_ms dict, n_voxels)."""
    pts = np.ascontiguousarray(points32)
    idx = None if indices is None else np.ascontiguousarray(indices, dtype=np.int32)
    tm = np.zeros(7)
    nv = C.c_int()
    h = lib().ago_localize(pts.ctypes.data_as(C.c_void_p), pts.strides[0], pts.shape[0], int(size_left),
                           C.byref(params), None if idx is None else _p(idx, C.c_int),
                           0 if idx is None else idx.shape[0], int(flags), None if svm is None else svm.h,
                           int(use_std_set), _p(tm, C.c_double), C.byref(nv))
    if not h:
        raise RuntimeError(_err())
    names = ("preprocess", "tree", "normals_all", "quadrics", "hands", "classify", "total")
    return Hands(h), dict(zip(names, tm.tolist())), nv.value


def find_handles(grasps, min_inliers, min_length):
    """HandleSearch::findHandles + Handle on grasp records -> (handles: HANDLE_DTYPE array, inlier index arrays)"""
    from agile_grasp_b200.ctypes_defs import HANDLE_DTYPE
    g = np.ascontiguousarray(grasps)
    hp, ip_, nh, ni = C.c_void_p(), C.POINTER(C.c_int32)(), C.c_int(), C.c_int()
    rc = lib().ago_find_handles(g.ctypes.data_as(C.POINTER(AgGrasp)), g.shape[0], int(min_inliers), C.c_double(min_length),
                                C.byref(hp), C.byref(nh), C.byref(ip_), C.byref(ni))
    if rc:
        raise RuntimeError(_err())
    H = np.frombuffer(C.string_at(hp, nh.value * HANDLE_DTYPE.itemsize), dtype=HANDLE_DTYPE).copy()
    flat = np.ctypeslib.as_array(ip_, shape=(max(ni.value, 1),))[:ni.value].copy()
    lib().ago_free(hp)
    lib().ago_free(ip_)
    return H, [flat[h["inlier_offset"]:h["inlier_offset"] + h["n_inliers"]] for h in H]


def remove_plane(xyz, seed, max_iterations=100, thresh=0.01):
    """uses_clustering (localization.cpp:51-98): (keep mask, per-iteration inlier counts, refitted plane) or None"""
    X = np.ascontiguousarray(xyz, dtype=np.float32)
    keep = np.zeros(X.shape[0], np.uint8)
    counts = np.zeros(max_iterations, np.int32)
    plane = np.zeros(4)
    rc = lib().ago_remove_plane(_p(X, C.c_float), X.shape[0], C.c_uint64(seed), int(max_iterations), C.c_double(thresh),
                                _p(keep, C.c_uint8), _p(counts, C.c_int32), _p(plane, C.c_double))
    return None if rc else (keep.astype(bool), counts, plane)


def glibc_rand(seed, n):
    out = np.zeros(n, np.int32)
    lib().ago_glibc_rand(C.c_uint32(seed), int(n), _p(out, C.c_int32))
    return out


def draw_samples(n, num_samples, seed):
    out = np.zeros(min(n, num_samples), np.int32)
    k = lib().ago_draw_samples(int(n), int(num_samples), C.c_uint64(seed), _p(out, C.c_int32))
    return out[:k]
