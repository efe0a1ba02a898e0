"""ctypes mirrors of the POD structs in include/ag_b200.h (single source of truth: the header)."""
import ctypes as C

import numpy as np

AG_IMAGE_COLS = 100
AG_IMAGE_ROWS = 80
AG_IMAGE_WORDS = 250
AG_HOG_DIM = 3528
AG_FLAG_CALC_ANTIPODAL = 1
AG_FLAG_KEEP_POINTS = 2


class AgParams(C.Structure):
    _fields_ = [
        ("finger_width", C.c_double),
        ("hand_outer_diameter", C.c_double),
        ("hand_depth", C.c_double),
        ("hand_height", C.c_double),
        ("init_bite", C.c_double),
        ("workspace", C.c_double * 6),
        ("cam_tf_left", C.c_double * 16),
        ("cam_tf_right", C.c_double * 16),
        ("nn_radius_taubin", C.c_double),
        ("nn_radius_hands", C.c_double),
        ("nn_radius_normals", C.c_double),
        ("voxel_size", C.c_double),
        ("num_samples", C.c_int32),
        ("num_threads", C.c_int32),
        ("deterministic_normals", C.c_int32),
        ("filters_boundaries", C.c_int32),
        ("fix_cam_source", C.c_int32),
        ("shard_interleave", C.c_int32),
        ("seed", C.c_uint64),
        ("shard_index", C.c_int32),
        ("shard_count", C.c_int32),
    ]


class AgGrasp(C.Structure):
    _fields_ = [
        ("axis", C.c_double * 3),
        ("approach", C.c_double * 3),
        ("binormal", C.c_double * 3),
        ("bottom", C.c_double * 3),
        ("surface", C.c_double * 3),
        ("width", C.c_double),
        ("score", C.c_float),
        ("sample_index", C.c_int32),
        ("sample_slot", C.c_int32),
        ("orientation", C.c_int32),
        ("cam_source", C.c_int32),
        ("num_points", C.c_int32),
        ("image_id", C.c_int32),
        ("half_antipodal", C.c_uint8),
        ("full_antipodal", C.c_uint8),
        ("label", C.c_uint8),
        ("reserved", C.c_uint8),
    ]


class AgFrame(C.Structure):
    _fields_ = [
        ("normal", C.c_double * 3),
        ("axis", C.c_double * 3),
        ("binormal", C.c_double * 3),
        ("num_neighbors", C.c_int32),
        ("majority_cam", C.c_int32),
    ]


class AgTimings(C.Structure):
    _fields_ = [(n, C.c_float) for n in (
        "h2d_ms", "preprocess_ms", "grid_ms", "normals_all_ms", "quadric_ms", "sweep_ms", "hog_svm_ms",
        "d2h_ms", "total_ms")] + [(n, C.c_int32) for n in ("n_in", "n_voxels", "n_samples", "n_hyp")] + [
        (n, C.c_int64) for n in ("taubin_neighbor_points", "hand_neighbor_points", "taubin_candidates",
                                 "hand_candidates")] + [("moments_ms", C.c_float), ("axes_ms", C.c_float),
                                                        ("kernel_launches", C.c_int32), ("search_ms", C.c_float)]


GRASP_DTYPE = np.dtype([
    ("axis", "<f8", 3), ("approach", "<f8", 3), ("binormal", "<f8", 3), ("bottom", "<f8", 3),
    ("surface", "<f8", 3), ("width", "<f8"), ("score", "<f4"), ("sample_index", "<i4"),
    ("sample_slot", "<i4"), ("orientation", "<i4"), ("cam_source", "<i4"), ("num_points", "<i4"),
    ("image_id", "<i4"), ("half_antipodal", "u1"), ("full_antipodal", "u1"), ("label", "u1"),
    ("reserved", "u1")], align=True)
HANDLE_DTYPE = np.dtype([("axis", "<f8", 3), ("center", "<f8", 3), ("approach", "<f8", 3), ("binormal", "<f8", 3),
                         ("hands_center", "<f8", 3), ("width", "<f8"), ("n_inliers", "<i4"),
                         ("inlier_offset", "<i4")], align=True)
assert HANDLE_DTYPE.itemsize == 136
FRAME_DTYPE = np.dtype([("normal", "<f8", 3), ("axis", "<f8", 3), ("binormal", "<f8", 3),
                        ("num_neighbors", "<i4"), ("majority_cam", "<i4")], align=True)
assert GRASP_DTYPE.itemsize == C.sizeof(AgGrasp) == 160, (GRASP_DTYPE.itemsize, C.sizeof(AgGrasp))
assert FRAME_DTYPE.itemsize == C.sizeof(AgFrame) == 80

# Baxter camera matrices used by every reference executable (src/nodes/test.cpp:47-56)
BASE_TF = np.array([[0, 0.445417, 0.895323, 0.215], [1, 0, 0, -0.015], [0, 0.895323, -0.445417, 0.23],
                    [0, 0, 0, 1]], dtype=np.float64)
SQRT_TF = np.array([[0.9366, -0.0162, 0.3500, -0.2863], [0.0151, 0.9999, 0.0058, 0.0058],
                    [-0.3501, -0.0002, 0.9367, 0.0554], [0, 0, 0, 1]], dtype=np.float64)


def default_params(**kw) -> AgParams:
    """Defaults of find_grasps.cpp:7-23 / test.cpp:78-82, single camera = launch-file camera_pose."""
    p = AgParams()
    p.finger_width = 0.01
    p.hand_outer_diameter = 0.09
    p.hand_depth = 0.06
    p.hand_height = 0.02
    p.init_bite = 0.01
    p.workspace[:] = [-10, 10, -10, 10, -10, 10]
    p.cam_tf_left[:] = BASE_TF.reshape(-1).tolist()
    p.cam_tf_right[:] = BASE_TF.reshape(-1).tolist()  # single camera: right = left (App. B#1)
    p.nn_radius_taubin = 0.03
    p.nn_radius_hands = 0.08
    p.nn_radius_normals = 0.01
    p.voxel_size = 0.003
    p.num_samples = 2000
    p.num_threads = 1
    p.deterministic_normals = 1
    p.filters_boundaries = 0
    p.fix_cam_source = 0
    p.seed = 20150320
    for k, v in kw.items():
        if k in ("workspace", "cam_tf_left", "cam_tf_right"):
            getattr(p, k)[:] = np.asarray(v, dtype=np.float64).reshape(-1).tolist()
        else:
            setattr(p, k, v)
    return p
