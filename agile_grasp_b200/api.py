"""Thin ctypes binding of libag_b200.so (the extern "C" ABI in include/ag_b200.h).

This is host-side plumbing for tests and bench.py only; the C++ drop-in lives in
include/agile_grasp/*.h.  There is no fallback: if the CUDA library is missing or no GPU is
visible, construction fails loudly.
"""
import ctypes as C
import os

import numpy as np

from .ctypes_defs import (AG_HOG_DIM, AG_IMAGE_COLS, AG_IMAGE_ROWS, AG_IMAGE_WORDS, FRAME_DTYPE, GRASP_DTYPE, AgFrame,
                          AgGrasp, AgParams, AgTimings)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libag_b200.so")
_LIB = None

EXPORTS = [
    "ag_last_error", "ag_default_params", "ag_create", "ag_destroy", "ag_set_params", "ag_get_params",
    "ag_get_timings", "ag_set_stage_timing", "ag_free", "ag_svm_load", "ag_svm_free", "ag_svm_info", "ag_localize", "ag_localize_device",
    "ag_classify", "ag_set_svm", "ag_set_export_buffer", "ag_get_points", "ag_get_images", "ag_get_normals", "ag_train_features", "ag_remove_plane", "ag_preprocess", "ag_set_cloud", "ag_radius_search",
    "ag_fit_quadrics", "ag_hand_sweep", "ag_sweep_debug", "ag_hog_svm",
    "ag_find_handles", "ag_load_pcd", "ag_localize_batch", "ag_gather_slot_bytes", "ag_gather_create", "ag_gather_connect", "ag_gather_wait", "ag_gather_result", "ag_gather_destroy",
]


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with __graft_entry__.build() "
                           "(make -C agile_grasp_b200/csrc). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, ip, dp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double)
    L.ag_last_error.restype = C.c_char_p
    L.ag_create.restype = vp
    L.ag_create.argtypes = [C.c_int]
    L.ag_destroy.argtypes = [vp]
    L.ag_set_params.argtypes = [vp, C.POINTER(AgParams)]
    L.ag_get_params.argtypes = [vp, C.POINTER(AgParams)]
    L.ag_get_timings.argtypes = [vp, C.POINTER(AgTimings)]
    L.ag_set_stage_timing.argtypes = [vp, C.c_int]
    L.ag_free.argtypes = [vp]
    L.ag_svm_load.restype = vp
    L.ag_svm_load.argtypes = [C.c_char_p]
    L.ag_svm_free.argtypes = [vp]
    L.ag_svm_info.argtypes = [vp, ip, ip, ip, dp]
    loc_args = [vp, vp, C.c_int, C.c_int, C.c_int, ip, C.c_int, C.c_uint, C.POINTER(C.POINTER(AgGrasp)), ip]
    L.ag_localize.argtypes = loc_args
    L.ag_localize_device.argtypes = loc_args
    L.ag_classify.argtypes = [vp, vp, C.POINTER(AgGrasp), C.c_int, C.POINTER(C.c_uint8)]
    L.ag_set_svm.argtypes = [vp, vp]
    L.ag_set_export_buffer.argtypes = [vp, vp, C.c_size_t]
    L.ag_get_points.argtypes = [vp, C.c_int, C.POINTER(C.POINTER(C.c_double)), C.POINTER(C.POINTER(C.c_int32)), ip]
    L.ag_get_images.argtypes = [vp, C.POINTER(C.POINTER(C.c_uint32)), ip]
    L.ag_get_normals.argtypes = [vp, dp, C.c_int]
    L.ag_train_features.argtypes = [vp, C.POINTER(AgGrasp), C.c_int, C.POINTER(C.c_float)]
    L.ag_preprocess.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.POINTER(C.POINTER(C.c_float)),
                                C.POINTER(C.POINTER(C.c_int32)), ip]
    L.ag_set_cloud.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_int32), C.c_int]
    L.ag_remove_plane.argtypes = [vp, C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.POINTER(C.c_int32)), ip]
    L.ag_radius_search.argtypes = [vp, C.POINTER(C.c_float), C.c_double, C.POINTER(C.POINTER(C.c_int32)), ip]
    L.ag_fit_quadrics.argtypes = [vp, ip, C.c_int, C.c_double, C.POINTER(AgFrame)]
    L.ag_hand_sweep.argtypes = [vp, ip, C.c_int, C.POINTER(AgFrame), dp, C.c_uint, C.POINTER(C.POINTER(AgGrasp)), ip]
    L.ag_sweep_debug.argtypes = [vp, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.ag_hog_svm.argtypes = [vp, vp, C.POINTER(C.c_uint32), C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.ag_find_handles.argtypes = [vp, C.POINTER(AgGrasp), C.c_int, C.c_int, C.c_double, C.POINTER(C.c_void_p), ip,
                                  C.POINTER(C.POINTER(C.c_int32)), ip]
    L.ag_localize_batch.argtypes = [vp, C.c_int, C.POINTER(C.c_void_p), ip, ip, ip, C.c_uint,
                                    C.POINTER(C.POINTER(AgGrasp)), ip]
    L.ag_load_pcd.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), ip, ip, ip]
    L.ag_gather_slot_bytes.restype = C.c_size_t
    L.ag_gather_slot_bytes.argtypes = [C.c_int]
    L.ag_gather_create.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_char_p]
    L.ag_gather_connect.argtypes = [vp, C.c_char_p]
    L.ag_gather_wait.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    L.ag_gather_result.argtypes = [vp, C.POINTER(C.c_int32), ip, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    L.ag_gather_destroy.argtypes = [vp]
    _LIB = L
    return L


class AgError(RuntimeError):
    pass


def _check(rc):
    if rc != 0:
        raise AgError(f"[{rc}] " + lib().ag_last_error().decode())


def _grasps_from(ptr, n):
    # one flat copy out of the malloc'ed list (a field-wise copy of the structured dtype costs ten times as much)
    out = np.empty(n, GRASP_DTYPE)
    if n:
        C.memmove(out.ctypes.data, ptr, n * C.sizeof(AgGrasp))
    lib().ag_free(ptr)
    return out


def pack_images(images_u8):
    """(n, 80, 100) uint8 (0/255) -> (n, 250) uint32 bit images, bit index row*100+col."""
    imgs = np.asarray(images_u8).reshape(-1, AG_IMAGE_ROWS * AG_IMAGE_COLS) != 0
    bits = np.packbits(imgs, axis=1, bitorder="little")
    return np.ascontiguousarray(bits).view(np.uint32).reshape(-1, AG_IMAGE_WORDS)


def unpack_images(bits_u32):
    b = np.ascontiguousarray(bits_u32, dtype=np.uint32).view(np.uint8)
    px = np.unpackbits(b.reshape(-1, AG_IMAGE_WORDS * 4), axis=1, bitorder="little")
    return (px.reshape(-1, AG_IMAGE_ROWS, AG_IMAGE_COLS) * 255).astype(np.uint8)


class Svm:
    def __init__(self, path):
        h = lib().ag_svm_load(str(path).encode())
        if not h:
            raise AgError(lib().ag_last_error().decode())
        self.h = C.c_void_p(h)
        kt, vc, st, rho = C.c_int(), C.c_int(), C.c_int(), C.c_double()
        lib().ag_svm_info(self.h, C.byref(kt), C.byref(vc), C.byref(st), C.byref(rho))
        self.kernel, self.var_count, self.sv_total, self.rho = kt.value, vc.value, st.value, rho.value

    def __del__(self):
        if getattr(self, "h", None):
            lib().ag_svm_free(self.h)
            self.h = None


def load_pcd(path):
    """ag_load_pcd: PCD file -> (n, 8) float32 view of pcl::PointXYZRGBA records (x y z _ rgba-bits _ _ _), width, height"""
    p, n, w, h = C.c_void_p(), C.c_int(), C.c_int(), C.c_int()
    _check(lib().ag_load_pcd(str(path).encode(), C.byref(p), C.byref(n), C.byref(w), C.byref(h)))
    arr = np.frombuffer(C.string_at(p, n.value * 32), dtype=np.float32).reshape(n.value, 8).copy()
    lib().ag_free(p)
    return arr, w.value, h.value


class Context:
    """One GPU context = one `Localization` object of the reference."""

    def __init__(self, device=0, params: AgParams = None, stage_timing=True):
        h = lib().ag_create(int(device))
        if not h:
            raise AgError(lib().ag_last_error().decode())
        self.h = C.c_void_p(h)
        # (tests and tools read the per-stage times; the library default — and bench.py's headline — is off)
        lib().ag_set_stage_timing(self.h, 1 if stage_timing else 0)
        if params is not None:
            self.set_params(params)

    def close(self):
        if getattr(self, "h", None):
            lib().ag_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def set_params(self, p: AgParams):
        _check(lib().ag_set_params(self.h, C.byref(p)))

    def set_stage_timing(self, on):
        lib().ag_set_stage_timing(self.h, 1 if on else 0)

    def timings(self):
        t = AgTimings()
        lib().ag_get_timings(self.h, C.byref(t))
        return {n: getattr(t, n) for n, _ in AgTimings._fields_}

    # ---- full path
    def localize(self, points32, size_left, indices=None, flags=0):
        pts = np.ascontiguousarray(points32)
        idx = None if indices is None else np.ascontiguousarray(indices, dtype=np.int32)
        out = C.POINTER(AgGrasp)()
        n = C.c_int()
        _check(lib().ag_localize(self.h, pts.ctypes.data_as(C.c_void_p), pts.strides[0], pts.shape[0], int(size_left),
                                 None if idx is None else idx.ctypes.data_as(C.POINTER(C.c_int)),
                                 0 if idx is None else idx.shape[0], int(flags), C.byref(out), C.byref(n)))
        return _grasps_from(out, n.value)

    def localize_batch(self, clouds, size_lefts, flags=0):
        """ag_localize_batch: list of (n, 8) float32 clouds -> list of grasp arrays (samples drawn from params.seed)"""
        k = len(clouds)
        arrs = [np.ascontiguousarray(p) for p in clouds]
        ptrs = (C.c_void_p * k)(*[a.ctypes.data for a in arrs])
        strides = (C.c_int * k)(*[a.strides[0] for a in arrs])
        n_in = (C.c_int * k)(*[a.shape[0] for a in arrs])
        sl = (C.c_int * k)(*[int(v) for v in size_lefts])
        outs = (C.POINTER(AgGrasp) * k)()
        n_out = (C.c_int * k)()
        _check(lib().ag_localize_batch(self.h, k, ptrs, strides, n_in, sl, int(flags), outs, n_out))
        return [_grasps_from(outs[i], n_out[i]) for i in range(k)]

    def localize_device(self, dev_ptr, stride, n_in, size_left, indices=None, flags=0):
        idx = None if indices is None else np.ascontiguousarray(indices, dtype=np.int32)
        out = C.POINTER(AgGrasp)()
        n = C.c_int()
        _check(lib().ag_localize_device(self.h, C.c_void_p(dev_ptr), int(stride), int(n_in), int(size_left),
                                        None if idx is None else idx.ctypes.data_as(C.POINTER(C.c_int)),
                                        0 if idx is None else idx.shape[0], int(flags), C.byref(out), C.byref(n)))
        return _grasps_from(out, n.value)

    def classify(self, svm: Svm, grasps):
        g = np.ascontiguousarray(grasps)
        keep = np.zeros(g.shape[0], np.uint8)
        _check(lib().ag_classify(self.h, svm.h, g.ctypes.data_as(C.POINTER(AgGrasp)), g.shape[0],
                                 keep.ctypes.data_as(C.POINTER(C.c_uint8))))
        return g, keep

    def set_svm(self, svm):
        """fuse scoring into localize(); pass None to detach"""
        self._svm = svm  # keep the model alive while attached
        _check(lib().ag_set_svm(self.h, None if svm is None else svm.h))

    def points(self, image_id):
        """points_for_learning (3 x m) and the camera source of every column, reference column order"""
        pts = C.POINTER(C.c_double)()
        cam = C.POINTER(C.c_int32)()
        m = C.c_int()
        _check(lib().ag_get_points(self.h, int(image_id), C.byref(pts), C.byref(cam), C.byref(m)))
        if m.value == 0:
            P3, Cm = np.zeros((3, 0)), np.zeros(0, np.int32)
        else:
            P3 = np.ctypeslib.as_array(pts, shape=(m.value, 3)).T.copy()
            Cm = np.ctypeslib.as_array(cam, shape=(m.value,)).copy()
        lib().ag_free(pts)
        lib().ag_free(cam)
        return P3, Cm

    def find_handles(self, grasps, min_inliers, min_length):
        """HandleSearch::findHandles on grasp records -> (handles: HANDLE_DTYPE array, list of inlier index arrays)"""
        from .ctypes_defs import HANDLE_DTYPE
        g = np.ascontiguousarray(grasps)
        hp, ip_, nh, ni = C.c_void_p(), C.POINTER(C.c_int32)(), C.c_int(), C.c_int()
        _check(lib().ag_find_handles(self.h, g.ctypes.data_as(C.POINTER(AgGrasp)), g.shape[0], int(min_inliers),
                                     float(min_length), C.byref(hp), C.byref(nh), C.byref(ip_), C.byref(ni)))
        H = np.frombuffer(C.string_at(hp, nh.value * HANDLE_DTYPE.itemsize), dtype=HANDLE_DTYPE).copy()
        flat = np.ctypeslib.as_array(ip_, shape=(max(ni.value, 1),))[:ni.value].copy()
        lib().ag_free(hp)
        lib().ag_free(ip_)
        return H, [flat[h["inlier_offset"]:h["inlier_offset"] + h["n_inliers"]] for h in H]

    # ---- peer gather: the grasp-list all-gather fused into the export kernel (include/ag_b200.h)
    def gather_create(self, num_samples, world, rank):
        """allocates this rank's gather buffer; returns its 64-byte CUDA IPC handle"""
        hd = C.create_string_buffer(64)
        _check(lib().ag_gather_create(self.h, int(num_samples), int(world), int(rank), hd))
        self._gather_world = int(world)
        return hd.raw

    def gather_connect(self, handles):
        """handles: list of every rank's 64-byte IPC handle, in rank order"""
        _check(lib().ag_gather_connect(self.h, b"".join(handles)))

    def gather_result(self, copy=True):
        """the merged grasp list of the last localize() over all ranks (sample-major, as one unsharded call would
        return it): (n_hyp per rank, merged records from the mapped host copy, device address of the merged list)"""
        n = (C.c_int32 * self._gather_world)()
        tot, dptr, hptr = C.c_int(), C.c_void_p(), C.c_void_p()
        _check(lib().ag_gather_result(self.h, n, C.byref(tot), C.byref(dptr), C.byref(hptr) if copy else None))
        recs = None
        if copy:
            if tot.value == 0:
                recs = np.zeros(0, GRASP_DTYPE)
            else:
                recs = np.frombuffer(C.string_at(hptr.value, tot.value * C.sizeof(AgGrasp)), dtype=GRASP_DTYPE).copy()
        return list(n), recs, dptr.value

    def gather_wait(self):
        """the exchange completes inside localize(); reports it: (n_hyp per rank, device address of slot 0, slot bytes)"""
        n = (C.c_int32 * self._gather_world)()
        ptr, sb = C.c_void_p(), C.c_size_t()
        _check(lib().ag_gather_wait(self.h, n, C.byref(ptr), C.byref(sb)))
        return list(n), ptr.value, sb.value

    def set_export_buffer(self, dev_ptr, nbytes):
        """device buffer that receives [n_hyp, n_vox, n_samples, error][records] on every localize()"""
        _check(lib().ag_set_export_buffer(self.h, None if not dev_ptr else C.c_void_p(dev_ptr), int(nbytes)))

    def images(self):
        bits = C.POINTER(C.c_uint32)()
        n = C.c_int()
        _check(lib().ag_get_images(self.h, C.byref(bits), C.byref(n)))
        if n.value == 0:
            out = np.zeros((0, AG_IMAGE_WORDS), np.uint32)
        else:
            out = np.ctypeslib.as_array(bits, shape=(n.value, AG_IMAGE_WORDS)).copy()
        lib().ag_free(bits)
        return out

    def train_features(self, grasps):
        """HOG descriptors of the three training instances of each hypothesis (own image, camera 1 only, camera 2
        only): (n, 3, 3528) float32"""
        g = np.ascontiguousarray(grasps)
        out = np.zeros((g.shape[0], 3, AG_HOG_DIM), np.float32)
        _check(lib().ag_train_features(self.h, g.ctypes.data_as(C.POINTER(AgGrasp)), g.shape[0],
                                       out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def normals(self, n):
        """cloud_normals_ of the last localize / sweep call: (n, 3) float64, zero where no normal was computed"""
        out = np.zeros((int(n), 3), np.float64)
        _check(lib().ag_get_normals(self.h, out.ctypes.data_as(C.POINTER(C.c_double)), int(n)))
        return out

    # ---- stages
    def preprocess(self, points32, size_left):
        pts = np.ascontiguousarray(points32)
        xyz = C.POINTER(C.c_float)()
        cam = C.POINTER(C.c_int32)()
        n = C.c_int()
        _check(lib().ag_preprocess(self.h, pts.ctypes.data_as(C.c_void_p), pts.strides[0], pts.shape[0],
                                   int(size_left), C.byref(xyz), C.byref(cam), C.byref(n)))
        if n.value == 0:
            X, Cm = np.zeros((0, 3), np.float32), np.zeros(0, np.int32)
        else:
            X = np.ctypeslib.as_array(xyz, shape=(n.value, 3)).copy()
            Cm = np.ctypeslib.as_array(cam, shape=(n.value,)).copy()
        lib().ag_free(xyz)
        lib().ag_free(cam)
        return X, Cm

    def remove_plane(self):
        """uses_clustering as a stage: the current voxelised cloud without its dominant plane -> (xyz, cam)"""
        xyz = C.POINTER(C.c_float)()
        cam = C.POINTER(C.c_int32)()
        n = C.c_int()
        _check(lib().ag_remove_plane(self.h, C.byref(xyz), C.byref(cam), C.byref(n)))
        if n.value == 0:
            X, Cm = np.zeros((0, 3), np.float32), np.zeros(0, np.int32)
        else:
            X = np.ctypeslib.as_array(xyz, shape=(n.value, 3)).copy()
            Cm = np.ctypeslib.as_array(cam, shape=(n.value,)).copy()
        lib().ag_free(xyz)
        lib().ag_free(cam)
        return X, Cm

    def set_cloud(self, xyz, cam=None):
        X = np.ascontiguousarray(xyz, dtype=np.float32)
        Cm = None if cam is None else np.ascontiguousarray(cam, dtype=np.int32)
        _check(lib().ag_set_cloud(self.h, X.ctypes.data_as(C.POINTER(C.c_float)),
                                  None if Cm is None else Cm.ctypes.data_as(C.POINTER(C.c_int32)), X.shape[0]))

    def radius_search(self, q, radius):
        qq = np.ascontiguousarray(q, dtype=np.float32)
        idx = C.POINTER(C.c_int32)()
        n = C.c_int()
        _check(lib().ag_radius_search(self.h, qq.ctypes.data_as(C.POINTER(C.c_float)), float(radius), C.byref(idx),
                                      C.byref(n)))
        out = np.ctypeslib.as_array(idx, shape=(n.value,)).copy() if n.value else np.zeros(0, np.int32)
        lib().ag_free(idx)
        return out

    def fit_quadrics(self, indices, radius):
        idx = np.ascontiguousarray(indices, dtype=np.int32)
        frames = np.zeros(idx.shape[0], FRAME_DTYPE)
        _check(lib().ag_fit_quadrics(self.h, idx.ctypes.data_as(C.POINTER(C.c_int)), idx.shape[0], float(radius),
                                     frames.ctypes.data_as(C.POINTER(AgFrame))))
        return frames

    def hand_sweep(self, indices, frames, cloud_normals=None, flags=0):
        idx = np.ascontiguousarray(indices, dtype=np.int32)
        fr = np.ascontiguousarray(frames)
        cn = None if cloud_normals is None else np.ascontiguousarray(cloud_normals, dtype=np.float64)
        out = C.POINTER(AgGrasp)()
        n = C.c_int()
        _check(lib().ag_hand_sweep(self.h, idx.ctypes.data_as(C.POINTER(C.c_int)), idx.shape[0],
                                   fr.ctypes.data_as(C.POINTER(AgFrame)),
                                   None if cn is None else cn.ctypes.data_as(C.POINTER(C.c_double)), int(flags),
                                   C.byref(out), C.byref(n)))
        return _grasps_from(out, n.value)

    def sweep_debug(self, n_samples):
        slab = np.zeros(n_samples, np.int32)
        dbg = np.zeros((n_samples, 8), np.int32)
        _check(lib().ag_sweep_debug(self.h, n_samples, slab.ctypes.data_as(C.POINTER(C.c_int32)),
                                    dbg.ctypes.data_as(C.POINTER(C.c_int32))))
        u = dbg.view(np.uint32)
        return {"num_slab": slab, "status": (u & 0xF).astype(np.int32),
                "hand_idx": np.where((u & 0xF) == 2, (u >> 4) & 0xF, -1).astype(np.int32),
                "depth_steps": ((u >> 8) & 0xF).astype(np.int32), "finger_mask": (u >> 12).astype(np.int32)}

    def hog_svm(self, svm: Svm, images_bits, want_descriptors=False):
        bits = np.ascontiguousarray(images_bits, dtype=np.uint32).reshape(-1, AG_IMAGE_WORDS)
        n = bits.shape[0]
        scores = np.zeros(n, np.float32)
        desc = np.zeros((n, AG_HOG_DIM), np.float32) if want_descriptors else None
        _check(lib().ag_hog_svm(self.h, svm.h, bits.ctypes.data_as(C.POINTER(C.c_uint32)), n,
                                None if desc is None else desc.ctypes.data_as(C.POINTER(C.c_float)),
                                scores.ctypes.data_as(C.POINTER(C.c_float))))
        return scores, desc
