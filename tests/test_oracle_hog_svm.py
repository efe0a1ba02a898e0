"""HOG + SVM restatement pinned against cv2 itself and the cv2-made fixtures."""
import os

import numpy as np
import pytest

from agile_grasp_b200 import api

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_hog_matches_cv2_fixture_bit_for_bit(oracle):
    z = np.load(os.path.join(GOLD, "hog_svm_cv2.npz"))
    imgs = api.unpack_images(z["images_bits"])
    for k in range(len(imgs)):
        d = oracle.hog(imgs[k])
        assert (d.view(np.uint32) == z["descriptors"][k].view(np.uint32)).all()


def test_svm_decision_matches_cv2_fixture(oracle, linear_svm_path):
    z = np.load(os.path.join(GOLD, "hog_svm_cv2.npz"))
    svm = oracle.Svm(linear_svm_path)
    assert (svm.kernel, svm.var_count, svm.sv_total) == (0, 3528, 1)
    assert svm.rho == -3.1383255947302025e-01
    for k in range(len(z["descriptors"])):
        s = np.float32(svm.decision(z["descriptors"][k]))
        assert s == z["svm_raw"][k]
        # CvSVM::predict: label +1 <=> decision value <= 0 (learning.cpp:226 keeps prediction == 1)
        assert (1.0 if s <= 0 else -1.0) == z["svm_label"][k]


def test_hog_matches_live_cv2(oracle):
    cv2 = pytest.importorskip("cv2")
    hog = cv2.HOGDescriptor((64, 64), (16, 16), (8, 8), (8, 8), 9, 1, -1.0, 0, 0.2, True, 64, False)
    rng = np.random.default_rng(7)
    for t in range(20):
        img = np.zeros((80, 100), np.uint8)
        img[rng.random((80, 100)) < rng.uniform(0.01, 0.6)] = 255
        d = hog.compute(img, (32, 32), (0, 0)).ravel()
        assert d.shape == (3528,)
        assert (oracle.hog(img).view(np.uint32) == d.view(np.uint32)).all()
    # empty and full images: zero gradient everywhere -> zero descriptor
    for v in (0, 255):
        img = np.full((80, 100), v, np.uint8)
        assert (oracle.hog(img) == 0).all() and (hog.compute(img, (32, 32), (0, 0)) == 0).all()


def test_cart_to_polar_table_is_what_cv2_computes():
    cv2 = pytest.importorskip("cv2")
    s = np.float32(np.sqrt(np.float32(255)))
    mag_bits = [0x41b4aa5a, 0x417f7fe0, 0x41b4aa5a, 0x417f7fe0, 0x0, 0x417f7fe0, 0x41b4aa5a, 0x417f7fe0, 0x41b4aa5a]
    ang_bits = [0x407b5116, 0x4096cbe4, 0x40afef3d, 0x40490fdb, 0x0, 0x0, 0x4016ce9f, 0x3fc90fdb, 0x3f4904f0]
    dx = np.tile(np.array([-1, 0, 1] * 3, np.float32) * s, 8).reshape(1, -1)
    dy = np.tile(np.repeat(np.array([-1, 0, 1], np.float32), 3) * s, 8).reshape(1, -1)
    mag, ang = cv2.cartToPolar(dx, dy)
    assert (mag.view(np.uint32).reshape(8, 9) == np.array(mag_bits, np.uint32)).all()
    assert (ang.view(np.uint32).reshape(8, 9) == np.array(ang_bits, np.uint32)).all()


def test_poly_svm_file_round_trip(oracle, tmp_path):
    """POLY kernel (the launch-file default models) in the OpenCV 2.4 YAML format, checked against cv2."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    nsv = 5
    sv = (rng.random((nsv, 3528)) * (rng.random((nsv, 3528)) < 0.1)).astype(np.float32)
    alpha = rng.normal(size=nsv)
    path = tmp_path / "poly_svm"
    write_opencv_svm(path, sv, alpha, rho=0.25, kernel="POLY", degree=2, gamma=1.0, coef0=0.0)
    svm = oracle.Svm(path)
    assert (svm.kernel, svm.sv_total, svm.degree) == (1, nsv, 2)
    csvm = cv2.ml.SVM_load(str(path))
    for t in range(4):
        x = (rng.random(3528) * 0.2).astype(np.float32)
        raw = csvm.predict(x.reshape(1, -1), flags=cv2.ml.STAT_MODEL_RAW_OUTPUT)[1][0, 0]
        assert abs(svm.decision(x) - raw) <= 1e-6 * max(1.0, abs(raw))


def write_opencv_svm(path, sv, alpha, rho, kernel="LINEAR", degree=2, gamma=1.0, coef0=0.0):
    nsv, dim = sv.shape
    kern = "{ type:LINEAR }" if kernel == "LINEAR" else "{ type:POLY, degree:%r, gamma:%r, coef0:%r }" % (
        float(degree), float(gamma), float(coef0))
    rows = []
    for r in sv:
        rows.append("      - [ " + ", ".join("%.8e" % v for v in r) + " ]")
    txt = f"""%YAML:1.0
my_svm: !!opencv-ml-svm
   svm_type: C_SVC
   kernel: {kern}
   C: 1.
   term_criteria: {{ epsilon:1.1920928955078125e-07, iterations:1000 }}
   var_all: {dim}
   var_count: {dim}
   class_count: 2
   class_labels: !!opencv-matrix
      rows: 1
      cols: 2
      dt: i
      data: [ -1, 1 ]
   sv_total: {nsv}
   support_vectors:
""" + "\n".join(rows) + f"""
   decision_functions:
      -
         sv_count: {nsv}
         rho: {rho!r}
         alpha: [ {", ".join("%.17e" % a for a in alpha)} ]
         index: [ {", ".join(str(i) for i in range(nsv))} ]
"""
    with open(path, "w") as f:
        f.write(txt)


@pytest.mark.parametrize("name", ["svm_032015_20_20_same", "svm_032015_20_20"])
def test_shipped_poly_models_match_cv2_fixture(oracle, poly_svm_paths, name):
    """The reference's shipped POLY models (588 / 1190 support vectors; svm_032015_20_20_same is the launch-file
    default, launch/single_camera_grasps.launch:6): descriptors and raw decision values of the committed
    cv2-made fixture (tools/make_golden_poly.py), bit for bit."""
    z = np.load(os.path.join(GOLD, "poly_svm_cv2.npz"))
    svm = oracle.Svm(poly_svm_paths[name])
    assert svm.kernel == 1 and svm.degree == 2 and svm.var_count == 3528
    assert svm.sv_total == (588 if name.endswith("same") else 1190)
    imgs = api.unpack_images(z["images_bits"])
    raw = z["raw_" + name]
    for t in range(len(imgs)):
        d = oracle.hog(imgs[t])
        assert (d.view(np.uint32) == z["descriptors"][t].view(np.uint32)).all()
        assert np.float32(svm.decision(d)).view(np.uint32) == raw[t].view(np.uint32), t
    assert ((raw <= 0) == (z["label_" + name] == 1)).all()  # CvSVM::predict: label +1 <=> sum <= 0
