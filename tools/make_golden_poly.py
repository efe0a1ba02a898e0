"""Fixture for the reference's shipped POLY-kernel SVM models (svm_032015_20_20_same = the launch-file default,
launch/single_camera_grasps.launch:6; svm_032015_20_20), run in the build container:

 * tests/golden/svm_032015_20_20{,_same}.xz — the two model files, xz-compressed copies of the reference's on-disk
   YAML (data, 11 / 25 MB uncompressed); tests decompress them and load them through ag_svm_load / the oracle.
 * tests/golden/poly_svm_cv2.npz — produced by cv2 4.13 itself: grasp images (the 12 of hog_svm_cv2.npz plus the
   grasp images of the pipeline fixture), their cv2.HOGDescriptor descriptors and the raw decision values
   cv2.ml.SVM_load(model).predict(RAW_OUTPUT) of both models (CvSVM::predict, learning.cpp:225).
"""
import lzma, os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cv2
from agile_grasp_b200 import api

G = os.path.join(ROOT, "tests", "golden")
hog = cv2.HOGDescriptor((64, 64), (16, 16), (8, 8), (8, 8), 9, 1, -1.0, 0, 0.2, True, 64, False)
bits = np.concatenate([np.load(os.path.join(G, "hog_svm_cv2.npz"))["images_bits"],
                       np.load(os.path.join(G, "pipeline_small.npz"), allow_pickle=True)["images_bits"]])
imgs = api.unpack_images(bits)
desc = np.stack([hog.compute(im, (32, 32), (0, 0)).ravel() for im in imgs]).astype(np.float32)
out = dict(images_bits=bits, descriptors=desc)
for name in ("svm_032015_20_20_same", "svm_032015_20_20"):
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, name)
        with open(path, "wb") as f:
            f.write(lzma.open(os.path.join(G, name + ".xz")).read())
        assert open(path, "rb").read() == open("/root/reference/" + name, "rb").read()
        svm = cv2.ml.SVM_load(path)
    raw = np.array([svm.predict(d.reshape(1, -1), flags=cv2.ml.STAT_MODEL_RAW_OUTPUT)[1][0, 0] for d in desc], np.float32)
    lab = np.array([svm.predict(d.reshape(1, -1))[1][0, 0] for d in desc], np.float32)
    out["raw_" + name], out["label_" + name] = raw, lab
    print(name, "sv", svm.getSupportVectors().shape, "positives", int((lab == 1).sum()), "of", len(lab))
np.savez_compressed(os.path.join(G, "poly_svm_cv2.npz"), **out)
print(os.path.getsize(os.path.join(G, "poly_svm_cv2.npz")))
