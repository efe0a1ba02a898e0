"""Generates the committed fixtures under tests/golden/ (run in the build container).

 * hog_svm_cv2.npz   — produced by cv2 4.13 itself (independent of the oracle): 12 binary grasp images,
                       their cv2.HOGDescriptor descriptors (configured as learning.cpp:194-195,220) and
                       the cv2.ml.SVM raw decision values / labels of the reference's shipped linear model.
 * dggev_scipy.npz   — produced by scipy.linalg.lapack.dggev: Taubin (M, N) pencils of real neighbourhoods
                       with their generalized eigenvalues and the eigenvector the reference would select.
 * pipeline_small.npz— produced by the oracle on a seeded 200x150 scene: voxelised cloud, sample indices,
                       frames (reference dggev path and extended-precision check), grasp records, packed grasp
                       images and SVM scores.  Regression fixture + GPU parity target that needs no oracle.
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cv2
from scipy.linalg import lapack
from agile_grasp_b200 import scenes, api
from oracle import oracle as O

G = os.path.join(ROOT, "tests", "golden")
SVM = "/root/reference/svm_032015_linear_20_20_same"
rng = np.random.default_rng(20150320)

# ---- cv2 fixtures
hog = cv2.HOGDescriptor((64, 64), (16, 16), (8, 8), (8, 8), 9, 1, -1.0, 0, 0.2, True, 64, False)
svm = cv2.ml.SVM_load(SVM)
imgs = []
for t in range(12):
    img = np.zeros((80, 100), np.uint8)
    if t % 3 == 0:
        img[rng.random((80, 100)) < rng.uniform(0.02, 0.4)] = 255
    elif t % 3 == 1:
        for k in range(rng.integers(1, 5)):
            cy, cx, r = rng.integers(0, 80), rng.integers(0, 100), rng.integers(3, 25)
            yy, xx = np.ogrid[:80, :100]
            img[(yy - cy) ** 2 + (xx - cx) ** 2 < r * r] = 255
    else:
        img[rng.integers(0, 80, 300), rng.integers(0, 100, 300)] = 255
    imgs.append(img)
imgs = np.stack(imgs)
desc = np.stack([hog.compute(im, (32, 32), (0, 0)).ravel() for im in imgs]).astype(np.float32)
raw = np.array([svm.predict(d.reshape(1, -1), flags=cv2.ml.STAT_MODEL_RAW_OUTPUT)[1][0, 0] for d in desc], np.float32)
lab = np.array([svm.predict(d.reshape(1, -1))[1][0, 0] for d in desc], np.float32)
np.savez_compressed(os.path.join(G, "hog_svm_cv2.npz"), images_bits=api.pack_images(imgs), descriptors=desc,
                    svm_raw=raw, svm_label=lab)

# ---- pipeline fixture (oracle) + dggev fixture (scipy)
pts, size_left, P, S = scenes.config_cloud(2, small=(200, 150, 60))
xyz, cam = O.preprocess(pts, size_left, P, False)
tree = O.Tree(xyz)
idx = O.draw_samples(len(xyz), S, P.seed)
r = O.fit_quadrics(tree, cam, idx, 0.03, P, want_params=True, want_mn=True)
frames = r["frames"]
exact = O.fit_quadrics(tree, cam, idx, 0.03, P, sum_perm=-1)["frames"]
Ms, Ns, lams, vecs = [], [], [], []
for k in range(0, len(idx), 10):
    M, N = r["MN"][k, 0].copy(), r["MN"][k, 1].copy()
    alphar, alphai, beta, vl, vr, work, info = lapack.dggev(M, N, compute_vl=0, compute_vr=1)
    with np.errstate(all="ignore"):
        lam = alphar / beta
    mi = int(np.argmin(np.where(np.isnan(lam[:9]), np.inf, lam[:9])))
    Ms.append(r["MN"][k, 0]); Ns.append(r["MN"][k, 1]); lams.append(lam); vecs.append(vr[:, mi])
np.savez_compressed(os.path.join(G, "dggev_scipy.npz"), M=np.stack(Ms), N=np.stack(Ns), lam=np.stack(lams),
                    vec=np.stack(vecs), sample_pos=np.arange(0, len(idx), 10))
normals = np.zeros((len(xyz), 3)); normals[idx] = frames["normal"]
H = O.find_hands(tree, cam, idx, frames, cam[idx], normals, P)
osvm = O.Svm(SVM)
keep = H.classify(osvm, P)
g = H.grasps
images = api.pack_images(np.stack([H.image(k, P) for k in range(len(g))]))
np.savez_compressed(os.path.join(G, "pipeline_small.npz"), points=pts[:, :3].astype(np.float32), size_left=size_left,
                    xyz=xyz, cam=cam, idx=idx, frames=frames, frames_exact=exact, grasps=g, images_bits=images,
                    keep=keep, lapack=str(O.lib()._lapack))
print("golden written:", {f: os.path.getsize(os.path.join(G, f)) for f in os.listdir(G)})
print("hyps", len(g), "kept", keep.sum())
