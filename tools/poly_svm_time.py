"""Times the batched POLY scoring at the size of the reference's launch-file model (1190 SVs x 3528, degree 2)
on H hypothesis images: k_hog_svm (descriptors) + k_svm_gemm + k_svm_decide, CUDA events via ag_timings is not
available for this entry point, so wall clock around ag_hog_svm with images resident on the host (small)."""
import os, sys, time, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from agile_grasp_b200 import api, scenes
from test_oracle_hog_svm import write_opencv_svm
rng = np.random.default_rng(3)
nsv, H = int(sys.argv[1]) if len(sys.argv) > 1 else 1190, int(sys.argv[2]) if len(sys.argv) > 2 else 1000
sv = (rng.random((nsv, 3528)) * (rng.random((nsv, 3528)) < 0.2)).astype(np.float32)
d = tempfile.mkdtemp()
path = os.path.join(d, "poly")
write_opencv_svm(path, sv, rng.normal(size=nsv), rho=0.3, kernel="POLY", degree=2, gamma=1.0, coef0=0.0)
svm = api.Svm(path)
pts, size_left, P, S = scenes.config_cloud(2, small=(200, 150, 60))
ctx = api.Context(0, P)
imgs = np.zeros((H, 80, 100), np.uint8)
for t in range(H):
    imgs[t][rng.random((80, 100)) < 0.2] = 255
bits = api.pack_images(imgs)
for rep in range(5):
    t0 = time.perf_counter(); sc, _ = ctx.hog_svm(svm, bits); dt = time.perf_counter() - t0
    print(f"rep {rep}: {H} hypotheses x {nsv} SVs: {dt*1e3:.3f} ms wall ({H*nsv*3528/dt/1e12:.2f} T MAC/s)")
