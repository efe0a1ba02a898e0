"""One profiled step of the bench workload between cudaProfilerStart/Stop (for ncu --profile-from-start off)."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from agile_grasp_b200 import api, scenes
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 3
pts, size_left, P, S = scenes.config_cloud(cfg)
P.deterministic_normals = int(sys.argv[3]) if len(sys.argv) > 3 else 0  # 0 = the reference's production mode
ctx = api.Context(0, P)
svm = api.Svm(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/golden/svm_032015_linear_20_20_same"))
rt = ctypes.CDLL("libcudart.so")
ctx.set_svm(svm)
for _ in range(warm):
    g = ctx.localize(pts, size_left); ctx.classify(svm, g)
rt.cudaProfilerStart()
g = ctx.localize(pts, size_left); gg, keep = ctx.classify(svm, g)
rt.cudaProfilerStop()
print("profiled step:", len(g), "hypotheses", ctx.timings())
