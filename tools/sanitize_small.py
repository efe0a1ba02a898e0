"""A small pass over every kernel of the hot path for compute-sanitizer (memcheck / racecheck / initcheck):
   compute-sanitizer --tool memcheck python tools/sanitize_small.py
production normal mode, fused linear scoring, the POLY model through ag_classify, a sample-sharded call (k_ball_over50),
calculates_antipodal (all-points pass + carry of the rand() stream), the training features and handle search."""
import lzma, os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from agile_grasp_b200 import api, scenes

pts, size_left, P, S = scenes.config_cloud(3, small=(200, 150, 120))
P.deterministic_normals = 0
lin = api.Svm(os.path.join(ROOT, "tests", "golden", "svm_032015_linear_20_20_same"))
d = tempfile.mkdtemp(prefix="ag_poly_")
pp = os.path.join(d, "svm_032015_20_20_same")
open(pp, "wb").write(lzma.open(os.path.join(ROOT, "tests", "golden", "svm_032015_20_20_same.xz")).read())
poly = api.Svm(pp)
for stage in (False, True):
    ctx = api.Context(0, P, stage_timing=stage)
    ctx.set_svm(lin)
    for _ in range(3):  # eager, capture, replay
        g = ctx.localize(pts, size_left)
    gl, keep = ctx.classify(lin, g)
    ctx.set_svm(None)
    g2 = ctx.localize(pts, size_left)
    gp, keep_p = ctx.classify(poly, g2)
    ga = ctx.localize(pts, size_left, flags=1)  # calculates_antipodal
    if len(g2):
        ctx.find_handles(g2, 3, 0.005)
    print("stage events", stage, "hyp", len(g), "linear positives", int(keep.sum()), "poly positives", int(keep_p.sum()),
          "with all-points normals", len(ga))
    ctx.close()
P.shard_index, P.shard_count, P.shard_interleave = 1, 2, 1
ctx = api.Context(0, P)
ctx.set_svm(lin)
for _ in range(3):
    gs = ctx.localize(pts, size_left)
print("shard 1 of 2:", len(gs), "hypotheses")
ctx.set_svm(None)
ctx.close()
print("done")
