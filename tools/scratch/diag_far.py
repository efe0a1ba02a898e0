import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from agile_grasp_b200 import api, scenes
from oracle import oracle as O
pts, size_left, P, S = scenes.config_cloud(2, small=(320, 240, 150))
far = pts.copy()
far[0, :3] = [-9.0, -9.0, -9.0]
far[size_left - 1, :3] = [9.0, 9.0, 9.0]
xo, co = O.preprocess(far, size_left, P, False)
idx_f = O.draw_samples(len(xo), 50, P.seed)
print("oracle voxels", len(xo), "idx max", idx_f.max())
for trial in range(3):
    c2 = api.Context(0, P)
    try:
        g = c2.localize(far, size_left, idx_f)
        print("trial", trial, "ok", len(g), c2.timings()["n_voxels"])
    except Exception as e:
        print("trial", trial, "ERR", e, c2.timings()["n_voxels"], c2.timings()["n_samples"])
        try:
            g = c2.localize(far, size_left, idx_f)
            print("   second call ok", len(g), c2.timings()["n_voxels"])
        except Exception as e2:
            print("   second call ERR", e2)
    c2.close()
c3 = api.Context(0, P)
g = c3.localize(far, size_left)   # drawn samples
print("drawn ok", len(g), c3.timings()["n_voxels"], c3.timings()["n_samples"])
