"""Parity statistics at the full BASELINE sizes (configs 1-3): the CUDA path (ag_localize + ag_classify on its own
frames) against the CPU oracle's full path on the same clouds and the same sample indices.  Stage-level bit-exact
parity (given shared frames) is what tests/test_gpu_parity.py asserts; this report shows what is left end to end,
where the only non-bit-exact stage (the Taubin eigen-solve: exact here, LAPACK dggev noise in the oracle, DESIGN.md
section 2) can flip a borderline decision.  Writes gpurun_out/parity_report.json."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from agile_grasp_b200 import api, scenes
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
svm_path = os.path.join(ROOT, "tests/golden/svm_032015_linear_20_20_same")
rows = []
for cfg in (1, 2, 3):
    pts, size_left, P, S = scenes.config_cloud(cfg)
    P.num_threads = os.cpu_count() or 1
    xo, co = O.preprocess(pts, size_left, P, False)
    idx = O.draw_samples(len(xo), S, P.seed)
    ctx = api.Context(0, P)
    xyz, cam = ctx.preprocess(pts, size_left)
    vox_equal = bool((xyz.view(np.uint32) == xo.view(np.uint32)).all() and (cam == co).all())
    g = ctx.localize(pts, size_left, idx)
    gg, keep = ctx.classify(api.Svm(svm_path), g)
    H, tm, nv = O.localize(pts, size_left, P, idx, 0, O.Svm(svm_path), False)
    go = H.grasps
    # frames: GPU vs the oracle's dggev frames and vs its extended-precision solve
    tree = O.Tree(xo)
    fo = O.fit_quadrics(tree, co, idx, 0.03, P)["frames"]
    fx = O.fit_quadrics(tree, co, idx, 0.03, P, sum_perm=-1)["frames"]
    fg = ctx.fit_quadrics(idx, 0.03)
    det = fo["num_neighbors"] >= 10
    d_ref = np.linalg.norm(fg["normal"] - fo["normal"], axis=1)[det]
    d_ex = np.linalg.norm(fg["normal"] - fx["normal"], axis=1)[det]
    ko = {(a, b): i for i, (a, b) in enumerate(zip(go["sample_index"].tolist(), go["orientation"].tolist()))}
    kg = {(a, b): i for i, (a, b) in enumerate(zip(gg["sample_index"].tolist(), gg["orientation"].tolist()))}
    common = sorted(set(ko) & set(kg))
    io = np.array([ko[k] for k in common], dtype=int)
    ig = np.array([kg[k] for k in common], dtype=int)
    sc_o, sc_g = go["score"][io], gg["score"][ig]
    rel = np.abs(sc_g - sc_o) / np.maximum(1.0, np.abs(sc_o))
    row = dict(config=cfg, points=int(len(pts)), voxels=int(len(xo)), voxel_cloud_bit_identical=vox_equal, samples=int(len(idx)),
               neighbour_counts_equal=bool(np.array_equal(fg["num_neighbors"], fo["num_neighbors"])),
               normals_vs_dggev_oracle=dict(median=float(np.median(d_ref)), p99=float(np.quantile(d_ref, 0.99)), max=float(d_ref.max()),
                                            frac_le_1e5=float((d_ref <= 1e-5).mean())),
               normals_vs_extended_precision_solve=dict(median=float(np.median(d_ex)), max=float(d_ex.max())),
               hypotheses_gpu=int(len(gg)), hypotheses_oracle=int(len(go)), common=int(len(common)),
               flags_equal_on_common=float(((gg["half_antipodal"][ig] == go["half_antipodal"][io]) &
                                            (gg["full_antipodal"][ig] == go["full_antipodal"][io])).mean()) if len(common) else None,
               box_point_counts_equal=float((gg["num_points"][ig] == go["num_points"][io]).mean()) if len(common) else None,
               labels_equal=float((gg["label"][ig] == go["label"][io]).mean()) if len(common) else None,
               score_rel_diff=dict(median=float(np.median(rel)), p99=float(np.quantile(rel, 0.99)), max=float(rel.max()),
                                   frac_le_1e5=float((rel <= 1e-5).mean())) if len(common) else None,
               positives_gpu=int(keep.sum()), positives_oracle=int((go["label"] == 1).sum()))
    rows.append(row)
    print(json.dumps(row), flush=True)
    ctx.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/parity_report.json", "w"), indent=1)
