import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def linear_svm_path():
    return os.path.join(ROOT, "tests", "golden", "svm_032015_linear_20_20_same")


@pytest.fixture(scope="session")
def small_scene(oracle):
    """320x240 single-camera scene, 150 samples: voxelised cloud + oracle tree, shared by many tests."""
    from agile_grasp_b200 import scenes
    pts, size_left, P, S = scenes.config_cloud(2, small=(320, 240, 150))
    xyz, cam = oracle.preprocess(pts, size_left, P, False)
    tree = oracle.Tree(xyz)
    idx = oracle.draw_samples(len(xyz), S, P.seed)
    return dict(pts=pts, size_left=size_left, P=P, xyz=xyz, cam=cam, tree=tree, idx=idx)


@pytest.fixture(scope="session")
def two_view_scene(oracle):
    """two registered 200x150 views (camera labels, NaN-shift quirk), 120 samples"""
    from agile_grasp_b200 import scenes
    pts, size_left, P, S = scenes.config_cloud(3, small=(200, 150, 120))
    xyz, cam = oracle.preprocess(pts, size_left, P, False)
    tree = oracle.Tree(xyz)
    idx = oracle.draw_samples(len(xyz), S, P.seed)
    return dict(pts=pts, size_left=size_left, P=P, xyz=xyz, cam=cam, tree=tree, idx=idx)


@pytest.fixture(scope="session")
def poly_svm_paths(tmp_path_factory):
    """the reference's two shipped POLY-kernel model files (committed xz-compressed: 11 / 25 MB of YAML text)"""
    import lzma
    d = tmp_path_factory.mktemp("poly_svm")
    out = {}
    for name in ("svm_032015_20_20_same", "svm_032015_20_20"):
        p = d / name
        p.write_bytes(lzma.open(os.path.join(ROOT, "tests", "golden", name + ".xz")).read())
        out[name] = str(p)
    return out
