"""The drop-in boundary: libag_b200.so loads without a GPU and exports every symbol the header declares."""
import ctypes
import os
import re

from agile_grasp_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "ag_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ag_[a-z_0-9]+)\s*\(", txt)))


def test_header_symbols_are_exported():
    L = ctypes.CDLL(api.LIB_PATH)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), n
    assert set(names) == set(api.EXPORTS)


def test_no_gpu_means_loud_failure_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        return
    L = api.lib()
    h = L.ag_create(0)
    assert not h
    msg = L.ag_last_error().decode()
    assert "no CUDA device" in msg and "no CPU fallback" in msg
    try:
        api.Context(0)
        assert False
    except api.AgError:
        pass


def test_struct_layouts_match_header():
    from agile_grasp_b200.ctypes_defs import AgFrame, AgGrasp, AgParams
    assert ctypes.sizeof(AgGrasp) == 160 and ctypes.sizeof(AgFrame) == 80
    assert ctypes.sizeof(AgParams) == 5 * 8 + 6 * 8 + 32 * 8 + 4 * 8 + 6 * 4 + 8 + 2 * 4  # + shard_index, shard_count
    p = AgParams()
    api.lib().ag_default_params(ctypes.byref(p))
    assert (p.finger_width, p.hand_outer_diameter, p.hand_depth, p.hand_height, p.init_bite) == (0.01, 0.09, 0.06, 0.02, 0.01)
    assert (p.nn_radius_taubin, p.nn_radius_hands, p.nn_radius_normals, p.voxel_size) == (0.03, 0.08, 0.01, 0.003)
    assert p.num_samples == 2000 and p.deterministic_normals == 1


def test_svm_loader_without_gpu(linear_svm_path, tmp_path):
    s = api.Svm(linear_svm_path)
    assert (s.kernel, s.var_count, s.sv_total) == (0, 3528, 1) and s.rho == -3.1383255947302025e-01
    try:
        api.Svm(tmp_path / "missing")
        assert False
    except api.AgError as e:
        assert "does not exist" in str(e)  # learning.cpp:172-178


def test_cpp_shim_compiles_and_links(tmp_path):
    """include/agile_grasp/localization.h (the reference's class API) builds against the C ABI with a
    plain host compiler, no CUDA/Eigen/PCL headers needed."""
    import subprocess
    exe = tmp_path / "test_svm"
    libdir = os.path.dirname(api.LIB_PATH)
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "test_svm.cpp"), "-o", str(exe), "-L" + libdir, "-lag_b200",
           "-Wl,-rpath," + libdir, "-L/usr/local/cuda/lib64", "-lcudart"]
    subprocess.check_call(cmd)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert "Usage: test_svm" in out.stdout
