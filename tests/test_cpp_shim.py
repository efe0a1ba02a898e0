"""The C++ drop-in interface (include/agile_grasp/*.h mirroring the reference's Localization / GraspHypothesis /
Handle / Grasp msgs) compiles with the system compiler against libag_b200.so; its host-only parts run here."""
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CXX = "/usr/bin/g++"


def _compile(src, out):
    env = {k: v for k, v in os.environ.items() if k not in ("CXX", "CC")}
    cmd = [CXX, "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), src, "-o", out,
           "-L", os.path.join(ROOT, "agile_grasp_b200"), "-lag_b200", "-Wl,-rpath," + os.path.join(ROOT, "agile_grasp_b200"),
           "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64", "-lcudart"]
    subprocess.check_call(cmd, env=env)


def test_shim_compiles_and_host_parts_run(tmp_path):
    from agile_grasp_b200 import api
    api.lib()  # builds nothing; fails loudly if the library is missing
    exe = str(tmp_path / "shim_test")
    _compile(os.path.join(ROOT, "tests", "cpp", "shim_test.cpp"), exe)
    n = 77
    rec = np.zeros(n, dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("rgba", "<u4")])
    rec["x"] = np.arange(n)
    pcd = tmp_path / "t.pcd"
    hdr = ("VERSION 0.7\nFIELDS x y z rgba\nSIZE 4 4 4 4\nTYPE F F F U\nCOUNT 1 1 1 1\nWIDTH %d\nHEIGHT 1\n"
           "VIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA binary\n" % (n, n)).encode()
    pcd.write_bytes(hdr + rec.tobytes())
    out = subprocess.run([exe, str(pcd), str(n)], capture_output=True, text=True)
    assert out.returncode == 0 and "shim ok" in out.stdout, out.stdout + out.stderr
    assert "Couldn't read pcd_filename_left file" in out.stdout


def test_example_cli_compiles(tmp_path):
    _compile(os.path.join(ROOT, "examples", "test_svm.cpp"), str(tmp_path / "test_svm"))
    _compile(os.path.join(ROOT, "examples", "train_svm.cpp"), str(tmp_path / "train_svm"))
