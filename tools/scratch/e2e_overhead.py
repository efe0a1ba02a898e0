import os, sys, time, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from agile_grasp_b200 import api, scenes
from agile_grasp_b200.ctypes_defs import AgGrasp
pts, size_left, P, S = scenes.config_cloud(2)
P.deterministic_normals = 0
pin = torch.from_numpy(np.ascontiguousarray(pts)).pin_memory()
host = pin.numpy()
ctx = api.Context(0, P)
svm = api.Svm("tests/golden/svm_032015_linear_20_20_same")
ctx.set_svm(svm)
for _ in range(5):
    g = ctx.localize(host, size_left); ctx.classify(svm, g)
L = api.lib()
def t(f, n=30):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    return 1e3 * float(np.median(ts))
print("python localize+classify", t(lambda: ctx.classify(svm, ctx.localize(host, size_left))))
print("python localize only    ", t(lambda: ctx.localize(host, size_left)))
out = C.POINTER(AgGrasp)(); n = C.c_int()
ptr = host.ctypes.data_as(C.c_void_p)
def raw():
    L.ag_localize(ctx.h, ptr, 32, host.shape[0], size_left, None, 0, 0, C.byref(out), C.byref(n)); L.ag_free(out)
print("raw ag_localize         ", t(raw))
dev = pin.cuda()
def rawdev():
    L.ag_localize_device(ctx.h, C.c_void_p(dev.data_ptr()), 32, host.shape[0], size_left, None, 0, 0, C.byref(out), C.byref(n)); L.ag_free(out)
for _ in range(3): rawdev()
print("raw ag_localize_device  ", t(rawdev), "timings total_ms", ctx.timings()["total_ms"])
d2 = torch.empty_like(dev)
def h2d():
    d2.copy_(pin, non_blocking=True); torch.cuda.synchronize()
print("H2D 9.8 MB pinned (torch)", t(h2d))
