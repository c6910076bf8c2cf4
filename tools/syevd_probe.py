#!/usr/bin/env python
"""How long does cusolverDnXsyevd take for the K x K core of the low-rank path (K = 200, FP64, vectors)?  Decides whether an
eigen-form core (diagonal Bc, symmetric positive definite K x K system per M-step) is worth its one-off cost."""
import ctypes, sys, time
import torch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
lib = ctypes.CDLL("libcusolver.so.11")
h = ctypes.c_void_p(); p = ctypes.c_void_p()
assert lib.cusolverDnCreate(ctypes.byref(h)) == 0
assert lib.cusolverDnCreateParams(ctypes.byref(p)) == 0
stream = torch.cuda.current_stream().cuda_stream
assert lib.cusolverDnSetStream(h, ctypes.c_void_p(stream)) == 0
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.rand(3000, 3, dtype=torch.float64, device="cuda", generator=g)
gram = torch.exp(-torch.cdist(x, x) ** 2 / 4.0)
q, _ = torch.linalg.qr(torch.rand(3000, n, dtype=torch.float64, device="cuda", generator=g))
bc = q.T @ gram @ q
bc = 0.5 * (bc + bc.T)
w = torch.empty(n, dtype=torch.float64, device="cuda")
wd, wh = ctypes.c_size_t(), ctypes.c_size_t()
R64 = 1
args = (h, p, 1, 0, ctypes.c_int64(n), R64)
a = bc.clone()
st = lib.cusolverDnXsyevd_bufferSize(h, p, 1, 0, ctypes.c_int64(n), R64, ctypes.c_void_p(a.data_ptr()), ctypes.c_int64(n), R64,
                                     ctypes.c_void_p(w.data_ptr()), R64, ctypes.byref(wd), ctypes.byref(wh))
assert st == 0, st
dbuf = torch.empty(max(wd.value, 8), dtype=torch.uint8, device="cuda")
hbuf = ctypes.create_string_buffer(max(wh.value, 8))
info = torch.zeros(1, dtype=torch.int32, device="cuda")
print("workspace: device %d B, host %d B" % (wd.value, wh.value))
for rep in range(6):
    a = bc.clone()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    st = lib.cusolverDnXsyevd(h, p, 1, 0, ctypes.c_int64(n), R64, ctypes.c_void_p(a.data_ptr()), ctypes.c_int64(n), R64,
                              ctypes.c_void_p(w.data_ptr()), R64, ctypes.c_void_p(dbuf.data_ptr()), ctypes.c_size_t(wd.value),
                              hbuf, ctypes.c_size_t(wh.value), ctypes.c_void_p(info.data_ptr()))
    torch.cuda.synchronize(); t1 = time.perf_counter()
    assert st == 0, st
    print("syevd n=%d: %.3f ms (info %d)" % (n, (t1 - t0) * 1e3, int(info.item())))
rec = (a.T * w) @ a        # row-major view: row j of `a` is eigenvector j
print("eigenvalues %.3e .. %.3e; |V diag(w) V^T - Bc| max %.2e" % (float(w.min()), float(w.max()), float((rec - bc).abs().max())))
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ww, vv = torch.linalg.eigh(bc)
    torch.cuda.synchronize(); print("torch.linalg.eigh: %.3f ms" % ((time.perf_counter() - t0) * 1e3))
