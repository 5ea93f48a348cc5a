"""GPU-box experiment (not a test): cost of cudaMalloc / cudaFree of a cfg3-sized buffer, call after call."""
import ctypes
import time

rt = ctypes.CDLL("libcudart.so.12")
rt.cudaMalloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t]
rt.cudaFree.argtypes = [ctypes.c_void_p]
rt.cudaMemset.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t]
rt.cudaFree(None)
for gib in (4, 64):
    for i in range(5):
        p = ctypes.c_void_p()
        t0 = time.perf_counter()
        rc = rt.cudaMalloc(ctypes.byref(p), gib << 30)
        t1 = time.perf_counter()
        rt.cudaMemset(p, 0, gib << 30); rt.cudaDeviceSynchronize()
        t2 = time.perf_counter()
        rt.cudaFree(p)
        t3 = time.perf_counter()
        print("%2d GiB  call %d: cudaMalloc %.3f s (rc %d)  memset %.3f s  cudaFree %.3f s" % (gib, i, t1 - t0, rc, t2 - t1, t3 - t2), flush=True)
