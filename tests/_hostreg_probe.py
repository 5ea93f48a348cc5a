"""GPU-box experiment (not a test): how long does cudaHostRegister of an existing pageable numpy array take, compared
with the staged copy of the pageable ingest path?"""
import time
import numpy as np
import torch

rt = torch.cuda.cudart()
torch.cuda.init()
for gib in (2, 8, 16):
    a = np.ones((gib << 28,), dtype=np.float32)          # touched pageable memory
    t0 = time.perf_counter()
    rc = rt.cudaHostRegister(a.ctypes.data, a.nbytes, 0)
    t1 = time.perf_counter()
    d = torch.empty(a.nbytes // 4, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    d.copy_(torch.from_numpy(a), non_blocking=True)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    rc2 = rt.cudaHostUnregister(a.ctypes.data)
    t4 = time.perf_counter()
    print("%2d GiB: register %.3f s (%.1f GB/s) rc=%s, H2D of the registered array %.3f s (%.1f GB/s), unregister %.3f s" % (
        gib, t1 - t0, a.nbytes / (t1 - t0) / 1e9, rc, t3 - t2, a.nbytes / (t3 - t2) / 1e9, t4 - t3), flush=True)
    del d, a
