import sys, numpy as np
sys.path.insert(0, "/root/repo")
import pymf_b200
for (d, n, k) in [(4096, 2048, 32), (4096, 4096, 32), (4096, 8192, 32), (4096, 16384, 32), (4096, 65536, 32)]:
    e = pymf_b200.Engine(d, n, k, path="tc")
    e.set_err_mode("trace")
    e.gen_x(1); e.gen_w(2); e.gen_h(3)
    e.enqueue(5); e.sync()
    e.kernel_timing(True)
    e.enqueue(30); e.sync()
    th, nh = e.kernel_timing_read(0); tx, nx = e.kernel_timing_read(1)
    xb = 4.0 * d * n
    print("d=%d n=%d X=%.0f MB: h_update %.1f us (%.2f TB/s of X) | xht %.1f us (%.2f TB/s of X)" % (d, n, xb / 1e6, th * 1e3, xb / th / 1e9, tx * 1e3, xb / tx / 1e9), flush=True)
    e.close()
