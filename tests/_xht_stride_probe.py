"""GPU experiment (not a test): does the X.H^T pass depend on the ROW STRIDE of X (TLB reach of its 128-row TMA boxes)?
Same bytes (16 GiB), same k, three aspect ratios."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pymf_b200  # noqa: E402

for d, n, k in ((8192, 524288, 64), (32768, 131072, 64), (65536, 65536, 64), (4096, 262144, 32), (16384, 65536, 32)):
    e = pymf_b200.Engine(d, n, k, path="tc")
    e.gen_x(1); e.gen_w(2); e.gen_h(3)
    e.enqueue(3); e.sync()
    e.kernel_timing(True)
    e.enqueue(6); e.sync()
    th, _ = e.kernel_timing_read(0)
    tx, _ = e.kernel_timing_read(1)
    gb = (4.0 * d * n + 8.0 * k * n) / 1e9
    print("d=%6d n=%7d k=%3d row stride %7.0f KB : H pass %.3f ms (%.2f TB/s)   X.H^T pass %.3f ms (%.2f TB/s)" % (
        d, n, k, 4.0 * n / 1024, th, gb / th, tx, gb / tx), flush=True)
    e.close()
