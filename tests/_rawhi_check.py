"""GPU experiment (not a test): one H update on a k = 128 (SS kernels) shape; saves H so that two builds of the library
(PYMFB_LIB=... with / without -DPYMFB_TC_RAW_HI=1) can be compared bit for bit.  usage: _rawhi_check.py out.npy"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pymf_b200  # noqa: E402

d, n, k = 1024, 4096, 128
e = pymf_b200.Engine(d, n, k, path="tc")
e.gen_x(1); e.gen_w(2); e.gen_h(3)
e.run(1, compute_w=False, compute_h=True, compute_err=False, early_stop=False)
H = e.get_h(np.float32)
np.save(sys.argv[1], H)
f, _ = e.run(3, early_stop=False)
print("ferr", f)
e.close()
