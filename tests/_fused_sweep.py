"""GPU experiment (not a test): sweep the one-pass fused kernel's schedule parameters on cfg2.

  python tests/_fused_sweep.py [d n k] > gpurun_out/fused_sweep.txt

Prints ms/iteration and the fused-kernel time for each (sbc, slab, lag, hint); the first line is the
two-pass baseline.  Parameters are read by fused_plan() from the environment at plan time.
"""
import itertools
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pymf_b200  # noqa: E402


def run(d, n, k, env, steps=10):
    for key in list(os.environ):
        if key.startswith("PYMFB_FUSED"):
            del os.environ[key]
    os.environ.update(env)
    eng = pymf_b200.Engine(d, n, k, device=0)
    eng.gen_x(1234); eng.gen_w(1235); eng.gen_h(1236)
    eng.enqueue(3); eng.sync()
    eng.kernel_timing(True)
    e0, e1 = eng.event(), eng.event()
    eng.record(e0); eng.enqueue(steps); eng.record(e1); eng.sync()
    ms = eng.elapsed_ms(e0, e1) / steps
    t0, n0 = eng.kernel_timing_read(0)
    t1, n1 = eng.kernel_timing_read(1)
    f = eng.frobenius()
    eng.close()
    return ms, t0, t1, f


def main():
    d, n, k = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (4096, 262144, 32)
    ms, t0, t1, f = run(d, n, k, {})
    print("two-pass            ms/iter %.3f  h %.3f  xht %.3f  ferr %.4f" % (ms, t0, t1, f), flush=True)
    grid = os.environ.get("SWEEP", "full")
    if grid == "none":
        return
    if grid == "full":
        combos = [(sbc, slab, bcols, lag, 1)
                  for sbc, slab, bcols in ((1024, 256, 256), (1024, 256, 128), (1024, 512, 256), (512, 256, 256),
                                           (512, 256, 128), (2048, 256, 256), (2048, 512, 512), (512, 512, 128))
                  for lag in (1, 2, 3, 4)]
    else:
        combos = [tuple(int(x) for x in c.split(",")) for c in grid.split(";")]
    for sbc, slab, bcols, lag, hint in combos:
        env = {"PYMFB_FUSED": "1", "PYMFB_FUSED_SBC": str(sbc), "PYMFB_FUSED_SLAB": str(slab),
               "PYMFB_FUSED_BCOLS": str(bcols), "PYMFB_FUSED_LAG": str(lag), "PYMFB_FUSED_HINT": str(hint)}
        try:
            ms, t0, t1, f = run(d, n, k, env)
            print("sbc %4d slab %3d bcols %4d lag %2d hint %d  window %5.1f MB  ms/iter %.3f  fused %.3f  ferr %.4f"
                  % (sbc, slab, bcols, lag, hint, (lag + 1) * d * sbc * 4 / 1e6, ms, t0, f), flush=True)
        except Exception as e:  # noqa: BLE001
            print("sbc %d slab %d bcols %d lag %d hint %d FAILED: %s" % (sbc, slab, bcols, lag, hint, e), flush=True)
            break


if __name__ == "__main__":
    main()
