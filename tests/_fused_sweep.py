"""GPU experiment (not a test): sweep the one-pass fused kernel's schedule parameters on cfg2.

  python tests/_fused_sweep.py [d n k] > gpurun_out/fused_sweep.txt

Prints ms/iteration and the fused-kernel time for each (depth, hint); the first line is the two-pass
baseline.  Parameters are read by fused_plan() from the environment at plan time.  SWEEP="d,h;d,h"
selects combinations, SWEEP=none prints only the baseline.
"""
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
    combos = ([(2, 1), (1, 1), (3, 1), (4, 1), (2, 0), (3, 0)] if grid == "full"
              else [tuple(int(x) for x in c.split(",")) for c in grid.split(";")])
    for combo in combos:
        depth, hint = combo[0], combo[1]
        pf = combo[2] if len(combo) > 2 else 1
        env = {"PYMFB_FUSED": "1", "PYMFB_FUSED_DEPTH": str(depth), "PYMFB_FUSED_HINT": str(hint), "PYMFB_FUSED_PF": str(pf)}
        try:
            ms, t0, t1, f = run(d, n, k, env)
            print("fused depth %d hint %d pf %d  ms/iter %.3f  fused(+gh) %.3f  ferr %.4f" % (depth, hint, pf, ms, t0, f), flush=True)
        except Exception as e:  # noqa: BLE001
            print("fused depth %d hint %d FAILED: %s" % (depth, hint, e), flush=True)
            break


if __name__ == "__main__":
    main()
