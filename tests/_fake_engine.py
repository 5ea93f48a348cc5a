"""Test double for pymf_b200.engine.Engine backed by the CPU oracle (float64).

Used ONLY by the CPU tests of the host logic (attribute semantics, flag handling, early-stop
bookkeeping, column sharding over torch.distributed/gloo).  It is test infrastructure: the
product never imports it and has no CPU path.
"""
import numpy as np

from oracle import nmf_oracle as O


class FakeEngine(object):
    def __init__(self, d, n_local, k, device=0, n_global=None, col0=0, path=None):
        self.d, self.n_local, self.k = d, n_local, k
        self.n_global = n_local if n_global is None else n_global
        self.col0 = col0
        self.world, self.rank = 1, 0
        self.X = self.W = self.H = None
        self.uploads = {"x": 0, "w": 0, "h": 0}
        self.lam = {"W": 0.0, "H": 0.0}
        self.inc = (1.0, 1.0)
        self.variant = "nmf"

    # comm -------------------------------------------------------------------------------
    @staticmethod
    def comm_unique_id():
        return b"\0" * 128

    def comm_init(self, uid, world, rank):
        self.world, self.rank = world, rank

    def _allreduce(self, a):
        if self.world == 1:
            return a
        import torch
        import torch.distributed as dist
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64).copy())
        dist.all_reduce(t)
        return t.numpy()

    def set_penalty(self, lamb_w, lamb_h, increase_w=1.0, increase_h=1.0):
        self.lam = {"W": float(lamb_w), "H": float(lamb_h)}
        self.inc = (float(increase_w), float(increase_h))

    def set_variant(self, variant):
        self.variant = variant

    def get_penalty(self):
        return self.lam["W"], self.lam["H"]

    # data -------------------------------------------------------------------------------
    def upload_x(self, x):
        self.X = np.array(x, dtype=np.float64)
        self.uploads["x"] += 1

    def upload_x_panels(self, source, panel_bytes=64 << 20):
        from pymf_b200.engine import panel_ranges
        dt = np.dtype(getattr(source, "dtype", np.float64))
        pw, ranges = panel_ranges(self.d, self.n_local, 4 if dt == np.float32 else 8, panel_bytes)
        self.X = np.empty((self.d, self.n_local))
        for c0, w in ranges:
            self.X[:, c0:c0 + w] = source[:, c0:c0 + w]
        self.uploads["x"] += 1
        return pw

    def set_w(self, w):
        self.W = np.array(w, dtype=np.float64)
        self.uploads["w"] += 1

    def set_h(self, h):
        self.H = np.array(h, dtype=np.float64)
        self.uploads["h"] += 1

    def get_w(self, dtype=np.float64, out=None):
        return self.W.astype(dtype)

    def get_h(self, dtype=np.float64, out=None):
        return self.H.astype(dtype)

    # loop -------------------------------------------------------------------------------
    def _err(self):
        if self.world == 1:
            return O.frobenius_norm(self.X, self.W, self.H)
        loc = np.sum((self.X - self.W.dot(self.H)) ** 2)
        return float(np.sqrt(self._allreduce(np.array([loc]))[0]))

    def run(self, niter, compute_w=True, compute_h=True, compute_err=True, early_stop=True):
        ferr = np.zeros(niter)
        done = 0
        nf = niter if compute_err else 0
        for i in range(niter):
            if compute_w and self.world > 1:
                # sharded columns: only the d x k / k x k partials are summed over ranks
                A = self._allreduce(self.X.dot(self.H.T))
                B = self._allreduce(self.H.dot(self.H.T))
                if self.variant == "snmf":
                    self.W = A.dot(np.linalg.inv(B))
                elif self.lam["W"] != 0.0 or self.lam["H"] != 0.0:
                    lw = self.lam["W"]
                    W1 = A + 3.0 * lw * (self.W ** 2)
                    W2 = self.W.dot(B) + 2.0 * lw * (self.W ** 3) + lw * self.W + O.EPS_DENOM
                    self.W *= W1 / W2
                else:
                    W2 = self.W.dot(B) + O.EPS_DENOM
                    self.W *= A
                    self.W /= W2
            elif compute_w and self.variant == "snmf":
                self.W = O.snmf_update_w(self.X, self.W, self.H)
            elif compute_w:
                if self.lam["W"] != 0.0 or self.lam["H"] != 0.0:
                    O.bnmf_update_w(self.X, self.W, self.H, self.lam)
                else:
                    O.update_w(self.X, self.W, self.H)
            if compute_h and self.variant == "snmf":
                O.snmf_update_h(self.X, self.W, self.H)
            elif compute_h:
                if self.lam["W"] != 0.0 or self.lam["H"] != 0.0:
                    assert self.inc == (O.LAMB_INCREASE_W, O.LAMB_INCREASE_H)
                    O.bnmf_update_h(self.X, self.W, self.H, self.lam)
                else:
                    O.update_h(self.X, self.W, self.H)
            if compute_err:
                ferr[i] = self._err()
            done = i + 1
            if early_stop and compute_err and i > 1:
                if abs(ferr[i] - ferr[i - 1]) / self.n_global < O.EPS_CONV:
                    nf = i
                    break
        return ferr[:nf].copy(), done

    def frobenius(self):
        return self._err()
