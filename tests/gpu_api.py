"""
Test helper: the reference's free-function signatures (X (V,D), W (V,k), H (k,D) numpy arrays)
implemented by calls through the C ABI, so GPU parity tests read like reference
tests/test_utils_klnmf.py.  Nothing here computes on the CPU.
"""

from __future__ import annotations

import numpy as np
import torch

from salamander_b200._device import (
    PASS_OBJECTIVE,
    PASS_POISSON,
    PASS_SAMPLEWISE,
    PASS_UPDATE_H,
    PASS_WNUM,
    Workspace,
)


class GpuKL:
    def __init__(self, dtype=torch.float64, math="fma", device="cuda:0"):
        self.dtype, self.math, self.dev = dtype, math, torch.device(device)

    def _up(self, a, transpose=True):
        if a is None:
            return None
        a = np.asarray(a, dtype=np.float64)
        a = a.T if transpose else a
        return torch.as_tensor(np.ascontiguousarray(a), dtype=self.dtype).to(self.dev).contiguous()

    def _setup(self, X, W, H):
        V, D = X.shape
        k = W.shape[1]
        ws = Workspace(V, D, k, self.dtype, self.dev, math=self.math)
        return ws, self._up(X), self._up(W), self._up(H)

    def kl_divergence(self, X, W, H, weights=None, weights_lhalf=None):
        ws, Xd, Wd, Hd = self._setup(X, W, H)
        obj = torch.zeros(1, dtype=torch.float64, device=self.dev)
        ws.klnmf_pass(Xd, Wd, Hd, PASS_OBJECTIVE, w_kl=self._up(weights, False), w_lhalf=self._up(weights_lhalf, False), objective=obj)
        out = float(obj.item())
        ws.close()
        return out

    def samplewise_kl_divergence(self, X, W, H, weights=None):
        ws, Xd, Wd, Hd = self._setup(X, W, H)
        per = torch.zeros(X.shape[1], dtype=self.dtype, device=self.dev)
        ws.klnmf_pass(Xd, Wd, Hd, PASS_SAMPLEWISE, per_sample=per)
        out = per.double().cpu().numpy()
        ws.close()
        return out if weights is None else out * np.asarray(weights)

    def poisson_llh_wo_factorial(self, X, W, H):
        ws, Xd, Wd, Hd = self._setup(X, W, H)
        obj = torch.zeros(1, dtype=torch.float64, device=self.dev)
        ws.klnmf_pass(Xd, Wd, Hd, PASS_POISSON, objective=obj)
        out = float(obj.item())
        ws.close()
        return out

    def update_W(self, X, W, H, weights_kl=None, n_given_signatures=0):
        ws, Xd, Wd, Hd = self._setup(X, W, H)
        Wnum = torch.zeros_like(Wd)
        if n_given_signatures < W.shape[1]:
            ws.klnmf_pass(Xd, Wd, Hd, PASS_WNUM, w_kl=self._up(weights_kl, False), Wnum=Wnum)
        Wout = torch.empty_like(Wd)
        ws.w_epilogue(Wd, Wnum, n_given_signatures, False, Wout)
        out = Wout.double().cpu().numpy().T
        ws.close()
        return out

    def update_H(self, X, W, H, weights_kl=None, weights_lhalf=None):
        ws, Xd, Wd, Hd = self._setup(X, W, H)
        ws.klnmf_pass(Xd, Wd, Hd, PASS_UPDATE_H, H_out=Hd, w_kl=self._up(weights_kl, False), w_lhalf=self._up(weights_lhalf, False))
        out = Hd.double().cpu().numpy().T
        ws.close()
        return out

    def update_WH(self, X, W, H, weights_kl=None, weights_lhalf=None, n_given_signatures=0):
        ws, Xd, Wd, Hd = self._setup(X, W, H)
        Wnum = torch.zeros_like(Wd)
        flags = PASS_UPDATE_H | (PASS_WNUM if n_given_signatures < W.shape[1] else 0)
        ws.klnmf_pass(Xd, Wd, Hd, flags, H_out=Hd, w_kl=self._up(weights_kl, False), w_lhalf=self._up(weights_lhalf, False), Wnum=Wnum)
        ws.w_epilogue(Wd, Wnum, n_given_signatures, True, Wd)
        outW, outH = Wd.double().cpu().numpy().T, Hd.double().cpu().numpy().T
        ws.close()
        return outW, outH
