"""On-device L-BFGS (row f-1 of SURVEY.md section 8).

`DeviceLBFGS` follows torch.optim.LBFGS.step (torch/optim/lbfgs.py — the optimiser the reference constructs at
attack_PCFA.py:97,114 with max_iter=10 and defaults lr=1, max_eval=12, tolerance_grad=1e-7, tolerance_change=1e-9,
history_size=100, no line search) decision for decision, but keeps the parameters, the gradient and the (s, y)
history in flat device buffers and runs the vector algebra in three launches per iteration instead of ~4*history:
    pcfa_lbfgs_store_pair  (y, s into the ring buffers, <y,s>, <y,y>)
    pcfa_lbfgs_direction   (two-loop recursion, <g,d>, max|d|: one cooperative launch)
    flat_param += t * d
Host round trips per iteration: two small D2H reads (the curvature test and the stopping tests need the same scalars
torch's implementation syncs on, one by one)."""
from __future__ import annotations

import torch

from . import _lib


class DeviceLBFGS:
    def __init__(self, flat_param: torch.Tensor, flat_grad: torch.Tensor, lr=1.0, max_iter=10, max_eval=None,
                 tolerance_grad=1e-7, tolerance_change=1e-9, history_size=100):
        if not (flat_param.is_cuda and flat_param.dtype == torch.float32 and flat_param.dim() == 1 and flat_param.is_contiguous()):
            raise RuntimeError("DeviceLBFGS: flat_param must be a contiguous 1-D CUDA float32 tensor (pcfa_b200 has no CPU path)")
        if flat_grad.shape != flat_param.shape or not flat_grad.is_cuda or flat_grad.dtype != torch.float32:
            raise RuntimeError("DeviceLBFGS: flat_grad must match flat_param")
        if history_size > 128:
            raise ValueError("history_size > 128 is not supported")
        self.p, self.g = flat_param, flat_grad
        self.lr, self.max_iter = float(lr), int(max_iter)
        self.max_eval = int(max_eval) if max_eval is not None else self.max_iter * 5 // 4
        self.tolerance_grad, self.tolerance_change, self.m = float(tolerance_grad), float(tolerance_change), int(history_size)
        n, dev = flat_param.numel(), flat_param.device
        self.n = n
        self.S = self.Y = None                                   # allocated on first use ([m, n] each)
        self.ro = torch.zeros(self.m, device=dev)
        self.hdiag = torch.ones(1, device=dev)
        self.d = torch.empty(n, device=dev)
        self.g_prev = torch.empty(n, device=dev)
        self.sc = torch.zeros(4, device=dev)
        lib = _lib.load()
        self.ws = torch.empty(lib.pcfa_lbfgs_workspace_bytes(), device=dev, dtype=torch.uint8)
        self.state = dict(func_evals=0, n_iter=0, t=None, start=0, num_old=0, prev_loss=None, have_prev=False)

    # history slot that the next pair goes to, ring semantics of old_dirs.pop(0)/append
    def _push_slot(self):
        st = self.state
        if st["num_old"] < self.m:
            slot = (st["start"] + st["num_old"]) % self.m
            st["num_old"] += 1
        else:
            slot = st["start"]
            st["start"] = (st["start"] + 1) % self.m
        return slot

    def step(self, closure):
        lib, st, s = _lib.load(), self.state, _lib.stream()
        P = _lib.ptr
        orig_loss = closure()
        loss, gmax = torch.stack([orig_loss.detach().reshape(()), self.g.abs().max()]).tolist()
        current_evals = 1
        st["func_evals"] += 1
        if gmax <= self.tolerance_grad:
            return orig_loss
        t = st["t"]
        n_iter = 0
        while n_iter < self.max_iter:
            n_iter += 1
            st["n_iter"] += 1
            if st["n_iter"] == 1:
                torch.neg(self.g, out=self.d)
                st["start"], st["num_old"] = 0, 0
                self.hdiag.fill_(1.0)
                gtd = -float(self.g.dot(self.g))
                dmax = None
            else:
                if self.S is None:
                    self.S = torch.empty(self.m, self.n, device=self.p.device)
                    self.Y = torch.empty(self.m, self.n, device=self.p.device)
                # candidate pair into the slot it would occupy; the ring only advances if <y,s> > 1e-10 (lbfgs.py)
                cand = (st["start"] + st["num_old"]) % self.m if st["num_old"] < self.m else st["start"]
                if st["num_old"] == self.m:
                    # the slot to be overwritten still belongs to the history if the pair is rejected: stage in d-sized scratch
                    s_slot, y_slot = self._scratch()
                else:
                    s_slot, y_slot = self.S[cand], self.Y[cand]
                _lib.check(lib.pcfa_lbfgs_store_pair(P(self.g), P(self.g_prev), P(self.d), float(t), P(s_slot), P(y_slot), P(self.sc),
                                                     P(self.ws), self.n, s), "pcfa_lbfgs_store_pair")
                ys, yy = self.sc[:2].tolist()
                if ys > 1e-10:
                    slot = self._push_slot()
                    if s_slot.data_ptr() != self.S[slot].data_ptr():
                        self.S[slot].copy_(s_slot); self.Y[slot].copy_(y_slot)
                    self.ro[slot] = 1.0 / ys
                    self.hdiag.fill_(ys / yy)
                _lib.check(lib.pcfa_lbfgs_direction(P(self.S), P(self.Y), P(self.ro), P(self.g), P(self.hdiag), P(self.d), P(self.sc[2:]),
                                                    P(self.ws), self.n, self.m, st["start"], st["num_old"], s), "pcfa_lbfgs_direction")
                gtd, dmax = self.sc[2:4].tolist()
            if not st["have_prev"] or st["n_iter"] == 1:
                self.g_prev.copy_(self.g)
                st["have_prev"] = True
            prev_loss = loss
            if st["n_iter"] == 1:
                t = min(1.0, 1.0 / float(self.g.abs().sum())) * self.lr
            else:
                t = self.lr
            if gtd > -self.tolerance_change:
                break
            self.p.add_(self.d, alpha=t)
            ls_func_evals = 0
            if n_iter != self.max_iter:
                with torch.enable_grad():
                    l = closure()
                loss, gmax = torch.stack([l.detach().reshape(()), self.g.abs().max()]).tolist()
                ls_func_evals = 1
            current_evals += ls_func_evals
            st["func_evals"] += ls_func_evals
            if n_iter == self.max_iter:
                break
            if current_evals >= self.max_eval:
                break
            if gmax <= self.tolerance_grad:
                break
            if dmax is None:
                dmax = float(self.d.abs().max())
            if dmax * t <= self.tolerance_change:
                break
            if abs(loss - prev_loss) < self.tolerance_change:
                break
        st["t"], st["prev_loss"] = t, prev_loss
        return orig_loss

    def _scratch(self):
        if not hasattr(self, "_scr"):
            self._scr = torch.empty(2, self.n, device=self.p.device)
        return self._scr[0], self._scr[1]
