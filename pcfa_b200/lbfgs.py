"""On-device L-BFGS (row f-1 of SURVEY.md section 8).

`DeviceLBFGS` follows torch.optim.LBFGS.step (torch/optim/lbfgs.py — the optimiser the reference constructs at
attack_PCFA.py:97,114 with max_iter=10 and defaults lr=1, max_eval=12, tolerance_grad=1e-7, tolerance_change=1e-9,
history_size=100, no line search) decision for decision, but keeps the parameters, the gradient and the (s, y)
history in flat device buffers and runs the vector algebra in four launches per iteration instead of ~4*history:
    pcfa_lbfgs_update_history  (<y,s>, <y,y>; curvature test, ring-buffer update, ro, H_diag decided on the device)
    pcfa_lbfgs_direction_compact (default; the same H_k in the compact Byrd-Nocedal-Schnabel form: one pass for S^T g, Y^T g,
                                a single-CTA double-precision solve, one pass for d and param += t*d unless
                                <g,d> > -tolerance_change — no chain of 2*history grid barriers), or
    pcfa_lbfgs_direction_step  (direction="two_loop": torch's two-loop recursion literally, one cooperative launch)
Host round trips: ONE small D2H read per inner iteration, after its closure (the stopping tests torch syncs on one by
one, packed), none on the last iteration of a step."""
from __future__ import annotations

import torch

from . import _lib


class DeviceLBFGS:
    def __init__(self, flat_param: torch.Tensor, flat_grad: torch.Tensor, lr=1.0, max_iter=10, max_eval=None,
                 tolerance_grad=1e-7, tolerance_change=1e-9, history_size=100, direction="compact"):
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
        self.sc = torch.zeros(4, device=dev)                     # {<y,s>, <y,y>, <g,d>, max|d|}
        self.ring = torch.zeros(2, 4, dtype=torch.int32, device=dev)   # double-buffered {start, num_old, accepted, -}
        self.cur = 0
        if direction not in ("compact", "two_loop"):
            raise ValueError("direction must be 'compact' or 'two_loop'")
        self.direction = direction
        lib = _lib.load()
        self.ws = torch.empty(lib.pcfa_lbfgs_workspace_bytes(), device=dev, dtype=torch.uint8)
        # compact direction: S^T Y, Y^T Y (double) and the previous S^T g, Y^T g live here between iterations
        self.cstate = torch.zeros(lib.pcfa_lbfgs_compact_workspace_bytes(self.m) // 8 + 1, device=dev, dtype=torch.float64)
        self.state = dict(func_evals=0, n_iter=0, t=None, prev_loss=None)

    def step(self, closure):
        """torch.optim.LBFGS.step (no line search) with ONE host round trip per inner iteration: after each closure the
        host reads one packed block {loss, max|g|, <y,s>, <y,y>, <g,d>, max|d|}; the history bookkeeping
        (pcfa_lbfgs_update_history) and the pre-update break test (pcfa_lbfgs_direction_step) are device-side predicates.
        The very first iteration of the optimiser's life (d = -g, t = min(1, 1/|g|_1)) keeps its extra reads.
        One mechanical difference: when `<g,d> > -tolerance_change` fires (torch breaks before the update), the update is
        skipped on the device but the closure of that iteration has already been launched; it re-evaluates the unchanged
        point, is not counted in func_evals, and changes neither the parameters nor the history."""
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
            first = st["n_iter"] == 1
            if first:
                torch.neg(self.g, out=self.d)
                self.ring.zero_()
                self.cur = 0
                self.cstate.zero_()
                self.hdiag.fill_(1.0)
                self.g_prev.copy_(self.g)
                gtd = -float(self.g.dot(self.g))
                t = min(1.0, 1.0 / float(self.g.abs().sum())) * self.lr
                if gtd > -self.tolerance_change:
                    prev_loss = loss
                    break
                self.p.add_(self.d, alpha=t)
            else:
                if self.S is None:
                    self.S = torch.empty(self.m, self.n, device=self.p.device)
                    self.Y = torch.empty(self.m, self.n, device=self.p.device)
                ring_in, ring_out = self.ring[self.cur], self.ring[self.cur ^ 1]
                _lib.check(lib.pcfa_lbfgs_update_history(P(self.g), P(self.g_prev), P(self.d), float(t), P(self.S), P(self.Y), P(self.ro),
                                                         P(self.hdiag), P(ring_in), P(ring_out), P(self.sc), P(self.ws), self.n, self.m, s),
                           "pcfa_lbfgs_update_history")
                self.cur ^= 1
                t = self.lr
                if self.direction == "compact":
                    _lib.check(lib.pcfa_lbfgs_direction_compact(P(self.S), P(self.Y), P(self.g), P(self.hdiag), P(self.d), P(ring_out), P(self.sc),
                                                                P(self.p), float(t), float(self.tolerance_change), P(self.sc[2:]),
                                                                P(self.cstate), self.n, self.m, s), "pcfa_lbfgs_direction_compact")
                else:
                    _lib.check(lib.pcfa_lbfgs_direction_step(P(self.S), P(self.Y), P(self.ro), P(self.g), P(self.hdiag), P(self.d), P(ring_out),
                                                             P(self.p), float(t), float(self.tolerance_change), P(self.sc[2:]), P(self.ws),
                                                             self.n, self.m, s), "pcfa_lbfgs_direction_step")
            prev_loss = loss
            if n_iter == self.max_iter:
                break                                         # torch evaluates no closure on the last iteration: no read-back either
            with torch.enable_grad():
                l = closure()
            vals = torch.cat([torch.stack([l.detach().reshape(()).float(), self.g.abs().max()]), self.sc]).tolist()
            new_loss, gmax = vals[0], vals[1]
            if not first:
                gtd, dmax = vals[4], vals[5]
                if gtd > -self.tolerance_change:              # the update was skipped on the device: torch's pre-update break
                    break
            else:
                dmax = None
            loss = new_loss
            current_evals += 1
            st["func_evals"] += 1
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

    @property
    def history(self):
        """(start, num_old) of the device-side ring (one D2H read; diagnostics and tests only)."""
        a = self.ring[self.cur].tolist()
        return a[0], a[1]
