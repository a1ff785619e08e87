"""PCFA optimisation loops on the fused closure.

Mirrors attack_PCFA.py of the reference: `pcfa_attack` (:40-294, one image pair, disjoint or joint
perturbation, L-BFGS with max_iter=10 per outer step, best-delta bookkeeping :226-243),
`attack_l2` (:570-701, loop over pairs) and `attack_l2_universal` (:297-566, one shared delta,
one L-BFGS instance for the whole run).  Differences in mechanism, not in the optimisation problem:

  * a closure evaluation is `FusedObjective.evaluate` (optionally replayed from a CUDA graph), not
    ~200 eager launches plus a device→host→device hop of the flow (ownutilities.py:297);
  * the backward the reference runs before every optimizer.step and then discards
    (attack_PCFA.py:173, zeroed by the closure's zero_grad) is not executed;
  * torch.autograd.set_detect_anomaly (attack_PCFA.py:41) is off;
  * per-step metrics are reduced on the device and fetched with one copy per outer step;
  * multi-GPU: pairs are sharded over ranks (no communication); in universal mode every closure
    all-reduces [grad | loss] so all ranks take identical L-BFGS decisions.
"""
from __future__ import annotations

import math
import time
from dataclasses import dataclass, field

import torch

from . import objective as J
from .adapter import compute_flow, model_takes_unit_input, preprocess_img
from .dist import pack_reduce_unpack


def resolve_mu(mu: float, delta_bound: float, target: str) -> float:
    """attack_PCFA.py:302-307 / :578-583."""
    if mu == -1.:
        mu = 2500. / delta_bound
        if target not in ['zero']:
            mu = 1.5 * mu
    return mu


def get_target(target_name, flow_pred_init, custom_target_path="", device=None):
    """helper_functions/targets.py:89-114 (custom targets are .npy flows of matching size)."""
    if target_name == 'zero':
        return torch.zeros_like(flow_pred_init)
    if target_name == 'neg_flow':
        return -flow_pred_init
    if target_name == 'custom':
        # targets.py:60-80: [2,H,W] flow, cropped or reflect-padded (right / bottom) to the flow's size, repeated over
        # the batch.  Offline build: the file is a .npy array ([2,H,W], [H,W,2] or [1,2,H,W]) instead of a .flo/.png.
        import numpy as np
        import torch.nn.functional as F
        t = torch.from_numpy(np.load(custom_target_path).astype("float32")).to(flow_pred_init.device)
        if t.dim() == 4:
            t = t[0]
        if t.dim() != 3:
            raise ValueError("custom target must be a flow field, got shape %s" % (tuple(t.shape),))
        if t.shape[0] != 2 and t.shape[-1] == 2:
            t = t.permute(2, 0, 1)
        Hf, Wf = flow_pred_init.shape[-2:]
        if Wf < t.shape[-1]:
            t = t[:, :, :Wf]
        elif Wf > t.shape[-1]:
            t = F.pad(t[None], (0, Wf - t.shape[-1]), "reflect")[0]
        if Hf < t.shape[-2]:
            t = t[:, :Hf, :]
        elif Hf > t.shape[-2]:
            t = F.pad(t[None], (0, 0, 0, Hf - t.shape[-2]), "reflect")[0]
        if flow_pred_init.dim() == 4:
            t = t[None].repeat(flow_pred_init.shape[0], 1, 1, 1)
        return t.contiguous()
    raise ValueError("The specified target type '%s' is not defined" % target_name)


def avg_epe(a, b):
    """helper_functions/losses.py:3-30 (device tensor result)."""
    return torch.sum((a - b) ** 2, dim=-3).sqrt().mean()


@dataclass
class PairResult:
    aee_tgt: float = 0.
    aee_adv_tgt: float = 0.
    aee_adv_pred: float = 0.
    l2_delta1: float = 0.
    l2_delta2: float = 0.
    l2_delta12: float = 0.
    aee_adv_tgt_min: float = float('inf')
    aee_adv_pred_min: float = 0.
    l2_delta12_min: float = float('inf')
    closure_evals: int = 0
    delta1_best: torch.Tensor | None = None
    delta2_best: torch.Tensor | None = None
    flow_best: torch.Tensor | None = None
    history: list = field(default_factory=list)


class GraphedEvaluate:
    """FusedObjective.evaluate behind a CUDA graph: variables and gradients live in static buffers."""

    def __init__(self, fo: J.FusedObjective, var1, var2, use_graph=True, warmup=3, g1=None, g2=None):
        self.fo, self.var1, self.var2 = fo, var1, var2
        self.g1 = torch.empty_like(var1) if g1 is None else g1
        self.g2 = None if var2 is None else (torch.empty_like(var2) if g2 is None else g2)
        self.graph = None
        if use_graph:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(warmup):
                    fo.evaluate(var1, var2, self.g1, self.g2)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                fo.evaluate(var1, var2, self.g1, self.g2)

    def __call__(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self.fo.evaluate(self.var1, self.var2, self.g1, self.g2)
        return self.fo.terms[0]


class GraphedPredict:
    """FusedObjective.predict(var1, var2, want_delta=True) behind a CUDA graph (the per-outer-step re-prediction of
    attack_PCFA.py:212-224 is ~600 eager launches otherwise: as long as ten graphed closure evaluations' worth of
    host time)."""

    def __init__(self, fo: J.FusedObjective, var1, var2, use_graph=True):
        self.fo, self.var1, self.var2 = fo, var1, var2
        self.graph, self.out = None, None
        if use_graph:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side), torch.no_grad():
                fo.predict(var1, var2, want_delta=True)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph), torch.no_grad():
                self.out = fo.predict(var1, var2, want_delta=True)

    def __call__(self):
        if self.graph is not None:
            self.graph.replay()
            return self.out
        with torch.no_grad():
            return self.fo.predict(self.var1, self.var2, want_delta=True)


def _net_forward(model, net_name, iters=None):
    kw = {}
    if iters is not None and net_name in ("RAFT",):
        kw["iters"] = iters
    return lambda a, b: compute_flow(model, net_name, a, b, test_mode=True, **kw)


def pcfa_attack(model, net_name, image1, image2, *, steps=20, delta_bound=0.005, mu=-1., target='zero',
                loss='aee', joint_perturbation=False, boxconstraint='change_of_variables', eps_box=1e-7,
                custom_target_path="", use_graph=True, iters=None, lbfgs_max_iter=10, keep_best=True,
                reduce_hook=None, lbfgs="device"):
    """One image pair ([B,3,H,W] in [0,255] on the device).  Returns a PairResult.
    Follows attack_PCFA.py:40-294 step for step (see module docstring for the mechanical differences)."""
    device = image1.device
    unit = model_takes_unit_input(net_name)
    if not unit:                                   # attack_PCFA.py:58-60; unit-input nets are scaled in preprocess_img
        image1, image2 = image1 / 255., image2 / 255.
    padder, (image1, image2) = preprocess_img(net_name, image1, image2)
    image1, image2 = image1.contiguous(), image2.contiguous()
    mode = J.box_mode(boxconstraint, joint=joint_perturbation)
    mu = resolve_mu(mu, delta_bound, target)
    scale = 1.0 if unit else 255.0
    fwd = _net_forward(model, net_name, iters)

    # variables (attack_PCFA.py:84-114)
    if joint_perturbation:
        var1, var2 = torch.zeros_like(image1), None
    elif mode == J.BOX_COV:
        var1 = torch.atanh(2. * (1. - eps_box) * image1 - (1 - eps_box)).contiguous()
        var2 = torch.atanh(2. * (1. - eps_box) * image2 - (1 - eps_box)).contiguous()
    else:
        var1, var2 = image1.clone(), image2.clone()

    H, W = padder.get_dimensions()
    B = image1.shape[0]
    fo = J.FusedObjective(fwd, image1, image2, torch.zeros(B, 2, H, W, device=device), mode=mode,
                          joint=joint_perturbation, pad=padder.top_left, eps_box=eps_box, scale=scale,
                          delta_bound=delta_bound, mu=mu, loss=loss)
    flow_init = padder.unpad(fo.predict(var1, var2)).contiguous().clone()
    fo.target.copy_(get_target(target, flow_init, custom_target_path))
    res = PairResult(aee_tgt=float(avg_epe(fo.target, flow_init)))

    counter = [0]
    if lbfgs == "device":
        # f-1: variables and gradients are slices of two flat device buffers; pcfa_b200.lbfgs.DeviceLBFGS runs
        # torch.optim.LBFGS's update rule on them in three launches per iteration
        from .lbfgs import DeviceLBFGS
        n1 = var1.numel()
        n2 = 0 if var2 is None else var2.numel()
        flat_p, flat_g = torch.empty(n1 + n2, device=device), torch.zeros(n1 + n2, device=device)
        flat_p[:n1].copy_(var1.reshape(-1))
        var1 = flat_p[:n1].view_as(var1)
        g1, g2 = flat_g[:n1].view_as(var1), None
        if var2 is not None:
            flat_p[n1:].copy_(var2.reshape(-1))
            var2 = flat_p[n1:].view_as(var2)
            g2 = flat_g[n1:].view_as(var2)
        ev = GraphedEvaluate(fo, var1, var2, use_graph=use_graph, g1=g1, g2=g2)
        optimizer = DeviceLBFGS(flat_p, flat_g, max_iter=lbfgs_max_iter)

        def closure():
            counter[0] += 1
            loss_t = ev()
            if reduce_hook is not None:
                loss_t = reduce_hook(loss_t, ev.g1, ev.g2)
            return loss_t
    else:
        ev = GraphedEvaluate(fo, var1, var2, use_graph=use_graph)
        params = [var1] if var2 is None else [var1, var2]
        for p in params:
            p.requires_grad_(True)
        optimizer = torch.optim.LBFGS(params, max_iter=lbfgs_max_iter)

        def closure():
            counter[0] += 1
            loss_t = ev()
            if reduce_hook is not None:
                loss_t = reduce_hook(loss_t, ev.g1, ev.g2)
            var1.grad = ev.g1
            if var2 is not None:
                var2.grad = ev.g2
            return loss_t

    below = False
    repredict = GraphedPredict(fo, var1.detach(), None if var2 is None else var2.detach(), use_graph=use_graph)
    for step in range(steps):
        optimizer.step(closure)
        with torch.no_grad():
            flow_pred = padder.unpad(repredict())
            d1, d2 = fo.delta1, fo.delta2 if fo.delta2 is not None else fo.delta1
            stats = torch.stack([avg_epe(flow_pred, fo.target), avg_epe(flow_pred, flow_init),
                                 d1.pow(2).sum(), d2.pow(2).sum()]).tolist()     # one D2H per outer step
        aee_adv_tgt, aee_adv_pred, s1, s2 = stats
        n1 = d1.numel()
        l2_1, l2_2 = math.sqrt(s1 / n1), math.sqrt(s2 / d2.numel())
        l2_12 = math.sqrt(s1 + s2) / math.sqrt(n1 + d2.numel())                   # losses.py:91-107
        res.aee_adv_tgt, res.aee_adv_pred = aee_adv_tgt, aee_adv_pred
        res.l2_delta1, res.l2_delta2, res.l2_delta12 = l2_1, l2_2, l2_12
        update = False                                                           # attack_PCFA.py:226-243
        if not below:
            if l2_12 < res.l2_delta12_min or (l2_12 == res.l2_delta12_min and aee_adv_tgt < res.aee_adv_tgt_min):
                update = True
                if l2_12 <= delta_bound:
                    below = True
        elif l2_12 <= delta_bound and aee_adv_tgt < res.aee_adv_tgt_min:
            update = True
        if update:
            res.l2_delta12_min, res.aee_adv_tgt_min, res.aee_adv_pred_min = l2_12, aee_adv_tgt, aee_adv_pred
            if keep_best:
                res.delta1_best, res.delta2_best = d1.detach().clone(), d2.detach().clone()
                res.flow_best = flow_pred.detach().clone()
        res.history.append(dict(step=step, aee_adv_tgt=aee_adv_tgt, aee_adv_pred=aee_adv_pred, l2_delta12=l2_12,
                                loss=float(fo.terms[0]), closure_evals=counter[0], t=time.perf_counter()))
    res.closure_evals = counter[0]
    return res


class UniversalAttack:
    """attack_l2_universal (attack_PCFA.py:297-566): one delta (or delta1, delta2) shared by every
    pair, clipping box constraint, one L-BFGS instance whose history persists across batches/epochs.
    With torch.distributed initialised, each rank evaluates its shard of the batch and the closure
    all-reduces [grad_delta1 | grad_delta2 | loss] (sum, then / world_size)."""

    def __init__(self, model, net_name, image_shape, device, *, delta_bound=0.005, mu=-1., target='zero', loss='aee',
                 joint_perturbation=False, eps_box=1e-7, iters=None, lbfgs_max_iter=10, use_graph=False, lbfgs="device",
                 custom_target_path=""):
        self.model, self.net_name, self.device = model, net_name, device
        self.custom_target_path = custom_target_path
        self.delta_bound, self.target, self.loss = delta_bound, target, loss
        self.mu = resolve_mu(mu, delta_bound, target)
        self.joint, self.eps_box, self.iters, self.use_graph = joint_perturbation, eps_box, iters, use_graph
        self.unit = model_takes_unit_input(net_name)
        dummy = torch.zeros(1, 3, *image_shape, device=device)
        self.padder, (padded,) = preprocess_img(net_name, dummy)
        chw = padded.shape[1:]
        n1 = int(torch.Size(chw).numel())
        n = n1 if joint_perturbation else 2 * n1
        self.flat = torch.zeros(n + 1, device=device)            # fused all-reduce buffer [grads | loss]
        self.device_lbfgs = lbfgs == "device" and torch.device(device).type == "cuda"
        if self.device_lbfgs:
            # f-1: deltas are slices of one flat parameter buffer, their gradients slices of the all-reduce buffer
            from .lbfgs import DeviceLBFGS
            self.flat_p = torch.zeros(n, device=device)
            self.delta1 = self.flat_p[:n1].view(chw)
            self.delta2 = None if joint_perturbation else self.flat_p[n1:].view(chw)
            self.g1 = self.flat[:n1].view(chw)
            self.g2 = None if joint_perturbation else self.flat[n1:n].view(chw)
            self.optimizer = DeviceLBFGS(self.flat_p, self.flat[:n], max_iter=lbfgs_max_iter)
        else:
            self.delta1 = torch.zeros(chw, device=device, requires_grad=True)
            self.delta2 = None if joint_perturbation else torch.zeros(chw, device=device, requires_grad=True)
            params = [self.delta1] if self.delta2 is None else [self.delta1, self.delta2]
            self.optimizer = torch.optim.LBFGS(params, max_iter=lbfgs_max_iter)
            self.g1 = self.g2 = None
        self.closure_evals = 0

    def _dist(self):
        import torch.distributed as dist
        return dist if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 else None

    def run_batch(self, image1, image2, steps):
        """image1/2: this rank's shard of the batch, [b,3,H,W] in [0,255]."""
        if not self.unit:
            image1, image2 = image1 / 255., image2 / 255.
        _, (image1, image2) = preprocess_img(self.net_name, image1, image2)
        image1, image2 = image1.contiguous(), image2.contiguous()
        H, W = self.padder.get_dimensions()
        b = image1.shape[0]
        scale = 1.0 if self.unit else 255.0
        fwd = _net_forward(self.model, self.net_name, self.iters)
        zero = torch.zeros_like(self.delta1)
        fo = J.FusedObjective(fwd, image1, image2, torch.zeros(b, 2, H, W, device=self.device), mode=J.BOX_UNIVERSAL,
                              joint=self.joint, pad=self.padder.top_left, eps_box=self.eps_box, scale=scale,
                              delta_bound=self.delta_bound, mu=self.mu, loss=self.loss)
        flow_init = self.padder.unpad(fo.predict(zero, None if self.joint else zero)).contiguous().clone()
        fo.target.copy_(get_target(self.target, flow_init, self.custom_target_path))
        d1 = self.delta1.detach()
        d2 = None if self.delta2 is None else self.delta2.detach()
        ev = GraphedEvaluate(fo, d1, d2, use_graph=self.use_graph, g1=self.g1, g2=self.g2)
        dist = self._dist()
        def closure():
            self.closure_evals += 1
            loss_t = ev()
            if dist is not None:
                loss_t = pack_reduce_unpack(self.flat, loss_t, ev.g1, ev.g2)
            if not self.device_lbfgs:
                self.delta1.grad = ev.g1
                if self.delta2 is not None:
                    self.delta2.grad = ev.g2
            return loss_t

        out = []
        for _ in range(steps):
            self.optimizer.step(closure)
            with torch.no_grad():
                flow_pred = self.padder.unpad(fo.predict(d1, d2))
                out.append(torch.stack([avg_epe(flow_pred, fo.target), avg_epe(flow_pred, flow_init)]))
        stats = torch.stack(out)
        if dist is not None:
            dist.all_reduce(stats)
            stats /= dist.get_world_size()
        return stats.tolist()

    def l2_norms(self):
        d1 = self.delta1.detach()
        d2 = d1 if self.delta2 is None else self.delta2.detach()
        s1, s2 = float(d1.pow(2).sum()), float(d2.pow(2).sum())
        return math.sqrt(s1 / d1.numel()), math.sqrt(s2 / d2.numel()), math.sqrt((s1 + s2) / (d1.numel() + d2.numel()))
