"""Fused PCFA objective on B200: box-constraint input transform + loss + penalty.

Two layers:

* reference-shaped, autograd-aware callables with the reference's names and argument meaning —
  `scaled_input` (the pre-processing of ScaledInputModel.forward, helper_functions/own_models.py:
  62-85), `extract_deltas`, `extract_deltas_joint` (attack_PCFA.py:20-37) and
  `loss_delta_constraint` (helper_functions/losses.py:200-230) — each backed by the CUDA kernels;
* `FusedObjective`, the closure the attack loop actually runs: 2 box kernels, the network, 2 loss
  kernels, the network's backward, 2 box-gradient kernels.  No .cpu() hop, no host sync, every
  buffer preallocated, so the whole evaluation can be captured in a CUDA graph.
"""
from __future__ import annotations

import torch
from torch.autograd import Function

from . import _lib

BOX_COV, BOX_CLIP, BOX_JOINT, BOX_UNIVERSAL = 0, 1, 2, 3
BOX_PARTIALS = 1024
LOSS_TYPES = {"aee": 0, "mse": 1, "cosim": 2}


def box_mode(boxconstraint: str, joint: bool = False, universal: bool = False) -> int:
    if universal:
        return BOX_UNIVERSAL
    if joint:
        if boxconstraint == "change_of_variables":
            # attack_PCFA.py:91-92
            raise ValueError("Training a --joint_perturbation with --boxconstraint=change_of_variables "
                             "is not defined. Please use --boxconstraint=clipping.")
        return BOX_JOINT
    return BOX_COV if boxconstraint == "change_of_variables" else BOX_CLIP


# ------------------------------------------------------------------------------- raw kernel calls
def _box_forward(var, image, amax, amin, net_in, delta_out, partials, mode, eps_box, scale):
    lib = _lib.load()
    B = image.shape[0]
    chw = image[0].numel()
    _lib.check(lib.pcfa_box_forward(_lib.ptr(var), _lib.ptr(image), _lib.ptr(amax), _lib.ptr(amin),
                                    _lib.ptr(net_in), _lib.ptr(delta_out), _lib.ptr(partials), mode, B,
                                    chw, float(eps_box), float(scale), _lib.stream()), "pcfa_box_forward")


def _box_backward(var, image, amax, amin, gnet, loss_terms, gvar, accumulate, mode, eps_box, scale):
    lib = _lib.load()
    B = image.shape[0]
    chw = image[0].numel()
    _lib.check(lib.pcfa_box_backward(_lib.ptr(var), _lib.ptr(image), _lib.ptr(amax), _lib.ptr(amin),
                                     _lib.ptr(gnet), _lib.ptr(loss_terms), _lib.ptr(gvar),
                                     int(accumulate), mode, B, chw, float(eps_box), float(scale),
                                     _lib.stream()), "pcfa_box_backward")


def _objective_loss(flow, target, p1, p2, w1, w2, terms, gflow, ws, loss_type, pad_top, pad_left,
                    numel_total, delta_bound, mu):
    lib = _lib.load()
    B, _, Hp, Wp = flow.shape
    _, _, H, W = target.shape
    _lib.check(lib.pcfa_objective_loss(_lib.ptr(flow), _lib.ptr(target), _lib.ptr(p1), _lib.ptr(p2),
                                       float(w1), float(w2), _lib.ptr(terms), _lib.ptr(gflow),
                                       _lib.ptr(ws), loss_type, B, H, W, Hp, Wp, pad_top, pad_left,
                                       float(numel_total), float(delta_bound), float(mu),
                                       _lib.stream()), "pcfa_objective_loss")


def _loss_workspace(device):
    lib = _lib.load()
    return torch.empty(lib.pcfa_objective_workspace_bytes(), device=device, dtype=torch.uint8)


# ------------------------------------------------------------------- reference-shaped autograd ops
class _ScaledInputFn(Function):
    """net_in = scale * clamp(transform(var [, image]), 0, 1)  with the transform chosen by mode."""

    @staticmethod
    def forward(ctx, var, image, mode, eps_box, scale):
        var, image = var.contiguous(), image.contiguous()
        _lib.require_cuda(var, image, name="ScaledInputModel")
        net_in = torch.empty_like(image)
        partials = torch.empty(BOX_PARTIALS, device=image.device, dtype=torch.float32)
        # JOINT here only needs x = clamp(image + delta); pass image as both aux tensors (unused for x)
        aux = image if mode == BOX_JOINT else None
        _box_forward(var, image, aux, aux, net_in, None, partials, mode, eps_box, scale)
        ctx.save_for_backward(var, image)
        ctx.cfg = (mode, eps_box, scale)
        return net_in

    @staticmethod
    def backward(ctx, gnet):
        var, image = ctx.saved_tensors
        mode, eps_box, scale = ctx.cfg
        gvar = torch.empty_like(var)
        aux = image if mode == BOX_JOINT else None
        _box_backward(var, image, aux, aux, gnet.contiguous(), None, gvar, False, mode, eps_box, scale)
        return gvar, None, None, None, None


def scaled_input(image, delta=None, *, var_change=False, eps_box=0.0, make_unit_input=False):
    """Pre-processing of ScaledInputModel.forward for ONE image (own_models.py:62-85):
    optional +delta (a [C,H,W] or [1,C,H,W] delta is broadcast over the batch like .repeat),
    optional change of variables, clamp to [0,1], optional x255."""
    scale = 255.0 if make_unit_input else 1.0
    if delta is not None:
        if var_change:
            raise ValueError("delta together with variable_change is not a configuration the attack uses")
        d = delta
        universal = d.dim() == 3 or (d.dim() == 4 and d.shape[0] == 1 and image.shape[0] != 1)
        if universal:
            return _ScaledInputFn.apply(d.reshape(image.shape[1:]), image, BOX_UNIVERSAL, eps_box, scale)
        return _ScaledInputFn.apply(d.reshape(image.shape), image, BOX_JOINT, eps_box, scale)
    mode = BOX_COV if var_change else BOX_CLIP
    return _ScaledInputFn.apply(image, image, mode, eps_box, scale)


class _ExtractDeltaFn(Function):
    @staticmethod
    def forward(ctx, var, image, amax, amin, mode, eps_box):
        var, image = var.contiguous(), image.contiguous()
        _lib.require_cuda(var, image, amax, amin, name="extract_deltas")
        delta = torch.empty_like(image)
        scratch = torch.empty_like(image)
        partials = torch.empty(BOX_PARTIALS, device=image.device, dtype=torch.float32)
        _box_forward(var, image, amax, amin, scratch, delta, partials, mode, eps_box, 1.0)
        ctx.save_for_backward(var, amax, amin)
        ctx.cfg = (mode, eps_box)
        return delta

    @staticmethod
    def backward(ctx, gdelta):
        lib = _lib.load()
        var, amax, amin = ctx.saved_tensors
        mode, eps_box = ctx.cfg
        gvar = torch.empty_like(var)
        _lib.check(lib.pcfa_box_delta_backward(_lib.ptr(var), _lib.ptr(amax), _lib.ptr(amin),
                                               _lib.ptr(gdelta.contiguous()), _lib.ptr(gvar), 0, mode,
                                               var.numel(), float(eps_box), _lib.stream()),
                   "pcfa_box_delta_backward")
        return gvar, None, None, None, None, None


def extract_deltas(nw_input1, nw_input2, image1, image2, boxconstraint, eps_box=0.0):
    """attack_PCFA.py:20-29."""
    mode = BOX_COV if boxconstraint in ["change_of_variables"] else BOX_CLIP
    return (_ExtractDeltaFn.apply(nw_input1, image1, None, None, mode, eps_box),
            _ExtractDeltaFn.apply(nw_input2, image2, None, None, mode, eps_box))


def extract_deltas_joint(nw_delta, images_max, images_min):
    """attack_PCFA.py:32-37 (returns the same tensor twice, like the reference)."""
    d = _ExtractDeltaFn.apply(nw_delta, images_max.contiguous(), images_max.contiguous(),
                              images_min.contiguous(), BOX_JOINT, 0.0)
    return d, d


class _LossDeltaConstraintFn(Function):
    @staticmethod
    def forward(ctx, pred, target, delta1, delta2, delta_bound, mu, loss_type):
        lib = _lib.load()
        squeeze = pred.dim() == 3
        p4 = (pred[None] if squeeze else pred).contiguous()
        t4 = (target[None] if target.dim() == 3 else target).contiguous()
        d1, d2 = delta1.contiguous(), delta2.contiguous()
        _lib.require_cuda(p4, t4, d1, d2, name="loss_delta_constraint")
        dev = p4.device
        parts = torch.empty((2, BOX_PARTIALS), device=dev, dtype=torch.float32)
        for k, d in enumerate((d1, d2)):
            _lib.check(lib.pcfa_sumsq_partials(_lib.ptr(d), d.numel(), _lib.ptr(parts[k]), _lib.stream()),
                       "pcfa_sumsq_partials")
        terms = torch.empty(4, device=dev, dtype=torch.float32)
        gflow = torch.empty_like(p4)
        _objective_loss(p4, t4, parts[0], parts[1], 1.0, 1.0, terms, gflow, _loss_workspace(dev),
                        loss_type, 0, 0, d1.numel() + d2.numel(), delta_bound, mu)
        ctx.save_for_backward(gflow, terms, d1, d2)
        ctx.squeeze = squeeze
        ctx.shapes = (delta1.shape, delta2.shape)
        return terms[0].clone()

    @staticmethod
    def backward(ctx, g):
        gflow, terms, d1, d2 = ctx.saved_tensors
        gp = gflow * g
        if ctx.squeeze:
            gp = gp[0]
        coef = terms[2] * g
        return gp, None, (d1 * coef).view(ctx.shapes[0]), (d2 * coef).view(ctx.shapes[1]), None, None, None


def loss_delta_constraint(pred, target, delta1, delta2, device=None, delta_bound=0.001, mu=100.,
                          f_type="aee"):
    """helper_functions/losses.py:200-230.  `device` is accepted for signature parity."""
    if f_type not in LOSS_TYPES:
        raise NotImplementedError("The requested loss type %s does not exist. Please choose one of "
                                  "'aee', 'mse' or 'cosim'" % (f_type))
    return _LossDeltaConstraintFn.apply(pred, target, delta1, delta2, float(delta_bound), float(mu),
                                        LOSS_TYPES[f_type])


# ------------------------------------------------------------------------------- the fused closure
class FusedObjective:
    """One PCFA closure evaluation (forward + backward) with preallocated buffers.

    net_forward(net_in1, net_in2) -> padded flow [B,2,Hp,Wp] (differentiable torch graph).
    images are the padded, [0,1]-ranged inputs [B,C,Hp,Wp]; target is the unpadded [B,2,H,W] flow;
    `pad` = (pad_top, pad_left) of InputPadder (ownutilities.py:26-33).

    Variables per mode (see box_mode):
      COV / CLIP : var1, var2 of image shape      JOINT: var1 = delta of image shape
      UNIVERSAL  : var1 (and var2 unless joint) of shape [C,Hp,Wp]
    """

    def __init__(self, net_forward, image1, image2, target, *, mode, joint, pad, eps_box, scale,
                 delta_bound, mu, loss="aee"):
        self.net_forward = net_forward
        self.image1, self.image2 = image1.contiguous(), image2.contiguous()
        self.target = target.contiguous()
        _lib.require_cuda(self.image1, self.image2, self.target, name="FusedObjective")
        self.mode, self.joint = mode, bool(joint)
        self.pad_top, self.pad_left = pad
        self.eps_box, self.scale = float(eps_box), float(scale)
        self.delta_bound, self.mu = float(delta_bound), float(mu)
        self.loss_type = LOSS_TYPES[loss]
        dev = self.image1.device
        self.amax = self.amin = None
        if mode == BOX_JOINT:
            self.amax = torch.max(self.image1, self.image2).contiguous()
            self.amin = torch.min(self.image1, self.image2).contiguous()
        var_numel = self.image1[0].numel() if mode == BOX_UNIVERSAL else self.image1.numel()
        self.numel_total = 2 * var_numel              # numel(delta1) + numel(delta2), losses.py:122-126
        self.net_in1 = torch.empty_like(self.image1)
        self.net_in2 = torch.empty_like(self.image2)
        self.partials = torch.empty((2, BOX_PARTIALS), device=dev, dtype=torch.float32)
        self.terms = torch.zeros(4, device=dev, dtype=torch.float32)
        self.ws = _loss_workspace(dev)
        self.gflow = None
        self.delta1 = self.delta2 = None

    def _forward_boxes(self, var1, var2, want_delta):
        v2 = var1 if self.joint else var2
        d1 = d2 = None
        if want_delta:
            shape = self.image1.shape[1:] if self.mode == BOX_UNIVERSAL else self.image1.shape
            d1 = torch.empty(shape, device=self.image1.device, dtype=torch.float32)
            d2 = d1 if self.joint and self.mode == BOX_JOINT else torch.empty_like(d1)
        _box_forward(var1, self.image1, self.amax, self.amin, self.net_in1, d1, self.partials[0],
                     self.mode, self.eps_box, self.scale)
        _box_forward(v2, self.image2, self.amax, self.amin, self.net_in2,
                     None if d2 is d1 else d2, self.partials[1], self.mode, self.eps_box, self.scale)
        self.delta1, self.delta2 = d1, d2

    def _loss(self, flow, want_grad):
        if want_grad and (self.gflow is None or self.gflow.shape != flow.shape):
            self.gflow = torch.empty_like(flow)
        # per-pair joint: delta' is the same for both images (both partial sets hold the same sum);
        # universal-joint: delta2 is delta1.  Either way weights (1, 1) count it twice over
        # numel_total = 2*numel, exactly like delta1 = delta2 in the reference.
        _objective_loss(flow, self.target, self.partials[0], self.partials[1], 1.0, 1.0, self.terms,
                        self.gflow if want_grad else None, self.ws, self.loss_type, self.pad_top,
                        self.pad_left, self.numel_total, self.delta_bound, self.mu)

    @torch.no_grad()
    def predict(self, var1, var2=None, want_delta=False):
        """Forward only: padded flow for the current variables (the reference's re-prediction,
        attack_PCFA.py:207-212); also refreshes loss terms."""
        self._forward_boxes(var1, var2, want_delta)
        flow = self.net_forward(self.net_in1, self.net_in2).contiguous()
        self._loss(flow, False)
        return flow

    def evaluate(self, var1, var2=None, grad1=None, grad2=None):
        """loss (0-dim device tensor view) and dL/dvar (written into grad1/grad2 if given)."""
        self._forward_boxes(var1, var2, False)
        n1 = self.net_in1.detach().requires_grad_(True)
        n2 = self.net_in2.detach().requires_grad_(True)
        with torch.enable_grad():
            flow = self.net_forward(n1, n2)
        flow_c = flow.detach().contiguous()
        self._loss(flow_c, True)
        g1, g2 = torch.autograd.grad(flow, [n1, n2], self.gflow)
        if grad1 is None:
            grad1 = torch.empty_like(var1)
        _box_backward(var1, self.image1, self.amax, self.amin, g1.contiguous(), self.terms, grad1, False,
                      self.mode, self.eps_box, self.scale)
        if self.joint:
            _box_backward(var1, self.image2, self.amax, self.amin, g2.contiguous(), self.terms, grad1,
                          True, self.mode, self.eps_box, self.scale)
            return self.terms[0], grad1, None
        if grad2 is None:
            grad2 = torch.empty_like(var2)
        _box_backward(var2, self.image2, self.amax, self.amin, g2.contiguous(), self.terms, grad2, False,
                      self.mode, self.eps_box, self.scale)
        return self.terms[0], grad1, grad2
