"""numpy front-end of the C oracle (oracle/oracle.c) plus the numpy restatement of the PCFA
objective.  TEST INFRASTRUCTURE ONLY — see oracle/__init__.py."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(_build.build()))
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


# ---------------------------------------------------------------- spatial correlation sampler
def scs_forward(in1, in2, kernel_size=1, patch_size=1, stride=1, padding=0, dilation=1, dilation_patch=1):
    in1, in2 = _f(in1), _f(in2)
    B, Cc, iH, iW = in1.shape
    kH, kW = _pair(kernel_size); pH, pW = _pair(patch_size); dH, dW = _pair(stride)
    padH, padW = _pair(padding); dilH, dilW = _pair(dilation); dpH, dpW = _pair(dilation_patch)
    oH = (iH + 2 * padH - ((kH - 1) * dilH + 1)) // dH + 1
    oW = (iW + 2 * padW - ((kW - 1) * dilW + 1)) // dW + 1
    out = np.zeros((B, pH, pW, oH, oW), np.float32)
    lib().oracle_scs_forward(_p(in1), _p(in2), _p(out), B, Cc, iH, iW, kH, kW, pH, pW, padH, padW, dilH,
                             dilW, dpH, dpW, dH, dW)
    return out


def scs_backward(in1, in2, gout, kernel_size=1, patch_size=1, stride=1, padding=0, dilation=1, dilation_patch=1):
    in1, in2, gout = _f(in1), _f(in2), _f(gout)
    B, Cc, iH, iW = in1.shape
    kH, kW = _pair(kernel_size); pH, pW = _pair(patch_size); dH, dW = _pair(stride)
    padH, padW = _pair(padding); dilH, dilW = _pair(dilation); dpH, dpW = _pair(dilation_patch)
    g1, g2 = np.zeros_like(in1), np.zeros_like(in2)
    lib().oracle_scs_backward(_p(in1), _p(in2), _p(gout), _p(g1), _p(g2), B, Cc, iH, iW, kH, kW, pH, pW,
                              padH, padW, dilH, dilW, dpH, dpW, dH, dW)
    return g1, g2


# ---------------------------------------------------------------- FlowNet2 correlation
def fn2corr_sizes(H, W, pad, ks, md, s1, s2):
    oc, oh, ow = C.c_int(), C.c_int(), C.c_int()
    lib().oracle_fn2corr_sizes(H, W, pad, ks, md, s1, s2, C.byref(oc), C.byref(oh), C.byref(ow))
    return oc.value, oh.value, ow.value


def fn2corr_forward(in1, in2, pad, ks, md, s1, s2):
    in1, in2 = _f(in1), _f(in2)
    B, Cc, H, W = in1.shape
    oc, oh, ow = fn2corr_sizes(H, W, pad, ks, md, s1, s2)
    out = np.zeros((B, oc, oh, ow), np.float32)
    lib().oracle_fn2corr_forward(_p(in1), _p(in2), _p(out), B, Cc, H, W, pad, ks, md, s1, s2)
    return out


def fn2corr_backward(in1, in2, gout, pad, ks, md, s1, s2):
    in1, in2, gout = _f(in1), _f(in2), _f(gout)
    B, Cc, H, W = in1.shape
    g1, g2 = np.zeros_like(in1), np.zeros_like(in2)
    lib().oracle_fn2corr_backward(_p(in1), _p(in2), _p(gout), _p(g1), _p(g2), B, Cc, H, W, pad, ks, md, s1, s2)
    return g1, g2


# ---------------------------------------------------------------- resample2d / channelnorm
def resample2d_forward(img, flow, bilinear=True):
    img, flow = _f(img), _f(flow)
    B, Cc, H, W = img.shape
    _, _, oH, oW = flow.shape
    out = np.zeros((B, Cc, oH, oW), np.float32)
    lib().oracle_resample2d_forward(_p(img), _p(flow), _p(out), B, Cc, H, W, oH, oW, int(bilinear))
    return out


def resample2d_backward(img, flow, gout):
    img, flow, gout = _f(img), _f(flow), _f(gout)
    B, Cc, H, W = img.shape
    _, _, oH, oW = flow.shape
    gi, gf = np.zeros_like(img), np.zeros_like(flow)
    lib().oracle_resample2d_backward(_p(img), _p(flow), _p(gout), _p(gi), _p(gf), B, Cc, H, W, oH, oW)
    return gi, gf


def channelnorm_forward(x):
    x = _f(x)
    B, Cc, H, W = x.shape
    out = np.zeros((B, 1, H, W), np.float32)
    lib().oracle_channelnorm_forward(_p(x), _p(out), B, Cc, H, W)
    return out


def channelnorm_backward(x, out, gout):
    x, out, gout = _f(x), _f(out), _f(gout)
    B, Cc, H, W = x.shape
    gx = np.zeros_like(x)
    lib().oracle_channelnorm_backward(_p(x), _p(out), _p(gout), _p(gx), B, Cc, H, W)
    return gx


# ---------------------------------------------------------------- all-pairs pyramid + lookup
def pyramid_layout(B, H, W, levels):
    off = (C.c_int64 * (levels + 1))(); hs = (C.c_int * levels)(); ws = (C.c_int * levels)()
    lib().oracle_pyramid_layout.restype = C.c_int64
    lib().oracle_pyramid_layout(B, H, W, levels, off, hs, ws)
    return list(off), list(hs), list(ws)


def corr_pyramid_forward(f1, f2, levels=4):
    f1, f2 = _f(f1), _f(f2)
    B, Cc, H, W = f1.shape
    off, _, _ = pyramid_layout(B, H, W, levels)
    pyr = np.zeros(off[-1], np.float32)
    lib().oracle_corr_pyramid_forward(_p(f1), _p(f2), _p(pyr), B, Cc, H, W, levels)
    return pyr


def corr_pyramid_backward(gpyr, f1, f2, levels=4):
    gpyr, f1, f2 = _f(gpyr), _f(f1), _f(f2)
    B, Cc, H, W = f1.shape
    g1, g2 = np.zeros_like(f1), np.zeros_like(f2)
    lib().oracle_corr_pyramid_backward(_p(gpyr), _p(f1), _p(f2), _p(g1), _p(g2), B, Cc, H, W, levels)
    return g1, g2


def corr_lookup_forward(pyr, coords, levels=4, radius=4):
    pyr, coords = _f(pyr), _f(coords)
    B, _, H, W = coords.shape
    D = 2 * radius + 1
    out = np.zeros((B, levels * D * D, H, W), np.float32)
    lib().oracle_corr_lookup_forward(_p(pyr), _p(coords), _p(out), B, H, W, levels, radius)
    return out


def corr_lookup_backward(gout, coords, levels=4, radius=4, gpyr=None):
    gout, coords = _f(gout), _f(coords)
    B, _, H, W = coords.shape
    off, _, _ = pyramid_layout(B, H, W, levels)
    if gpyr is None:
        gpyr = np.zeros(off[-1], np.float32)
    lib().oracle_corr_lookup_backward(_p(gout), _p(coords), _p(gpyr), B, H, W, levels, radius)
    return gpyr


# ---------------------------------------------------------------- PWC warp
def pwc_warp_forward(x, flow):
    x, flow = _f(x), _f(flow)
    B, Cc, H, W = x.shape
    out = np.zeros_like(x)
    lib().oracle_pwc_warp_forward(_p(x), _p(flow), _p(out), B, Cc, H, W)
    return out


def pwc_warp_backward(x, flow, gout):
    x, flow, gout = _f(x), _f(flow), _f(gout)
    B, Cc, H, W = x.shape
    gx, gf = np.zeros_like(x), np.zeros_like(flow)
    lib().oracle_pwc_warp_backward(_p(x), _p(flow), _p(gout), _p(gx), _p(gf), B, Cc, H, W)
    return gx, gf


# ---------------------------------------------------------------- PCFA objective (numpy)
def cov_transform(w, eps_box):
    """(1./2.) * 1./(1.-eps) * (tanh(w) + (1-eps))  — own_models.py:73-75, attack_PCFA.py:23-24.
    float32 arithmetic with the python scalars rounded to fp32, like torch does."""
    w = np.asarray(w, np.float32)
    c = np.float32(0.5 * 1.0 / (1.0 - eps_box))
    ome = np.float32(1.0 - eps_box)
    return c * (np.tanh(w) + ome)


def box_forward(var, image, mode, eps_box=0.0, scale=1.0, amax=None, amin=None):
    """Returns (net_in, delta, sum(delta^2)) for one image.  Modes as include/pcfa_b200.h:
    0 COV, 1 CLIP, 2 JOINT (per-pair shared delta), 3 UNIVERSAL (broadcast delta)."""
    image = np.asarray(image, np.float32)
    var = np.asarray(var, np.float32)
    if mode == 0:
        u = cov_transform(var, eps_box)
        x = np.clip(u, 0, 1); d = u - image
    elif mode == 1:
        x = np.clip(var, 0, 1); d = x - image
    elif mode == 2:
        x = np.clip(image + var, 0, 1)
        up = np.clip(var + amax, 0, 1) - amax                     # attack_PCFA.py:34
        d = np.clip(up + amin, 0, 1) - amin                       # attack_PCFA.py:35
    else:
        x = np.clip(image + var[None], 0, 1); d = var
    return (np.float32(scale) * x).astype(np.float32), d.astype(np.float32), float(np.sum(d.astype(np.float64) ** 2))


def loss_delta_constraint(pred, target, delta1, delta2, delta_bound, mu, f_type="aee"):
    """helper_functions/losses.py:200-230 in float64; returns (loss, sim, penalty_active, grad_pred)."""
    p = np.asarray(pred, np.float64); t = np.asarray(target, np.float64)
    if f_type == "aee":                                            # losses.py:20-27
        n = np.sqrt(((p - t) ** 2).sum(axis=1))
        sim = n.mean()
        with np.errstate(divide="ignore", invalid="ignore"):
            g = np.where(n[:, None] > 0, (p - t) / n[:, None], 0.0) / n.size
    elif f_type == "mse":                                          # losses.py:44
        sim = ((p - t) ** 2).mean(); g = 2 * (p - t) / p.size
    elif f_type == "cosim":                                        # losses.py:88 (sic)
        A, P, T = (p * t).sum(), (p * p).sum(), (t * t).sum()
        sim = 1 - A / np.sqrt(P) * np.sqrt(T)
        g = -np.sqrt(T) * (t / np.sqrt(P) - A * p / P ** 1.5)
    else:
        raise NotImplementedError(f_type)
    d1 = np.asarray(delta1, np.float64); d2 = np.asarray(delta2, np.float64)
    mean_sq = ((d1 ** 2).sum() + (d2 ** 2).sum()) / (d1.size + d2.size)          # losses.py:122-126
    excess = mean_sq - delta_bound ** 2
    loss = sim + mu * max(0.0, excess)                                           # losses.py:195-197,230
    return loss, sim, excess > 0, g
