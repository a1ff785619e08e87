"""Forward-only evaluation of stored perturbations (row f-3 of SURVEY.md section 8): host-side mirror of
evaluate_PCFA.py:21-79 (discovery of `NNNNN_delta{1,2}_eE.npy` files, re-padding between network families) and
:86-299 (perturbed vs unperturbed prediction, AEE(f_adv, f_init), L2 of the perturbation).  The forward pass is the
same fused path the attack uses (box kernel in universal mode -> network -> loss kernel), no gradients."""
from __future__ import annotations

import math
import os
import re

import numpy as np
import torch

from . import objective as J
from .adapter import model_takes_unit_input, preprocess_img

RAFT_PADDED = ("RAFT", "GMA")
FLOWNET_PADDED = ("PWCNet", "SpyNet", "FlowNet2")


def extract_epoch_patchlist(path: str):
    """(epochs, delta1 paths, delta2 paths).  `path` is one .npy file (one epoch) or a run folder whose `patches/`
    holds `BBBBB_delta1_eEE.npy` / `BBBBB_delta2_eEE.npy` as written by attack_PCFA.py --universal_perturbation."""
    if os.path.isfile(path):
        if os.path.splitext(path)[1] != ".npy":
            raise ValueError("Invalid extension %s for perturbation file, please use a .npy file instead of %s"
                             % (os.path.splitext(path)[1], path))
        return 1, [path], []
    base = os.path.join(path, "patches")
    names = sorted(os.listdir(base))
    d1 = [os.path.join(base, n) for n in names if re.fullmatch(r"[0-9]{5}_delta1_e[0-9]*\.npy", n)]
    d2 = [os.path.join(base, n) for n in names if re.fullmatch(r"[0-9]{5}_delta2_e[0-9]*\.npy", n)]
    if not d1:
        raise FileNotFoundError("no NNNNN_delta1_eE.npy files under %s" % base)
    last_epoch = int(re.search(r"_e([0-9]*)\.npy$", d1[-1]).group(1))
    return last_epoch + 1, d1, d2                       # epochs are counted from 0


def convert_perturbationsizes(delta: torch.Tensor, image_hw, network_training: str, network_eval: str) -> torch.Tensor:
    """Re-pad a perturbation ([3, Hp, Wp], image units [0,1]) trained for one padding family (RAFT/GMA: multiples of 8,
    centred; PWCNet/SpyNet/FlowNet2: multiples of 64) for a network of the other family.  `image_hw` is the unpadded
    image size of the dataset.  Unit-input networks divide by 255 in their pre-processing, which is undone here."""
    same = ((network_training in FLOWNET_PADDED and network_eval in FLOWNET_PADDED)
            or (network_training in RAFT_PADDED and network_eval in RAFT_PADDED))
    if same:
        return delta
    probe = torch.zeros(1, 3, *image_hw, dtype=delta.dtype, device=delta.device)
    padder_train, _ = preprocess_img(network_training, probe)
    unpadded = padder_train.unpad(delta).unsqueeze(0)
    _, (repadded,) = preprocess_img(network_eval, unpadded.detach().clone())
    if model_takes_unit_input(network_eval):
        repadded = repadded * 255.
    return repadded[0] if repadded.dim() == 4 and delta.dim() == 3 else repadded


def l2_metrics(delta1: torch.Tensor, delta2: torch.Tensor):
    """sqrt(mean(delta^2)) per perturbation and over both (losses.py:91-126)."""
    s1, s2 = float(delta1.pow(2).sum()), float(delta2.pow(2).sum())
    n1, n2 = delta1.numel(), delta2.numel()
    return math.sqrt(s1 / n1), math.sqrt(s2 / n2), math.sqrt((s1 + s2) / (n1 + n2))


@torch.no_grad()
def evaluate_perturbation(model, net_name: str, delta1: torch.Tensor, delta2: torch.Tensor | None, batches, *,
                          joint: bool, iters=None, boxconstraint: str = "clipping", eps_box: float = 1e-7):
    """batches: iterable of (image1, image2) in [0,255], [b,3,H,W] on the device.  delta2=None or joint=True applies
    delta1 to both frames.  Returns dict(aee_adv_pred=mean AEE(f_adv, f_init), images=count).

    boxconstraint="change_of_variables" reproduces the reference's default evaluation (evaluate_PCFA.py:151-154 builds
    the ScaledInputModel with variable_change=True): image+delta goes through the tanh transform as if it were the
    w-variable (own_models.py:62-85), for the unperturbed prediction as well.  That composition is forward-only and
    outside the attack's hot path, so it runs as four element-wise torch ops in front of the network."""
    from .attack import _net_forward, avg_epe
    unit = model_takes_unit_input(net_name)
    fwd = _net_forward(model, net_name, iters)
    total, count = 0.0, 0
    d1 = delta1.contiguous()
    d2 = None if (joint or delta2 is None) else delta2.contiguous()
    for image1, image2 in batches:
        if not unit:
            image1, image2 = image1 / 255., image2 / 255.
        padder, (image1, image2) = preprocess_img(net_name, image1, image2)
        image1, image2 = image1.contiguous(), image2.contiguous()
        H, W = padder.get_dimensions()
        b = image1.shape[0]
        fo = J.FusedObjective(fwd, image1, image2, torch.zeros(b, 2, H, W, device=image1.device), mode=J.BOX_UNIVERSAL,
                              joint=d2 is None, pad=padder.top_left, eps_box=0.0, scale=1.0 if unit else 255.0,
                              delta_bound=1.0, mu=0.0, loss="aee")
        if boxconstraint == "change_of_variables":
            sc = 1.0 if unit else 255.0

            def cov(img, d):
                x = img if d is None else img + d
                x = 0.5 / (1.0 - eps_box) * (torch.tanh(x) + (1.0 - eps_box))
                return (torch.clamp(x, 0.0, 1.0) * sc).contiguous()
            dd2 = d1 if d2 is None else d2
            flow_init = padder.unpad(fwd(cov(image1, None), cov(image2, None))).contiguous().clone()
            flow_adv = padder.unpad(fwd(cov(image1, d1), cov(image2, dd2)))
        else:
            zero = torch.zeros_like(d1)
            flow_init = padder.unpad(fo.predict(zero, None if d2 is None else zero)).contiguous().clone()
            flow_adv = padder.unpad(fo.predict(d1, d2))
        for i in range(b):
            total += float(avg_epe(flow_adv[i:i + 1], flow_init[i:i + 1]))
        count += b
    return dict(aee_adv_pred=total / max(count, 1), images=count)


def load_delta(path: str, device) -> torch.Tensor:
    return torch.from_numpy(np.load(path)).float().to(device)
