"""Per-kernel device timing of the C-ABI entry points with CUDA events (on the launching stream), and
the algorithmic-bytes model used for the roofline fractions (DESIGN.md "Kernels and rooflines")."""
from __future__ import annotations

import torch

from . import _lib


def algorithmic_bytes(name: str, *, B: int, C: int, H: int, W: int, levels: int = 4, radius: int = 4,
                      img_numel: int = 0, flow_numel: int = 0):
    """Minimum HBM bytes one launch of `name` must move, or None if not modelled.
    H, W are the feature-map (1/8) sizes for the correlation kernels."""
    N = H * W
    pyr, h, w = 0, H, W
    for _ in range(levels):
        pyr += B * N * h * w
        h, w = h // 2, w // 2
    D, F = 2 * radius + 1, 2 * radius + 2
    if name.endswith("_occ"):                             # sparse-backward variants: same per-unit figure (SURVEY.md §8d counts the
        name = name[:-4]                                  # dense gradient pyramid; skipped blocks are bytes NOT moved)
    if name == "pcfa_corr_pyramid_forward":
        return 4 * (2 * B * C * N + pyr)
    if name == "pcfa_corr_pyramid_backward":
        return 4 * (pyr + 4 * B * C * N)
    if name.endswith("_cl"):                              # channels-last variants move the same bytes
        name = name[:-3]
    if name == "pcfa_corr_lookup_forward":
        return 4 * B * N * (levels * F * F + levels * D * D + 2)
    if name == "pcfa_corr_lookup_backward":
        return 4 * B * N * (levels * D * D + 2 + 2 * levels * F * F)
    if name == "pcfa_box_forward":
        return 4 * 3 * img_numel            # read var + image, write net_in
    if name == "pcfa_box_backward":
        return 4 * 4 * img_numel            # read var, image, grad_net_in; write grad_var
    if name == "pcfa_objective_loss":
        return 4 * 3 * flow_numel           # read flow + target, write grad_flow
    return None


def call_bytes(name: str, args):
    """Minimum HBM bytes of ONE call from its own arguments (entry points whose shape changes from call to call)."""
    if len(args) == 2 and args[0] == "bytes":            # caller-provided (_lib.hint_bytes)
        return args[1]
    if name == "pcfa_gru_gates_x_forward":               # (zr, P, h, m, z, r, rhm, C, Cm, npix): r is not counted
        C, Cm, n = args[7], args[8], args[9]
        return 4 * n * (7 * C + 2 * Cm)
    if name == "pcfa_gru_gates_x_backward":              # (z, r, h, gz, grhm, gzr, gh, C, Cm, npix)
        return 4 * args[9] * 8 * args[7]
    if name == "pcfa_gru_blend_x_forward":               # (z, q_pre, P, h, m, q, hn, hm, C, Cm, npix)
        C, Cm, n = args[8], args[9], args[10]
        return 4 * n * (6 * C + (C + 2 * Cm if args[7] is not None else 0))
    if name == "pcfa_gru_blend_x_backward":              # (z, q, h, ghn, ghm, gz, gq, gh, C, Cm, npix)
        return 4 * args[10] * 8 * args[8]
    if name == "pcfa_gru_gates_forward":                 # (zr, h, z, r, rh, B, n, ...): read zr (2), h; write z, rh
        return 4 * 5 * args[5] * args[6]
    if name == "pcfa_gru_gates_backward":                # (z, r, h, gz, grh, gzr, gh, B, n, ...): read 5, write 3
        return 4 * 8 * args[7] * args[8]
    if name == "pcfa_gru_blend_forward":                 # (z, q_pre, h, q, h_new, numel, ...): read 3, write 2
        return 4 * 5 * args[5]
    if name == "pcfa_gru_blend_backward":                # (z, q, h, ghn, gz, gq, gh, numel, ...): read 4, write 3
        return 4 * 7 * args[7]
    if name == "pcfa_gru_gates_x_backward_acc":          # (z, r, h, gz, grhm, gzr, gh, acc, mode, C, Cm, npix): + RMW of the accumulator
        return 4 * args[11] * (8 + (4 if args[8] == 2 else 2 if args[8] == 1 else 0)) * args[9]
    if name == "pcfa_gru_blend_x_backward_acc":          # (z, q, h, ga, gb, ghm, gz, gq, gh, acc, mode, C, Cm, npix)
        n, C = args[13], args[11]
        return 4 * n * C * (3 + 1 + (1 if args[4] is not None else 0) + (1 if args[5] is not None else 0) + 3 + (2 if args[10] == 2 else 1 if args[10] == 1 else 0))
    if name == "pcfa_gru_step_combine":                  # (gh_a, gh_b, c0..c3, gh, gm, C, Cm, npix): read 2C + C + 4Cm, write C + Cm
        C, Cm, n = args[8], args[9], args[10]
        return 4 * n * (4 * C + 5 * Cm)
    if name == "pcfa_bias_act_forward":                  # (x, bias, n, C, inner, relu, slope, dtype): in place
        return args[2] * 2 * (4 if args[7] == 0 else 2)
    if name == "pcfa_relu_mask_backward":                # (y, gy, gx, n, slope, dtype)
        return args[3] * 3 * (4 if args[5] == 0 else 2)
    if name == "pcfa_convex_upsample_forward":           # (flow, mask, up, N, H, W, ...): mask 576 + flow 2 in, 128 out per coarse pixel
        return 4 * args[3] * args[4] * args[5] * (576 + 2 + 128)
    if name == "pcfa_convex_upsample_backward":          # (flow, mask, gup, gflow, gmask, ws, wsb, N, H, W, ...): mask, gup in; gmask, gflow out
        return 4 * args[7] * args[8] * args[9] * (576 + 128 + 576 + 2 + 2 + 36)
    if name == "pcfa_cat_channels_last_pad":             # (inputs, channels, n, out, npix, out_channels)
        return 2 * 4 * args[4] * args[5]
    if name == "pcfa_softmax_rows_f16_forward":          # (sim, attn, rows, cols)
        return 2 * 2 * args[2] * args[3]
    if name == "pcfa_softmax_rows_f16_backward":         # (attn, gattn, gsim, rows, cols)
        return 3 * 2 * args[3] * args[4]
    if name == "pcfa_instnorm_forward":                  # (x, y, stats, ws, B, C, H, W, ...): read x, write y
        return 4 * 2 * args[4] * args[5] * args[6] * args[7]
    if name == "pcfa_instnorm_backward":                 # (x, gy, stats, gx, ws, B, C, H, W, ...): read x, gy, write gx
        return 4 * 3 * args[5] * args[6] * args[7] * args[8]
    return None


def kernel_table(step_fn, n_steps: int, *, B: int, C: int, H: int, W: int, iters: int, peak_gbs: float,
                 img_numel: int = 0, flow_numel: int = 0):
    """Run `step_fn` n_steps times with event bracketing and return one row per entry point."""
    store: dict = {}
    step_fn()
    torch.cuda.synchronize()
    _lib.set_profile(store)
    try:
        for _ in range(n_steps):
            step_fn()
        torch.cuda.synchronize()
    finally:
        _lib.set_profile(None)
    overhead = _lib.event_overhead_us()
    rows = []
    for name, evs in store.items():
        us = [max(a.elapsed_time(b) * 1e3 - overhead, 0.1) for a, b, _ in evs]
        per_step = len(us) / n_steps
        avg = sum(us) / len(us)
        ab = algorithmic_bytes(name, B=B, C=C, H=H, W=W, img_numel=img_numel, flow_numel=flow_numel)
        if ab is None:                                   # shape varies per call: average the per-call minimum
            per_call = [call_bytes(name, args) for _, _, args in evs]
            if per_call and all(v is not None for v in per_call):
                ab = int(sum(per_call) / len(per_call))
        row = {"name": name, "launches_per_step": per_step, "avg_us": round(avg, 2),
               "total_us_per_step": round(avg * per_step, 1), "algorithmic_bytes": ab}
        if ab:
            row["achieved_gbs"] = round(ab / (avg * 1e-6) / 1e9, 1)
            row["frac_of_hbm_peak"] = round(row["achieved_gbs"] / peak_gbs, 4)
        else:
            row["achieved_gbs"] = 0.0
        rows.append(row)
    rows.sort(key=lambda r: -r["total_us_per_step"])
    return rows
