"""Host-side logic that needs no GPU: CLI surface, mu heuristic, box-mode rules, padding arithmetic,
targets, and the multi-rank reduction protocol of the universal mode (gloo, world_size 2)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pcfa_b200 import objective as J
from pcfa_b200.adapter import InputPadder, model_takes_unit_input, preprocess_img
from pcfa_b200.attack import avg_epe, get_target, resolve_mu
from pcfa_b200.parsing import create_parser


def test_cli_flags_and_defaults_match_reference():
    a = create_parser("training", "pcfa").parse_args([])
    # the reference's --net default is SpyNet (parsing_file.py:12), which has no cost volume and is not part of this build:
    # the default here is RAFT and SpyNet is rejected at argument parsing instead of crashing later
    assert (a.net, a.dataset, a.steps, a.boxconstraint, a.batch_size) == ("RAFT", "Kitti15", 20, "change_of_variables", 4)
    assert (a.delta_bound, a.mu, a.epochs, a.target, a.loss) == (0.005, -1, 25, "zero", "aee")
    assert not a.joint_perturbation and not a.universal_perturbation and a.output_folder == "experiment_data"
    a = create_parser("training", "pcfa").parse_args(["--net", "GMA", "--joint_perturbation", "--universal_perturbation",
                                                     "--target", "neg_flow", "--loss", "cosim", "--delta_bound", "0.01"])
    assert a.net == "GMA" and a.joint_perturbation and a.universal_perturbation and a.loss == "cosim"
    with pytest.raises(SystemExit):
        create_parser("training", "pcfa").parse_args(["--net", "LiteFlowNet"])
    with pytest.raises(SystemExit):
        create_parser("training", "pcfa").parse_args(["--net", "SpyNet"])
    with pytest.raises(ValueError):
        create_parser("nope", "pcfa")


def test_mu_heuristic_and_box_modes():
    assert resolve_mu(-1., 0.005, "zero") == 2500. / 0.005              # attack_PCFA.py:578-583
    assert resolve_mu(-1., 0.005, "neg_flow") == 1.5 * 2500. / 0.005
    assert resolve_mu(7., 0.005, "zero") == 7.
    assert J.box_mode("change_of_variables") == J.BOX_COV and J.box_mode("clipping") == J.BOX_CLIP
    assert J.box_mode("clipping", joint=True) == J.BOX_JOINT
    assert J.box_mode("change_of_variables", universal=True) == J.BOX_UNIVERSAL   # universal silently clips (:319)
    with pytest.raises(ValueError, match="not defined"):
        J.box_mode("change_of_variables", joint=True)                    # attack_PCFA.py:91-92


def test_padding_matches_reference_sizes():
    p = InputPadder((1, 3, 436, 1024))                                   # Sintel, divisor 8
    assert p._pad == [0, 0, 2, 2] and p.top_left == (2, 0)
    x = torch.arange(436 * 1024, dtype=torch.float32).view(1, 1, 436, 1024)
    (y,) = p.pad(x)
    assert y.shape == (1, 1, 440, 1024) and torch.equal(p.unpad(y), x) and torch.equal(y[0, 0, 0], x[0, 0, 0])
    p = InputPadder((1, 3, 375, 1242), divisor=64)                       # KITTI for PWCNet / FlowNet2
    (y,) = p.pad(torch.zeros(1, 3, 375, 1242))
    assert y.shape == (1, 3, 384, 1280)
    _, (a,) = preprocess_img("PWCNet", torch.full((1, 3, 375, 1242), 255.))
    assert a.shape == (1, 3, 384, 1280) and float(a.max()) == 1.0        # unit-input nets are scaled here
    assert model_takes_unit_input("PWCNet") and not model_takes_unit_input("RAFT")


def test_targets_and_epe():
    f = torch.randn(2, 2, 5, 7)
    assert torch.equal(get_target("zero", f), torch.zeros_like(f))
    assert torch.equal(get_target("neg_flow", f), -f)
    with pytest.raises(ValueError):
        get_target("sideways", f)
    ref = torch.sum((f - 1) ** 2, dim=1).sqrt().mean()
    assert float(avg_epe(f, torch.ones_like(f))) == pytest.approx(float(ref))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close(); return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pcfa_b200.dist import pack_reduce_unpack, shard_indices
    g = torch.Generator().manual_seed(100 + rank)
    g1, g2 = torch.randn(3, 4, 5, generator=g), torch.randn(3, 4, 5, generator=g)
    loss = torch.tensor(float(rank + 1))
    flat = torch.zeros(2 * 60 + 1)
    l = pack_reduce_unpack(flat, loss, g1, g2)
    out[rank] = (float(l), g1.clone(), g2.clone(), shard_indices(10, rank, world))
    dist.destroy_process_group()


def test_universal_reduction_protocol_gloo_world2():
    """Every rank must end up with the mean gradient and mean loss (identical L-BFGS decisions)."""
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    exp1 = sum(torch.randn(3, 4, 5, generator=torch.Generator().manual_seed(100 + r)) for r in range(2)) / 2
    for r in range(2):
        l, g1, g2, shard = out[r]
        assert l == pytest.approx(1.5)
        assert torch.allclose(g1, exp1)
        assert torch.equal(out[0][2], out[1][2])
    assert out[0][3] == [0, 2, 4, 6, 8] and out[1][3] == [1, 3, 5, 7, 9]


# ----------------------------------------------------------------------------------- evaluation path (f-3)
def test_extract_epoch_patchlist_and_repadding(tmp_path):
    """evaluate_PCFA.py:21-79: discovery of the per-epoch perturbation files and re-padding between the RAFT family
    (multiples of 8, centred) and the PWCNet/FlowNet2 family (multiples of 64)."""
    import numpy as np
    import torch
    from pcfa_b200.adapter import InputPadder
    from pcfa_b200.evaluate import convert_perturbationsizes, extract_epoch_patchlist, l2_metrics
    run = tmp_path / "run"
    (run / "patches").mkdir(parents=True)
    for e in range(3):
        np.save(run / "patches" / ("%05d_delta1_e%d.npy" % (e, e)), np.full((3, 4, 4), e, np.float32))
        np.save(run / "patches" / ("%05d_delta2_e%d.npy" % (e, e)), np.zeros((3, 4, 4), np.float32))
    np.save(run / "patches" / "00000_image1.npy", np.zeros(1))                      # ignored
    epochs, d1, d2 = extract_epoch_patchlist(str(run))
    assert epochs == 3 and len(d1) == 3 and len(d2) == 3 and d1[2].endswith("00002_delta1_e2.npy")
    assert extract_epoch_patchlist(d1[0]) == (1, [d1[0]], [])
    with pytest.raises(ValueError):
        (tmp_path / "x.txt").write_text("x")
        extract_epoch_patchlist(str(tmp_path / "x.txt"))
    # Sintel: RAFT pads 436 -> 440 (2 + 2 rows); PWCNet/FlowNet2 pad 436 -> 448 (6 + 6 rows)
    H, W = 436, 1024
    g = torch.Generator().manual_seed(0)
    delta = torch.rand(3, 440, 1024, generator=g) * 0.01
    same = convert_perturbationsizes(delta, (H, W), "RAFT", "GMA")
    assert same is delta
    for net in ("FlowNet2", "PWCNet"):
        out = convert_perturbationsizes(delta, (H, W), "RAFT", net)
        assert tuple(out.shape) == (3, 448, 1024)
        core = InputPadder((1, 3, H, W), divisor=64).unpad(out)
        torch.testing.assert_close(core, delta[:, 2:438], rtol=1e-6, atol=1e-9)      # unit-input /255 is undone
    back = convert_perturbationsizes(convert_perturbationsizes(delta, (H, W), "RAFT", "FlowNet2"), (H, W), "FlowNet2", "RAFT")
    torch.testing.assert_close(back[:, 2:438], delta[:, 2:438])
    l1, l2, l12 = l2_metrics(torch.full((3, 2, 2), 0.1), torch.zeros(3, 2, 2))
    assert l1 == pytest.approx(0.1) and l2 == 0.0 and l12 == pytest.approx(0.1 / 2 ** 0.5)


# ----------------------------------------------------------------------------------- encoder / GRU weight algebra (f-4)
def test_bn_fold_and_gru_weight_split_are_exact_algebra():
    """The frozen-weight rewrites the GPU path relies on, checked on CPU with plain torch ops:
    eval batch norm folded into the convolution, convz|convr as one convolution, and the context-feature share of
    the GRU convolutions hoisted out of the loop (conv([h|inp|motion]) = conv_hm([h|motion]) + conv_inp(inp) + b)."""
    import torch
    import torch.nn as nn
    import torch.nn.functional as F
    from pcfa_b200.networks.raft import SepConvGRU, _bn_folded, _zr_weights
    torch.manual_seed(0)
    conv, bn = nn.Conv2d(5, 7, 3, padding=1, stride=2), nn.BatchNorm2d(7)
    bn.running_mean.uniform_(-1, 1); bn.running_var.uniform_(0.5, 2); bn.weight.data.uniform_(0.5, 1.5); bn.bias.data.uniform_(-1, 1)
    bn.eval()
    x = torch.randn(2, 5, 9, 11)
    w, b = _bn_folded(conv, bn)
    torch.testing.assert_close(F.conv2d(x, w, b, conv.stride, conv.padding), bn(conv(x)), rtol=1e-5, atol=1e-5)
    with torch.no_grad():
        conv.weight.mul_(2.0)                                                   # in-place update (version counter) invalidates the cache
    w2, _ = _bn_folded(conv, bn)
    assert not torch.equal(w, w2)

    gru = SepConvGRU(hidden_dim=8, input_dim=8 + 12).eval()                     # x = [inp (8) | motion (12)]
    h, inp, motion = torch.randn(1, 8, 6, 7), torch.randn(1, 8, 6, 7), torch.randn(1, 12, 6, 7)
    wz, bz = _zr_weights(gru.convz1, gru.convr1)
    hx = torch.cat([h, inp, motion], 1)
    zr = F.conv2d(hx, wz, bz, 1, gru.convz1.padding)
    torch.testing.assert_close(zr[:, :8], gru.convz1(hx)); torch.testing.assert_close(zr[:, 8:], gru.convr1(hx))
    hoist = gru.hoisted(inp)
    (wzr1, pzr1, wq1, pq1, pad1), (wzr2, pzr2, wq2, pq2, pad2) = hoist
    hm = torch.cat([h, motion], 1)
    torch.testing.assert_close(F.conv2d(hm, wzr1, None, 1, pad1) + pzr1, zr, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(F.conv2d(hm, wq1, None, 1, pad1) + pq1, gru.convq1(hx), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(F.conv2d(hm, wzr2, None, 1, pad2) + pzr2,
                               torch.cat([gru.convz2(hx), gru.convr2(hx)], 1), rtol=1e-5, atol=1e-5)
    # the whole hoisted step equals the reference composition (models/raft/update.py:33-60) with torch ops
    def step(h, cz, cr, cq):
        x = torch.cat([inp, motion], 1)
        hx = torch.cat([h, x], 1)
        z, r = torch.sigmoid(cz(hx)), torch.sigmoid(cr(hx))
        q = torch.tanh(cq(torch.cat([r * h, x], 1)))
        return (1 - z) * h + z * q
    ref = step(step(h, gru.convz1, gru.convr1, gru.convq1), gru.convz2, gru.convr2, gru.convq2)
    def step_x(h, wzr, pzr, wq, pq, pad):
        zr = F.conv2d(torch.cat([h, motion], 1), wzr, None, 1, pad) + pzr
        z, r = torch.sigmoid(zr[:, :8]), torch.sigmoid(zr[:, 8:])
        q = torch.tanh(F.conv2d(torch.cat([r * h, motion], 1), wq, None, 1, pad) + pq)
        return (1 - z) * h + z * q
    out = step_x(step_x(h, *hoist[0]), *hoist[1])
    torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(gru(h, torch.cat([inp, motion], 1)), ref, rtol=1e-6, atol=1e-6)   # CPU module path


def test_evaluate_helpers_match_reference_outputs(golden, tmp_path):
    """f-3: convert_perturbationsizes and extract_epoch_patchlist against the outputs of the reference's own functions
    (evaluate_PCFA.py:21-79, executed in place by oracle/make_golden.py::golden_evaluate): re-padding between the
    RAFT/GMA (multiples of 8, centred) and PWCNet/FlowNet2 (multiples of 64) families incl. the x255 of unit-input
    networks, and epoch / file discovery in a folder written by attack_PCFA.py --universal_perturbation."""
    import json
    from pcfa_b200.evaluate import convert_perturbationsizes, extract_epoch_patchlist
    z = golden("evaluate")
    H, W = (int(v) for v in z["hw"])
    for key in [k for k in z.files if k.endswith("_delta")]:
        tr, ev = key.split("_")[:2]
        delta = torch.from_numpy(z[key])
        want = z[f"{tr}_{ev}_out"]
        got = convert_perturbationsizes(delta, (H, W), tr, ev).numpy()
        assert got.shape == want.shape[-3:], (key, got.shape, want.shape)     # the reference keeps a leading batch dim of 1
        np.testing.assert_array_equal(got, want.reshape(got.shape), err_msg=key)   # padding + exact scaling: bit-exact
    names = json.loads(bytes(z["patch_names"]).decode())
    (tmp_path / "patches").mkdir()
    for n in names:
        np.save(tmp_path / "patches" / n, np.zeros(1, np.float32))
    want = json.loads(bytes(z["patch_result"]).decode())
    epochs, d1, d2 = extract_epoch_patchlist(str(tmp_path))
    import os
    assert epochs == want["epochs"]
    assert [os.path.basename(p) for p in d1] == want["d1"] and [os.path.basename(p) for p in d2] == want["d2"]
