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
    assert (a.net, a.dataset, a.steps, a.boxconstraint, a.batch_size) == ("SpyNet", "Kitti15", 20, "change_of_variables", 4)
    assert (a.delta_bound, a.mu, a.epochs, a.target, a.loss) == (0.005, -1, 25, "zero", "aee")
    assert not a.joint_perturbation and not a.universal_perturbation and a.output_folder == "experiment_data"
    a = create_parser("training", "pcfa").parse_args(["--net", "GMA", "--joint_perturbation", "--universal_perturbation",
                                                     "--target", "neg_flow", "--loss", "cosim", "--delta_bound", "0.01"])
    assert a.net == "GMA" and a.joint_perturbation and a.universal_perturbation and a.loss == "cosim"
    with pytest.raises(SystemExit):
        create_parser("training", "pcfa").parse_args(["--net", "LiteFlowNet"])
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
