"""Network definitions of this package reproduce the reference networks' outputs (CPU, oracle ops
injected) — pins state-dict compatibility and the forward/backward graph, independent of the GPU."""
import numpy as np
import torch

from conftest import assert_close
from oracle import torch_ref as TR


def test_raft_definition_matches_reference_forward_backward(golden):
    from pcfa_b200.adapter import build_network
    from pcfa_b200.networks.weights import synthetic_pair
    z = golden("raft_e2e")
    net = build_network("RAFT", device="cpu", seed=0, ops=TR)
    i1, i2 = synthetic_pair(0, 128, 160)
    i1.requires_grad_(True); i2.requires_grad_(True)
    lo, up = net(i1, i2, iters=12, test_mode=True)
    assert_close(up.detach().numpy(), z["flow_up"], rtol=1e-4, atol_rms=1e-4, what="flow_up")
    assert_close(lo.detach().numpy(), z["flow_lo"], rtol=1e-4, atol_rms=1e-4, what="flow_lo")
    (up * torch.from_numpy(z["gout"])).sum().backward()
    assert_close(i1.grad.numpy(), z["g_img1"], rtol=1e-3, atol_rms=1e-3, what="g img1")
    assert_close(i2.grad.numpy(), z["g_img2"], rtol=1e-3, atol_rms=1e-3, what="g img2")


def test_raft_training_mode_returns_all_predictions():
    from pcfa_b200.adapter import build_network
    net = build_network("RAFT", device="cpu", seed=1, ops=TR)
    x = torch.rand(1, 3, 128, 136) * 255
    preds = net(x, x.flip(-1), iters=3, test_mode=False)
    assert len(preds) == 3 and preds[0].shape == (1, 2, 128, 136)


def test_raft_accepts_dataparallel_checkpoint_keys():
    from pcfa_b200.networks.raft import RAFT
    a = RAFT(corr_block=TR.CorrBlock)
    sd = {"module." + k: v for k, v in a.state_dict().items()}
    b = RAFT(corr_block=TR.CorrBlock)
    b.load_state_dict(sd)


def test_gma_definition_matches_reference(golden):
    from pcfa_b200.adapter import build_network
    from pcfa_b200.networks.weights import synthetic_pair
    z = golden("networks")
    net = build_network("GMA", device="cpu", seed=0, ops=TR, gain=0.5)
    i1, i2 = synthetic_pair(3, 128, 136)
    with torch.no_grad():
        up = net(i1, i2, iters=6, test_mode=True)[1]
    assert_close(up.numpy(), z["gma_flow"], rtol=1e-4, atol_rms=1e-4, what="GMA flow")


def test_pwcnet_definition_matches_reference_forward_backward(golden):
    from pcfa_b200.adapter import build_network
    from pcfa_b200.networks.weights import synthetic_pair
    z = golden("networks")
    net = build_network("PWCNet", device="cpu", seed=0, ops=TR)
    i1, i2 = synthetic_pair(4, 128, 192)
    a = (i1 / 255.).requires_grad_(True)
    flow = net(a, i2 / 255.)
    assert_close(flow.detach().numpy(), z["pwc_flow"], rtol=1e-4, atol_rms=1e-4, what="PWCNet flow")
    (flow * torch.from_numpy(z["pwc_gout"])).sum().backward()
    assert_close(a.grad.numpy(), z["pwc_g_img1"], rtol=1e-3, atol_rms=1e-3, what="PWCNet grad")


def test_flownet2_definition_matches_reference(golden):
    from pcfa_b200.adapter import build_network, compute_flow
    from pcfa_b200.networks.weights import synthetic_pair
    z = golden("networks")
    net = build_network("FlowNet2", device="cpu", seed=0, ops=TR, gain=0.7)
    i1, i2 = synthetic_pair(5, 64, 128)
    with torch.no_grad():
        flow = compute_flow(net, "FlowNet2", i1, i2)
    assert_close(flow.numpy(), z["fn2_flow"], rtol=1e-4, atol_rms=1e-4, what="FlowNet2 flow")
