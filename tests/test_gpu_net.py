"""Network-level parity on the GPU: RAFT with the CUDA CorrBlock against the reference's outputs."""
import numpy as np
import pytest
import torch

from conftest import assert_close

pytestmark = pytest.mark.gpu


@pytest.fixture()
def fp32_convs():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("impl", [1, 0])
def test_raft_gpu_matches_reference_flow_and_gradients(golden, fp32_convs, impl, monkeypatch):
    from pcfa_b200.adapter import build_network
    from pcfa_b200.networks.weights import synthetic_pair
    monkeypatch.setenv("PCFA_CORR_IMPL", str(impl))
    z = golden("raft_e2e")
    net = build_network("RAFT", device="cuda", seed=0)
    i1, i2 = synthetic_pair(0, 128, 160)
    i1, i2 = i1.cuda().requires_grad_(True), i2.cuda().requires_grad_(True)
    lo, up = net(i1, i2, iters=12, test_mode=True)
    # 12 recurrent iterations amplify fp32 reordering noise; rtol 1e-3 on flows is BASELINE's bar
    assert_close(up.detach().cpu().numpy(), z["flow_up"], rtol=1e-3, atol_rms=2e-3, what="flow_up")
    (up * torch.from_numpy(z["gout"]).cuda()).sum().backward()
    assert_close(i1.grad.cpu().numpy(), z["g_img1"], rtol=1e-2, atol_rms=1e-2, what="g img1")
    assert_close(i2.grad.cpu().numpy(), z["g_img2"], rtol=1e-2, atol_rms=1e-2, what="g img2")


def test_fused_closure_matches_torch_autograd_composition():
    """FusedObjective.evaluate == autograd through scaled_input → net → unpad → loss_delta_constraint."""
    from pcfa_b200 import objective as J
    from pcfa_b200.adapter import build_network, preprocess_img
    from pcfa_b200.networks.weights import synthetic_pair
    net = build_network("RAFT", device="cuda", seed=0)
    i1, i2 = synthetic_pair(1, 132, 170)
    padder, (a, b) = preprocess_img("RAFT", i1.cuda() / 255.0, i2.cuda() / 255.0)
    a, b = a.contiguous(), b.contiguous()
    target = torch.zeros(1, 2, 132, 170, device="cuda")
    eps = 1e-7
    w1 = (torch.atanh(2 * (1 - eps) * a - (1 - eps)) + 0.02 * torch.randn_like(a)).requires_grad_(True)
    w2 = (torch.atanh(2 * (1 - eps) * b - (1 - eps)) + 0.02 * torch.randn_like(b)).requires_grad_(True)
    fwd = lambda x, y: net(x, y, iters=3, test_mode=True)[1]
    fo = J.FusedObjective(fwd, a, b, target, mode=J.BOX_COV, joint=False, pad=padder.top_left, eps_box=eps,
                          scale=255.0, delta_bound=0.005, mu=5e5, loss="aee")
    loss, g1, g2 = fo.evaluate(w1.detach(), w2.detach())
    x1 = J.scaled_input(w1, var_change=True, eps_box=eps, make_unit_input=True)
    x2 = J.scaled_input(w2, var_change=True, eps_box=eps, make_unit_input=True)
    flow = padder.unpad(fwd(x1, x2))
    d1, d2 = J.extract_deltas(w1, w2, a, b, "change_of_variables", eps_box=eps)
    ref = J.loss_delta_constraint(flow, target, d1, d2, None, delta_bound=0.005, mu=5e5, f_type="aee")
    ref.backward()
    np.testing.assert_allclose(float(loss), float(ref), rtol=1e-4)
    assert_close(g1.cpu().numpy(), w1.grad.cpu().numpy(), rtol=1e-3, atol_rms=1e-3, what="gw1")
    assert_close(g2.cpu().numpy(), w2.grad.cpu().numpy(), rtol=1e-3, atol_rms=1e-3, what="gw2")


def test_closure_is_cuda_graph_capturable():
    from pcfa_b200 import objective as J
    from pcfa_b200.adapter import build_network, preprocess_img
    from pcfa_b200.networks.weights import synthetic_pair
    net = build_network("RAFT", device="cuda", seed=0)
    i1, i2 = synthetic_pair(2, 128, 160)
    padder, (a, b) = preprocess_img("RAFT", i1.cuda() / 255.0, i2.cuda() / 255.0)
    target = torch.zeros(1, 2, 128, 160, device="cuda")
    fo = J.FusedObjective(lambda x, y: net(x, y, iters=2, test_mode=True)[1], a.contiguous(), b.contiguous(), target,
                          mode=J.BOX_CLIP, joint=False, pad=padder.top_left, eps_box=0.0, scale=255.0,
                          delta_bound=0.005, mu=5e5, loss="aee")
    w1, w2 = fo.image1.clone(), fo.image2.clone()
    g1, g2 = torch.empty_like(w1), torch.empty_like(w2)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fo.evaluate(w1, w2, g1, g2)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    eager = (float(fo.terms[0]), g1.clone(), g2.clone())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        fo.evaluate(w1, w2, g1, g2)
    g1.zero_(); g2.zero_()
    graph.replay()
    torch.cuda.synchronize()
    np.testing.assert_allclose(float(fo.terms[0]), eager[0], rtol=1e-5)
    assert_close(g1.cpu().numpy(), eager[1].cpu().numpy(), rtol=1e-4, atol_rms=1e-4, what="graph g1")
    w1.add_(0.01)                     # replays pick up new variable values (static buffers)
    graph.replay()
    torch.cuda.synchronize()
    assert abs(float(fo.terms[0]) - eager[0]) > 0
