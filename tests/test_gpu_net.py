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


def _cos(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b)))


def _rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-300))


def _run_raft(ops, iters, gout):
    from pcfa_b200.adapter import build_network
    from pcfa_b200.networks.weights import synthetic_pair
    net = build_network("RAFT", device="cuda", seed=0, ops=ops)
    i1, i2 = synthetic_pair(0, 128, 160)
    i1, i2 = i1.cuda().requires_grad_(True), i2.cuda().requires_grad_(True)
    lo, up = net(i1, i2, iters=iters, test_mode=True)
    (up * gout).sum().backward()
    return up.detach().cpu().numpy(), i1.grad.cpu().numpy(), i2.grad.cpu().numpy()


@pytest.mark.parametrize("impl", [1, 0])
def test_raft_gpu_matches_reference_flow_and_gradients(golden, fp32_convs, impl, monkeypatch):
    """(1) 12-iteration flows against the reference's CPU outputs at BASELINE's rtol 1e-3;
    (2) flows and image gradients against the same network on the same GPU with the reference's
    torch-op CorrBlock (oracle/torch_ref.py), which isolates our kernels from cuDNN-vs-CPU convolution
    differences.  The random-weight recurrence amplifies 1e-6 forward differences in the gradient by
    ~10x per two iterations (measured: 2e-5 at 1 iteration, 1e-2 at 12, with bit-identical reruns),
    so the strict gradient bar is applied at 2 iterations and a direction bar at 12."""
    from oracle import torch_ref as TR
    monkeypatch.setenv("PCFA_CORR_IMPL", str(impl))
    z = golden("raft_e2e")
    gout = torch.from_numpy(z["gout"]).cuda()
    ours, ref = _run_raft(None, 12, gout), _run_raft(TR, 12, gout)
    assert_close(ours[0], z["flow_up"], rtol=1e-3, atol_rms=2e-3, what="flow_up vs reference (CPU)")
    assert_close(ours[0], ref[0], rtol=1e-3, atol_rms=1e-3, what="flow_up vs torch-op CorrBlock (GPU)")
    assert _rel_l2(ours[1], ref[1]) < 5e-2 and _rel_l2(ours[2], ref[2]) < 5e-2
    assert _cos(ours[1], z["g_img1"]) > 0.99 and _cos(ours[2], z["g_img2"]) > 0.99
    ours, ref = _run_raft(None, 2, gout), _run_raft(TR, 2, gout)
    assert_close(ours[0], ref[0], rtol=1e-4, atol_rms=1e-4, what="flow_up, 2 iterations")
    assert _rel_l2(ours[1], ref[1]) < 1e-3, _rel_l2(ours[1], ref[1])
    assert _rel_l2(ours[2], ref[2]) < 1e-3, _rel_l2(ours[2], ref[2])


def _cos(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b)))


def test_fused_closure_matches_torch_autograd_composition(fp32_convs):
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
    np.testing.assert_allclose(float(loss), float(ref.detach()), rtol=1e-4)
    assert _rel_l2(g1.cpu().numpy(), w1.grad.cpu().numpy()) < 2e-3
    assert _rel_l2(g2.cpu().numpy(), w2.grad.cpu().numpy()) < 2e-3


def test_closure_is_cuda_graph_capturable(fp32_convs):
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
    assert _rel_l2(g1.cpu().numpy(), eager[1].cpu().numpy()) < 1e-3
    w1.add_(0.01)                     # replays pick up new variable values (static buffers)
    graph.replay()
    torch.cuda.synchronize()
    assert abs(float(fo.terms[0]) - eager[0]) > 0


@pytest.mark.parametrize("name,shape,gain,kw", [("GMA", (128, 136), 0.5, {}), ("PWCNet", (128, 192), 1.0, {}),
                                                ("FlowNet2", (64, 128), 0.7, {})])
def test_other_networks_gpu_match_reference_flows(golden, fp32_convs, name, shape, gain, kw):
    """GMA / PWCNet / FlowNet2 with the CUDA operators against the reference networks' CPU flows
    (tests/golden/networks.npz) and, for gradients, against the oracle's torch ops on the same GPU."""
    from oracle import torch_ref as TR
    from pcfa_b200.adapter import build_network, compute_flow
    from pcfa_b200.networks.weights import synthetic_pair
    z = golden("networks")
    idx = {"GMA": 3, "PWCNet": 4, "FlowNet2": 5}[name]
    key = {"GMA": "gma_flow", "PWCNet": "pwc_flow", "FlowNet2": "fn2_flow"}[name]
    scale = 255. if name == "PWCNet" else 1.
    res = {}
    for tag, ops in (("cuda", None), ("torch", TR)):
        if tag == "torch" and name == "FlowNet2":
            continue                                       # the oracle's FlowNet2 ops are CPU-only (C)
        net = build_network(name, device="cuda", seed=0, ops=ops, gain=gain)
        if name == "GMA":
            net.args["mixed_precision"] = False            # compare in fp32 (the reference CPU run cannot autocast)
        i1, i2 = synthetic_pair(idx, *shape)
        a, b = (i1 / scale).cuda().requires_grad_(True), (i2 / scale).cuda()
        flow = compute_flow(net, name, a, b)
        g = torch.randn(flow.shape, generator=torch.Generator().manual_seed(16)).cuda() / flow.numel()
        (flow * g).sum().backward()
        res[tag] = (flow.detach().cpu().numpy(), a.grad.cpu().numpy())
    assert_close(res["cuda"][0], z[key], rtol=1e-3, atol_rms=2e-3, what=f"{name} flow vs reference (CPU)")
    if "torch" in res:
        assert_close(res["cuda"][0], res["torch"][0], rtol=1e-3, atol_rms=1e-3, what=f"{name} flow vs torch ops (GPU)")
        assert _rel_l2(res["cuda"][1], res["torch"][1]) < 2e-2, _rel_l2(res["cuda"][1], res["torch"][1])
    assert np.isfinite(res["cuda"][1]).all() and np.abs(res["cuda"][1]).sum() > 0


def test_gma_autocast_config_runs(fp32_convs):
    """GMA's shipped config enables fp16 autocast (models/_config/gma_config.json:5): the CorrBlock gets fp32."""
    from pcfa_b200.adapter import build_network, compute_flow
    from pcfa_b200.networks.weights import synthetic_pair
    net = build_network("GMA", device="cuda", seed=0, gain=0.5)
    assert net.args["mixed_precision"]
    i1, i2 = synthetic_pair(3, 128, 136)
    a = i1.cuda().requires_grad_(True)
    flow = compute_flow(net, "GMA", a, i2.cuda())
    flow.float().abs().mean().backward()
    assert flow.shape == (1, 2, 128, 136) and torch.isfinite(a.grad).all()


@pytest.mark.gpu
def test_raft_channels_last_update_block_matches_nchw(fp32_convs):
    """The NHWC update block (channels-last lookup, fused GRU, cat kernel) against the same network with NCHW
    activations: identical arithmetic up to convolution algorithm choice."""
    from pcfa_b200.adapter import build_network
    from pcfa_b200.networks.weights import synthetic_pair
    i1, i2 = synthetic_pair(3, 128, 160)
    i1, i2 = i1.cuda(), i2.cuda()
    flows, grads = [], []
    for cl in (True, False):
        net = build_network("RAFT", device="cuda", seed=0, gain=0.5, channels_last_update=cl)
        a = i1.clone().requires_grad_(True)
        _, flow = net(a, i2, iters=4, test_mode=True)
        flow.square().mean().backward()
        flows.append(flow.detach()); grads.append(a.grad.detach())
    assert flows[0].is_contiguous()
    assert_close(flows[0].cpu().numpy(), flows[1].cpu().numpy(), what="flow (NHWC vs NCHW update block)", rtol=1e-3, atol_rms=1e-3)
    cos = float((grads[0] * grads[1]).sum() / (grads[0].norm() * grads[1].norm()))
    assert cos > 0.999, cos
