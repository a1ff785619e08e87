"""Host-side semantics of the round-2 glue helpers on CPU tensors: the network definitions run on CPU with plain torch ops
(that is how reference flows with the oracle's operators are produced), so every fused helper has to reduce
to the reference's own expression there.  (The CUDA paths are tested against these same expressions in test_gpu_ops.py.)"""
import torch


def test_add_relu_is_relu_of_sum_on_cpu():
    from pcfa_b200.conv_ops import add_relu
    a, b = torch.randn(2, 8, 5, 6), torch.randn(2, 8, 5, 6)
    out = add_relu(a, b, twin=True)
    assert torch.equal(out, torch.relu(a + b)) and getattr(out, "_pcfa_twin", None) is None


def test_fork_is_identity_on_cpu():
    from pcfa_b200.conv_ops import fork
    x = torch.randn(1, 4, 3, 3, requires_grad=True)
    a, b = fork(x)
    assert a is x and b is x


def test_flow_step_and_padded_flow_on_cpu():
    """coords1 + delta and the zero-padded flow of models/raft/raft.py:126,131."""
    from pcfa_b200.conv_ops import flow_step, padded_flow
    c1, c0, d = torch.randn(2, 2, 4, 6), torch.randn(2, 2, 4, 6), torch.randn(2, 2, 4, 6)
    new, flow = flow_step(c1, c0, d, 8)
    assert torch.equal(new, c1 + d) and flow.shape == (2, 8, 4, 6)
    assert torch.equal(flow[:, :2], c1 + d - c0) and float(flow[:, 2:].abs().max()) == 0.0
    p = padded_flow(d, 4)
    assert p.shape == (2, 4, 4, 6) and torch.equal(p[:, :2], d) and float(p[:, 2:].abs().max()) == 0.0


def test_dense_conv_cat_is_cat_of_leaky_conv_and_input_on_cpu():
    """x = cat((LeakyReLU(conv(x)), x), 1) of models/PWCNet/PWCNet.py:253-257."""
    from pcfa_b200.conv_ops import dense_conv_cat
    conv = torch.nn.Conv2d(8, 12, 3, padding=1)
    x = torch.randn(1, 8, 5, 7)
    ref = torch.cat((torch.nn.functional.leaky_relu(conv(x), 0.1), x), 1)
    assert torch.allclose(dense_conv_cat(conv, x, 0.1), ref, atol=1e-6)


def test_residual_block_same_values_with_and_without_twin_handles_on_cpu():
    """ResidualBlock (models/raft/extractor.py:8-56) chained twice: the skip branch takes the twin handle when there is one."""
    from pcfa_b200.networks.raft import ResidualBlock
    torch.manual_seed(0)
    b1, b2 = ResidualBlock(8, 8, "instance"), ResidualBlock(8, 16, "instance", stride=2)
    x = torch.randn(1, 8, 12, 16)
    y = b2(b1(x))
    h = b1.relu(x + b1.relu(b1.norm2(b1.conv2(b1.relu(b1.norm1(b1.conv1(x)))))))
    ref = b2.relu(b2.downsample(h) + b2.relu(b2.norm2(b2.conv2(b2.relu(b2.norm1(b2.conv1(h)))))))
    assert torch.allclose(y, ref, atol=1e-5)
