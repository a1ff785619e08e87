"""On-device L-BFGS (pcfa_b200.lbfgs.DeviceLBFGS) against torch.optim.LBFGS — the optimiser the reference uses
(attack_PCFA.py:97,114) — on the same GPU: same update rule, so the iterates agree to reduction-order noise."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _problem(n, seed):
    g = torch.Generator().manual_seed(seed)
    a = torch.exp(6 * torch.rand(n, generator=g) - 3).cuda()            # condition number ~1e5: no convergence in 40 iterations
    b = torch.randn(n, generator=g).cuda()
    x0 = torch.randn(n, generator=g).cuda() * 0.1

    def f(x):
        return ((x * a - b) ** 2).sum() + 0.1 * (x ** 4).sum() + 0.3 * (x[:-1] * x[1:]).sum()
    return f, x0


@pytest.mark.parametrize("direction", ["compact", "two_loop"])
@pytest.mark.parametrize("n,history,steps", [(5000, 100, 3), (1003, 5, 4), (200000, 100, 2), (4096, 7, 6)])
def test_device_lbfgs_matches_torch_lbfgs(n, history, steps, direction):
    from pcfa_b200.lbfgs import DeviceLBFGS
    f, x0 = _problem(n, n + history)
    # torch
    p = x0.clone().requires_grad_(True)
    opt = torch.optim.LBFGS([p], max_iter=10, history_size=history)
    evals_t = [0]

    def closure_t():
        evals_t[0] += 1
        opt.zero_grad()
        l = f(p)
        l.backward()
        return l
    # device
    flat_p, flat_g = x0.clone(), torch.zeros_like(x0)
    dev = DeviceLBFGS(flat_p, flat_g, max_iter=10, history_size=history, direction=direction)
    evals_d = [0]

    def closure_d():
        evals_d[0] += 1
        xp = flat_p.detach().clone().requires_grad_(True)
        l = f(xp)
        (g,) = torch.autograd.grad(l, xp)
        flat_g.copy_(g)
        return l.detach()
    for s in range(steps):
        lt = float(opt.step(closure_t).detach())
        ld = float(dev.step(closure_d))
        assert ld == pytest.approx(lt, rel=1e-3, abs=1e-5), f"loss at the start of step {s}"
        rel = float((flat_p - p.detach()).norm() / p.detach().norm())
        assert rel < 5e-3, f"iterates after step {s}: rel diff {rel:.2e}"
        if s == 0:                                           # before any stopping test can fire on fp32 noise
            assert evals_d[0] == evals_t[0] == 10
            assert dev.state["func_evals"] == opt.state[p]["func_evals"] and dev.state["n_iter"] == opt.state[p]["n_iter"] == 10
    assert float(f(flat_p)) < 0.5 * float(f(x0))


def test_device_lbfgs_rejects_cpu_tensors():
    from pcfa_b200.lbfgs import DeviceLBFGS
    with pytest.raises(RuntimeError):
        DeviceLBFGS(torch.zeros(4), torch.zeros(4))
