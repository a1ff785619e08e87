import sys, torch
sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
from test_gpu_lbfgs import _problem
from pcfa_b200.lbfgs import DeviceLBFGS
for n, history in ((5000, 100), (1003, 5)):
    f, x0 = _problem(n, n + history)
    p = x0.clone().requires_grad_(True)
    opt = torch.optim.LBFGS([p], max_iter=10, history_size=history)
    def closure_t():
        opt.zero_grad(); l = f(p); l.backward(); return l
    flat_p, flat_g = x0.clone(), torch.zeros_like(x0)
    dev = DeviceLBFGS(flat_p, flat_g, max_iter=10, history_size=history)
    def closure_d():
        xp = flat_p.detach().clone().requires_grad_(True); l = f(xp); (g,) = torch.autograd.grad(l, xp); flat_g.copy_(g); return l.detach()
    for s in range(4):
        lt = opt.step(closure_t); ld = dev.step(closure_d)
        rel = float((flat_p - p.detach()).norm() / p.detach().norm())
        print(n, history, "step", s, "loss torch %.6f device %.6f" % (float(lt), float(ld)), "rel", rel, "f now %.6f %.6f" % (float(f(p)), float(f(flat_p))),
              "n_iter", opt.state[p]["n_iter"], dev.state["n_iter"], "evals", opt.state[p]["func_evals"], dev.state["func_evals"], "hist", dev.state["num_old"], len(opt.state[p].get("old_dirs", [])))
