"""Where does one outer L-BFGS step of pcfa_attack go?  (RAFT 436x1024; wall clock around synchronised sections.)"""
import sys, time, torch
sys.path.insert(0, '.')
import bench
from pcfa_b200 import _lib, attack as A, objective as J
from pcfa_b200.lbfgs import DeviceLBFGS
_lib.load()
dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
net, i1, i2 = bench.make_problem(dev, 0, 436, 1024)
padder, img1, img2 = bench.prepare_on_device(i1.pin_memory(), i2.pin_memory(), dev)
fo = J.FusedObjective(lambda a, b: net(a, b, iters=12, test_mode=True)[1], img1, img2, torch.zeros(1, 2, 436, 1024, device=dev),
                      mode=J.BOX_COV, joint=False, pad=padder.top_left, eps_box=1e-7, scale=255.0, delta_bound=0.005, mu=5e5, loss="aee")
n1 = img1.numel()
flat_p, flat_g = torch.empty(2 * n1, device=dev), torch.zeros(2 * n1, device=dev)
flat_p[:n1].copy_(bench.init_vars(img1).reshape(-1)); flat_p[n1:].copy_(bench.init_vars(img2).reshape(-1))
v1, v2 = flat_p[:n1].view_as(img1), flat_p[n1:].view_as(img2)
ev = A.GraphedEvaluate(fo, v1, v2, use_graph=True, g1=flat_g[:n1].view_as(img1), g2=flat_g[n1:].view_as(img2))
rp = A.GraphedPredict(fo, v1, v2, use_graph=True)
opt = DeviceLBFGS(flat_p, flat_g, max_iter=10)
def sync(): torch.cuda.synchronize(); return time.perf_counter()
for _ in range(3): ev()
t0 = sync()
for _ in range(10): ev()
t1 = sync(); print("10 graph replays back to back: %.2f ms" % ((t1 - t0) * 1e3))
for _ in range(10): ev(); torch.cuda.synchronize()
t2 = sync(); print("10 graph replays, synchronised after each: %.2f ms" % ((t2 - t1) * 1e3))
for k in range(6):
    t0 = sync(); opt.step(lambda: ev()); t1 = sync(); rp(); t2 = sync()
    h = opt.history
    print("outer step %d: optimizer.step %.2f ms, re-prediction %.2f ms, history %s" % (k, (t1 - t0) * 1e3, (t2 - t1) * 1e3, h))
# the optimiser's own kernels at the current history
P = _lib.ptr; lib = _lib.load(); s = _lib.stream()
t0 = sync()
for _ in range(10):
    lib.pcfa_lbfgs_update_history(P(opt.g), P(opt.g_prev), P(opt.d), 1.0, P(opt.S), P(opt.Y), P(opt.ro), P(opt.hdiag), P(opt.ring[opt.cur]), P(opt.ring[opt.cur ^ 1]), P(opt.sc), P(opt.ws), opt.n, opt.m, s)
t1 = sync()
for _ in range(10):
    lib.pcfa_lbfgs_direction_step(P(opt.S), P(opt.Y), P(opt.ro), P(opt.g), P(opt.hdiag), P(opt.d), P(opt.ring[opt.cur]), P(opt.p), 0.0, 1e-9, P(opt.sc[2:]), P(opt.ws), opt.n, opt.m, s)
t2 = sync()
print("update_history %.3f ms, direction_step %.3f ms per call at history %s" % ((t1 - t0) * 100, (t2 - t1) * 100, opt.history))
