"""One eager RAFT closure evaluation at the BASELINE shape inside a cudaProfilerStart/Stop range (cuDNN algorithms
autotuned during the warm-up, outside the range):
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file L.csv python scripts/run_closure.py
then scripts/summarise_launches.py L.csv."""
import sys, torch
sys.path.insert(0, '.')
import bench
from pcfa_b200 import _lib, objective as J
_lib.load()
device = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
net, i1, i2 = bench.make_problem(device, 0, 436, 1024)
padder, img1, img2 = bench.prepare_on_device(i1.pin_memory(), i2.pin_memory(), device)
target = torch.zeros(1, 2, 436, 1024, device=device)
fo = J.FusedObjective(lambda a, b: net(a, b, iters=12, test_mode=True)[1], img1, img2, target, mode=J.BOX_COV, joint=False,
                      pad=padder.top_left, eps_box=bench.EPS_BOX, scale=255.0, delta_bound=bench.DELTA_BOUND, mu=bench.MU, loss="aee")
w1, w2 = bench.init_vars(img1), bench.init_vars(img2)
g1, g2 = torch.empty_like(w1), torch.empty_like(w2)
for _ in range(3):
    fo.evaluate(w1, w2, g1, g2)
torch.cuda.synchronize()
torch.cuda.profiler.start()
fo.evaluate(w1, w2, g1, g2)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", _lib.launch_count())
