"""Which Python lines launch the remaining ATen element-wise kernels of one RAFT closure (copies, adds, clamps)?
torch.profiler with stacks: for every CPU op whose CUDA children include a matching kernel, print shapes and the first
repo frame.  usage: python scripts/attribute_glue.py [pattern ...]"""
import collections, sys, torch
sys.path.insert(0, '.')
import bench
from pcfa_b200 import _lib, objective as J
from torch.profiler import profile, ProfilerActivity
pats = sys.argv[1:] or ["direct_copy", "CUDAFunctor_add", "clamp", "BinaryFunctor", "FillFunctor", "Memset", "nhwcAddPadding"]
_lib.load()
device = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
net, i1, i2 = bench.make_problem(device, 0, 436, 1024)
net = net.to(memory_format=torch.channels_last)
padder, img1, img2 = bench.prepare_on_device(i1.pin_memory(), i2.pin_memory(), device)
target = torch.zeros(img1.shape[0], 2, 436, 1024, device=device)
fo = J.FusedObjective(lambda a, b: net(a, b, iters=12, test_mode=True)[1], img1, img2, target, mode=J.BOX_COV, joint=False,
                      pad=padder.top_left, eps_box=bench.EPS_BOX, scale=255.0, delta_bound=bench.DELTA_BOUND, mu=bench.MU, loss="aee")
v1 = torch.zeros_like(img1); v2 = torch.zeros_like(img2)
g1 = torch.empty_like(v1); g2 = torch.empty_like(v2)
for _ in range(3):
    fo.evaluate(v1, v2, g1, g2)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True, record_shapes=True) as prof:
    fo.evaluate(v1, v2, g1, g2)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type.name != "CPU" or not ev.kernels:
        continue
    for k in ev.kernels:
        hit = [p for p in pats if p in k.name]
        if not hit:
            continue
        frame = next((f for f in (ev.stack or []) if "/root/repo" in f or "pcfa_b200" in f or "gpurun" in f), (ev.stack or ["?"])[0] if ev.stack else "?")
        key = (hit[0], ev.name, str(ev.input_shapes)[:90], frame.strip()[-90:])
        agg[key][0] += 1
        agg[key][1] += k.duration
for key, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    print("%7.1f us %3d  %-16s %-28s %s\n%s%s" % (us, n, key[0], key[1], key[2], " " * 24, key[3]))
