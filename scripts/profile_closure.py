"""Kernel-time breakdown of one eager RAFT closure with torch.profiler (CUPTI; no serialisation, warm caches).
usage: python scripts/profile_closure.py [nchw|cl] [topN] [RAFT|GMA] [batch]"""
import sys, collections, torch
sys.path.insert(0, '.')
import bench
from pcfa_b200 import _lib, objective as J
from torch.profiler import profile, ProfilerActivity
fmt = sys.argv[1] if len(sys.argv) > 1 else "nchw"
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
_lib.load()
device = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
name = sys.argv[3] if len(sys.argv) > 3 else "RAFT"
batch = int(sys.argv[4]) if len(sys.argv) > 4 else 1
if name == "RAFT":
    net, i1, i2 = bench.make_problem(device, 0, 436, 1024)
else:
    from pcfa_b200.adapter import build_network
    from pcfa_b200.networks.weights import synthetic_pair
    net = build_network(name, device=device, seed=0)
    HW = (375, 1242) if name in ("PWCNet", "FlowNet2") else (436, 1024)
    prs = [synthetic_pair(i, *HW) for i in range(batch)]
    i1, i2 = torch.cat([p[0] for p in prs]), torch.cat([p[1] for p in prs])
if fmt == "cl":
    net = net.to(memory_format=torch.channels_last)
if name in ("RAFT", "GMA"):
    padder, img1, img2 = bench.prepare_on_device(i1.pin_memory(), i2.pin_memory(), device)
    target = torch.zeros(img1.shape[0], 2, 436, 1024, device=device)
    iters = 12 if name == "RAFT" else 6
    fwd, scale = (lambda a, b: net(a, b, iters=iters, test_mode=True)[1]), 255.0
else:
    from pcfa_b200.adapter import compute_flow, model_takes_unit_input, preprocess_img
    unit = model_takes_unit_input(name)
    a, b = (i1.to(device), i2.to(device)) if unit else (i1.to(device) / 255., i2.to(device) / 255.)
    padder, (img1, img2) = preprocess_img(name, a, b)
    img1, img2 = img1.contiguous(), img2.contiguous()
    target = torch.zeros(img1.shape[0], 2, *HW, device=device)
    fwd, scale = (lambda a, b: compute_flow(net, name, a, b, test_mode=True)), (1.0 if unit else 255.0)
fo = J.FusedObjective(fwd, img1, img2, target, mode=J.BOX_COV, joint=False,
                      pad=padder.top_left, eps_box=bench.EPS_BOX, scale=scale, delta_bound=bench.DELTA_BOUND, mu=bench.MU, loss="aee")
w1, w2 = bench.init_vars(img1), bench.init_vars(img2)
g1, g2 = torch.empty_like(w1), torch.empty_like(w2)
for _ in range(4):
    fo.evaluate(w1, w2, g1, g2)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    fo.evaluate(w1, w2, g1, g2)
e1.record(); torch.cuda.synchronize()
print("eager closure ms:", e0.elapsed_time(e1) / 3)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    fo.evaluate(w1, w2, g1, g2)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        agg[ev.name][0] += 1; agg[ev.name][1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
tot = sum(v[1] for v in agg.values())
print(f"{len(agg)} kernels, {sum(v[0] for v in agg.values())} launches, {tot:.0f} us of kernel time")
groups = collections.OrderedDict([("layout conversions", ("nchwToNhwc", "nhwcToNchw", "tensorTransform", "transpose")), ("norms", ("batch_norm", "bn_fw", "bn_bw")),
          ("convs/gemms", ("cutlass", "xmma", "implicit_convolve", "dgrad", "wgrad", "winograd", "sgemm", "gemm", "conv")), ("pcfa", ("pcfa::",)),
          ("elementwise/cat/other", ("",))])
gs = collections.OrderedDict((k, 0.0) for k in groups)
for name, (n, t) in agg.items():
    for gname, keys in groups.items():
        if any(k in name for k in keys):
            gs[gname] += t; break
print({k: round(v) for k, v in gs.items()})
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
    print(f"{t:9.0f} us {n:5d}  {name[:110]}")
