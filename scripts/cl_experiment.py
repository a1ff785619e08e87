"""Does torch.channels_last help the PWCNet / FlowNet2 closures (cuDNN's sm_100 convolution kernels are NHWC-only)?
Closure time from CUDA-graph replays, NCHW vs channels-last weights (activations follow the weights)."""
import statistics, sys, torch
sys.path.insert(0, '.')
from pcfa_b200 import _lib, objective as J
from pcfa_b200.adapter import build_network, preprocess_img, model_takes_unit_input
from pcfa_b200.attack import GraphedEvaluate, _net_forward, resolve_mu
from pcfa_b200.networks.weights import synthetic_pair
_lib.load()
dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
for net_name in sys.argv[1:] or ["PWCNet", "FlowNet2"]:
    for cl in (False, True):
        torch.cuda.empty_cache()
        model = build_network(net_name, device=dev, seed=0, gain=0.5)
        if cl:
            model = model.to(memory_format=torch.channels_last)
        H, W = 375, 1242
        i1, i2 = synthetic_pair(0, H, W)
        i1, i2 = i1.to(dev), i2.to(dev)
        unit = model_takes_unit_input(net_name)
        a, b = (i1, i2) if unit else (i1 / 255., i2 / 255.)
        padder, (a, b) = preprocess_img(net_name, a, b)
        a, b = a.contiguous(), b.contiguous()
        fo = J.FusedObjective(_net_forward(model, net_name, None), a, b, torch.zeros(1, 2, H, W, device=dev), mode=J.BOX_COV, joint=False,
                              pad=padder.top_left, eps_box=1e-7, scale=1.0 if unit else 255.0, delta_bound=0.005,
                              mu=resolve_mu(-1., 0.005, "zero"), loss="aee")
        v1 = torch.atanh(2. * (1. - 1e-7) * a - (1 - 1e-7)).contiguous(); v2 = torch.atanh(2. * (1. - 1e-7) * b - (1 - 1e-7)).contiguous()
        ev = GraphedEvaluate(fo, v1, v2, use_graph=True)
        for _ in range(5): ev()
        ts = []
        for _ in range(20):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ev(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        print(net_name, "channels_last" if cl else "nchw", "closure ms %.3f" % statistics.median(ts), "loss %.6f" % float(fo.terms[0]), flush=True)
        del ev, fo, model
