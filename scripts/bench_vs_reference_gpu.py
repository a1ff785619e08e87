"""'The kernel to beat on the same box' (BASELINE.md §3.5): the reference's own CUDA extensions compiled for
sm_100a (oracle/_ref, built by oracle/build_ref.py) and the reference's torch-op CorrBlock (oracle/torch_ref.py on
cuda) timed beside the pcfa_b200 entry points at the BASELINE shapes.  CUDA events on the launching stream, L2
flushed before every sample, the duration of an empty event bracket subtracted.  Measurement script (not product):
it is the one place outside tests/ and bench.py's CPU arm where oracle/ runs, as the thing being compared against.

    python scripts/bench_vs_reference_gpu.py  →  gpurun_out/bench_vs_reference_gpu.json
"""
import json
import statistics
import sys

import torch

sys.path.insert(0, '.')
from oracle import build_ref                                  # noqa: E402
from oracle import torch_ref as TR                            # noqa: E402
from pcfa_b200 import flownet2_ops as F2                      # noqa: E402
from pcfa_b200.corr_block import CorrBlock                    # noqa: E402
from pcfa_b200.spatial_correlation_sampler import spatial_correlation_sample  # noqa: E402

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
rows = []


def _time(fn, reps=10, warm=3):
    ts = []
    for i in range(reps + warm):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= warm:
            ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts)


EMPTY = _time(lambda: None, reps=30)


def compare(name, ours, ref, note=""):
    t_o = max(_time(ours) - EMPTY, 0.1)
    t_r = max(_time(ref) - EMPTY, 0.1) if ref is not None else None
    r = dict(op=name, pcfa_b200_us=round(t_o, 1), reference_us=None if t_r is None else round(t_r, 1),
             speedup=None if t_r is None else round(t_r / t_o, 2), note=note)
    rows.append(r)
    print(r, flush=True)


g = torch.Generator().manual_seed(0)


def rn(*shape, scale=1.0):
    return (scale * torch.randn(*shape, generator=g)).cuda()


# ---------------------------------------------------------------- RAFT CorrBlock (reference = torch ops on this GPU)
f1, f2 = rn(1, 256, 55, 128).requires_grad_(True), rn(1, 256, 55, 128).requires_grad_(True)
ys, xs = torch.meshgrid(torch.arange(55.), torch.arange(128.), indexing="ij")
coords = (torch.stack([xs, ys])[None] + 3 * torch.randn(1, 2, 55, 128, generator=g)).cuda()
torch.backends.cuda.matmul.allow_tf32 = False                 # the reference runs torch.matmul in fp32 (torch 1.7.1 default on its hardware)
with torch.no_grad():
    compare("CorrBlock build (all-pairs + 3 pools), 1x256x55x128", lambda: CorrBlock(f1, f2), lambda: TR.CorrBlock(f1, f2),
            "reference: torch.matmul fp32 + /sqrt(C) + 3 avg_pool2d (models/raft/corr.py:13-27,52-60)")
    ours_blk, ref_blk = CorrBlock(f1, f2), TR.CorrBlock(f1, f2)
    compare("CorrBlock lookup x1 (4 levels, r=4)", lambda: ours_blk(coords), lambda: ref_blk(coords),
            "reference: ~60 launches incl. 4 grid_sample (models/raft/corr.py:29-50)")


def closure(blk_cls, iters=12):
    def run():
        f1.grad = f2.grad = None
        blk = blk_cls(f1, f2)
        acc = 0
        for i in range(iters):
            acc = acc + blk(coords + 0.25 * i).sum()
        acc.backward()
    return run


compare("CorrBlock build + 12 lookups + full backward", closure(CorrBlock), closure(TR.CorrBlock),
        "the cost-volume share of one RAFT closure; reference = autograd over the torch ops")

# ---------------------------------------------------------------- FlowNet2 ops (reference = its CUDA extensions, sm_100a)
rc, rr, rn_ = (build_ref.load_cuda(n) for n in ("correlation_cuda", "resample2d_cuda", "channelnorm_cuda"))
a, b = rn(1, 256, 48, 160), rn(1, 256, 48, 160)
args = (20, 1, 20, 1, 2, 1)
e = lambda: a.new_empty(0)                                    # noqa: E731
out_o, out_r = e(), e()
F2.correlation_cuda.forward(a, b, e(), e(), out_o, *args)
go = torch.randn_like(out_o)
g1, g2, r1, r2 = e(), e(), e(), e()
compare("FlowNet2 correlation fwd 1x256x48x160", lambda: F2.correlation_cuda.forward(a, b, e(), e(), out_o, *args),
        (lambda: rc.forward(a, b, r1, r2, out_r, *args)) if rc else None,
        "reference: correlation_cuda_kernel.cu:73-147 incl. its two padded NHWC copies")
compare("FlowNet2 correlation bwd", lambda: F2.correlation_cuda.backward(a, b, e(), e(), go, g1, g2, *args),
        (lambda: rc.backward(a, b, r1, r2, go, e(), e(), *args)) if rc else None, "correlation_cuda_kernel.cu:150-334")

img, flow = rn(1, 3, 384, 1280), rn(1, 2, 384, 1280, scale=3.0)
o1, o2 = torch.zeros_like(img), torch.zeros_like(img)
gi, gf, go = torch.zeros_like(img), torch.zeros_like(flow), torch.randn_like(img)
compare("Resample2d fwd 1x3x384x1280", lambda: F2.resample2d_cuda.forward(img, flow, o1, 1, True),
        (lambda: rr.forward(img, flow, o2, 1, True)) if rr else None, "resample2d_kernel.cu:15-72")
compare("Resample2d bwd", lambda: F2.resample2d_cuda.backward(img, flow, go, gi, gf, 1, True),
        (lambda: rr.backward(img, flow, go, gi, gf, 1, True)) if rr else None, "resample2d_kernel.cu:75-198 (3 kernels)")
n1, n2 = torch.zeros(1, 1, 384, 1280, device="cuda"), torch.zeros(1, 1, 384, 1280, device="cuda")
gn, gx = torch.randn_like(n1), torch.zeros_like(img)
compare("ChannelNorm fwd 1x3x384x1280", lambda: F2.channelnorm_cuda.forward(img, n1, 2),
        (lambda: rn_.forward(img, n2, 2)) if rn_ else None, "channelnorm_kernel.cu:18-60")
compare("ChannelNorm bwd", lambda: F2.channelnorm_cuda.backward(img, n1, gn, gx, 2),
        (lambda: rn_.backward(img, n1, gn, gx, 2)) if rn_ else None, "channelnorm_kernel.cu:63-96")

# ---------------------------------------------------------------- PWCNet sampler (reference = its CUDA build, sm_100a)
rs = build_ref.load_cuda("spatial_correlation_sampler_backend_cuda")
P9 = (1, 1, 9, 9, 0, 0, 1, 1, 1, 1, 1, 1)
tot = dict(of=0.0, ob=0.0, rf=0.0, rb=0.0)
for (C, H, W) in [(196, 6, 20), (128, 12, 40), (96, 24, 80), (64, 48, 160), (32, 96, 320)]:
    a, b = rn(1, C, H, W).requires_grad_(True), rn(1, C, H, W).requires_grad_(True)
    with torch.no_grad():
        compare(f"SCS fwd C{C} {H}x{W} patch 9", lambda: spatial_correlation_sample(a, b, 1, 9, 1, 0, 1, 1),
                (lambda: rs.forward(a, b, *P9)) if rs else None)
    out = spatial_correlation_sample(a, b, 1, 9, 1, 0, 1, 1)
    go = torch.randn_like(out)
    compare(f"SCS bwd C{C} {H}x{W}", lambda: torch.autograd.grad(out, [a, b], go, retain_graph=True),
            (lambda: rs.backward(a.detach(), b.detach(), go, *P9)) if rs else None)
    tot["of"] += rows[-2]["pcfa_b200_us"]
    tot["ob"] += rows[-1]["pcfa_b200_us"]
    if rs:
        tot["rf"] += rows[-2]["reference_us"]
        tot["rb"] += rows[-1]["reference_us"]
rows.append(dict(op="SCS 5 PWCNet levels total fwd", pcfa_b200_us=round(tot["of"], 1), reference_us=round(tot["rf"], 1) or None))
rows.append(dict(op="SCS 5 PWCNet levels total bwd", pcfa_b200_us=round(tot["ob"], 1), reference_us=round(tot["rb"], 1) or None))

json.dump(dict(device=torch.cuda.get_device_name(0), empty_bracket_us=round(EMPTY, 2), rows=rows),
          open("gpurun_out/bench_vs_reference_gpu.json", "w"), indent=1)
