#!/usr/bin/env python
"""bench.py — PCFA closure evaluations per second (forward + backward through the flow network).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One "step" = one PCFA closure evaluation: box transform → network forward → unpad →
loss + penalty → backward to the perturbation variables.  Workload at every N: BASELINE.json
configs[1] — RAFT (12 GRU iterations, 4-level all-pairs pyramid), disjoint perturbations with the
change-of-variables box constraint, zero target, AEE loss, one Sintel-shaped 436x1024 pair per GPU
(weak scaling: pairs are independent, no data-path collective).  Weights are deterministic synthetic
values and the pair is synthetic (no network access for checkpoints/datasets).

Prints ONE JSON line (see the task contract): value = closures/s with inputs resident in HBM and
the closure replayed from a CUDA graph; e2e = the same through the public API with host buffers
(every step: H2D of its pair from pinned memory + input preparation, prefetched on a copy stream while the previous
closure runs, and D2H of its loss — all inside the timed region); roofline = the dominant pcfa_b200
kernel measured with CUDA events in an instrumented eager pass; cpu_baseline = the oracle port of
the reference closure timed on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "PCFA steps/sec (fwd+bwd, device-timed)"      # BASELINE.json metric; 1 step = 1 closure evaluation
UNIT = "steps/s"
H_IMG, W_IMG = 436, 1024
DELTA_BOUND, EPS_BOX = 0.005, 1e-7
MU = 2500.0 / DELTA_BOUND           # attack_PCFA.py:578-583, zero target


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of CUDA-graph replay")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-samples", type=int, default=3)
    ap.add_argument("--channels-last", action="store_true", help="run the network's convolutions in NHWC memory format")
    ap.add_argument("--no-cudnn-benchmark", action="store_true", help="disable cuDNN autotuning of the convolution algorithms during warm-up")
    ap.add_argument("--universal-pairs", type=int, default=8, help="pairs per GPU of the universal-perturbation record (0 = skip it)")
    ap.add_argument("--universal-steps", type=int, default=0, help="timed closures of the universal record (default: --steps)")
    ap.add_argument("--height", type=int, default=H_IMG)
    ap.add_argument("--width", type=int, default=W_IMG)
    return ap.parse_args()


def config_dict(args, n):
    return {"workload": "RAFT PCFA disjoint, change_of_variables, zero target, aee, delta_bound=0.005, "
                        "12 GRU iters, 4-level all-pairs pyramid, %dx%d (padded 440x1024), batch 1 per GPU" % (args.height, args.width),
            "pairs_per_gpu": 1, "parallelism": "pair-sharded x%d (no collective)" % n,
            "l2": "per-step working set ~1.6 GB (261 MB pyramid + 261 MB gradient pyramid + activations) >> 126 MB L2, no explicit flush"}


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------ workload
def make_problem(device, rank, H, W, ops=None):
    """Network + objective for one synthetic pair, exactly as pcfa_attack sets it up
    (attack_PCFA.py:55-131): /255, pad, w = atanh(2(1-eps)I - (1-eps)), zero target."""
    from pcfa_b200.adapter import build_network, preprocess_img
    from pcfa_b200.networks.weights import synthetic_pair
    net = build_network("RAFT", device=device, seed=0, ops=ops)
    i1, i2 = synthetic_pair(rank, H, W)
    return net, i1, i2


def prepare_on_device(i1, i2, device):
    from pcfa_b200.adapter import preprocess_img
    a, b = i1.to(device, non_blocking=True) / 255.0, i2.to(device, non_blocking=True) / 255.0
    padder, (a, b) = preprocess_img("RAFT", a, b)
    return padder, a.contiguous(), b.contiguous()


def init_vars(img):
    return torch.atanh(2.0 * (1.0 - EPS_BOX) * img - (1 - EPS_BOX)).contiguous()


def measured_traffic(entry_point):
    """dram bytes per launch from this round's ncu --set full capture (profiles/traffic_r*.json), or None."""
    files = sorted((ROOT / "profiles").glob("traffic_r*.json"))
    if not files:
        return None
    try:
        ep = json.loads(files[-1].read_text())["entry_points"]
        for name in (entry_point, entry_point[:-3] if entry_point.endswith("_cl") else None,
                     entry_point[:-4] if entry_point.endswith("_occ") else None):
            if name and name in ep:
                return ep[name]
        return None
    except Exception:
        return None


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------ own arm
def run_b200(args):
    import torch.distributed as dist
    from pcfa_b200 import _lib
    from pcfa_b200 import objective as J
    from pcfa_b200 import profiling
    from pcfa_b200.adapter import preprocess_img

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    _lib.load()

    torch.backends.cudnn.benchmark = not args.no_cudnn_benchmark
    net, i1_host, i2_host = make_problem(device, rank, args.height, args.width)
    if args.channels_last:
        net = net.to(memory_format=torch.channels_last)
    i1_pin, i2_pin = i1_host.pin_memory(), i2_host.pin_memory()
    padder, img1, img2 = prepare_on_device(i1_pin, i2_pin, device)
    target = torch.zeros(1, 2, args.height, args.width, device=device)
    fo = J.FusedObjective(lambda a, b: net(a, b, iters=12, test_mode=True)[1], img1, img2, target,
                          mode=J.BOX_COV, joint=False, pad=padder.top_left, eps_box=EPS_BOX, scale=255.0,
                          delta_bound=DELTA_BOUND, mu=MU, loss="aee")
    w1, w2 = init_vars(img1), init_vars(img2)
    w1 += 0.01 * torch.randn_like(w1)
    w2 += 0.01 * torch.randn_like(w2)
    g1, g2 = torch.empty_like(w1), torch.empty_like(w2)

    def step():
        return fo.evaluate(w1, w2, g1, g2)[0]

    warm = max(3, args.warmup)
    for _ in range(warm):
        step()
    torch.cuda.synchronize()

    graph = None
    if not args.no_graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step()
        for _ in range(2):
            graph.replay()
        torch.cuda.synchronize()
    run = graph.replay if graph is not None else step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-timed region (inputs resident in HBM)
    lc0 = _lib.launch_count()
    step()
    launches_per_step = _lib.launch_count() - lc0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with ClockSampler(local) as clocks:
        ev0.record()
        for _ in range(args.steps):
            run()
        ev1.record()
        barrier()
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms], device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * args.steps / (ms_total / 1e3)

    # ---- end to end through the public API with host buffers
    # Every step copies its pair from pinned host memory and reads its loss back.  The copy and the input preparation of step
    # k+1 run on a side stream (copy engine + a few small kernels into a double-buffered staging pair) while step k's closure
    # runs; the compute stream only does the 2 x 5 MB device-to-device hand-over into the closure's input buffers.
    loss_host = torch.empty(1).pin_memory()
    copy_stream = torch.cuda.Stream(device=device)
    stage = [[torch.empty_like(i1_pin, device=device), torch.empty_like(i2_pin, device=device), None, None] for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    main_stream = torch.cuda.current_stream()

    def prefetch(k):
        st = stage[k % 2]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[k % 2])          # the closure that read this staging pair has taken it over
            st[0].copy_(i1_pin, non_blocking=True)
            st[1].copy_(i2_pin, non_blocking=True)
            _, (st[2], st[3]) = preprocess_img("RAFT", st[0] / 255.0, st[1] / 255.0)
            ready[k % 2].record(copy_stream)

    def e2e_step(k):
        prefetch(k + 1)
        main_stream.wait_event(ready[k % 2])
        fo.image1.copy_(stage[k % 2][2])
        fo.image2.copy_(stage[k % 2][3])
        consumed[k % 2].record(main_stream)
        run()
        loss_host.copy_(fo.terms[:1], non_blocking=True)

    for e in consumed:
        e.record(main_stream)
    prefetch(0)
    for k in range(3):
        e2e_step(k)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for k in range(3, 3 + args.steps):
        e2e_step(k)
    e1.record()
    barrier()
    wall = time.perf_counter() - t0
    e2e_ms = max(e0.elapsed_time(e1), wall * 1e3)
    t = torch.tensor([e2e_ms], device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * args.steps / (float(t.item()) / 1e3)
    h2d = i1_pin.numel() * 4 + i2_pin.numel() * 4

    out = {"metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
           "warmup": warm, "ms_per_step": round(ms_total / args.steps, 4), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32 (tensors fp32 end to end; cuDNN convolutions with torch's default TF32; cost-volume build bf16x3 split with fp32 accumulation ~1e-5, its backward TF32 3e-4)",
           "data": "synthetic", "config": config_dict(args, world),
           "e2e": {"value": round(e2e_value, 3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
           "gpu_launches": int(launches_per_step * args.steps), "gpu_launches_per_step": int(launches_per_step),
           "cuda_graph": graph is not None, "cudnn_benchmark": not args.no_cudnn_benchmark, "clocks": clocks.summary(), "loss": float(fo.terms[0].item())}

    universal = None
    if args.universal_pairs > 0:
        try:
            universal = bench_universal(args, net, device, world, rank, barrier)
        except Exception as e:                                   # the headline line must not depend on it
            universal = {"error": repr(e)[:300]}
            if world > 1:
                raise
    if universal is not None:
        out["universal"] = universal

    if rank == 0:
        # ---- per-kernel roofline, instrumented eager pass on the launching stream
        peak, peak_src = peaks()
        table = profiling.kernel_table(step, n_steps=3, B=1, C=256, H=img1.shape[2] // 8, W=img1.shape[3] // 8,
                                       iters=12, peak_gbs=peak, img_numel=img1.numel(), flow_numel=2 * img1.shape[2] * img1.shape[3])
        out["kernels"] = table
        def roof(r):
            return {"bound": "hbm", "kernel": r["name"], "achieved": r["achieved_gbs"], "peak": peak, "unit": "GB/s",
                    "frac": round(r["achieved_gbs"] / peak, 4), "traffic": measured_traffic(r["name"]),
                    "algorithmic_bytes": r["algorithmic_bytes"], "avg_us": r["avg_us"],
                    "launches_per_step": r["launches_per_step"], "peak_source": peak_src}
        corr = [r for r in table if r["name"].startswith("pcfa_corr_") and r.get("algorithmic_bytes")]
        if corr:
            # BASELINE.json's metric is "corr kernel % roofline": the headline is the WORST correlation kernel; the
            # largest aggregate among the other pcfa_b200 kernels (glue) is reported beside it
            worst = min(corr, key=lambda r: r["achieved_gbs"])
            out["roofline"] = roof(worst)
            out["roofline"]["selection"] = "lowest fraction among the correlation entry points " + str(sorted(r["name"] for r in corr))
            out["roofline"]["how"] = ("CUDA events around each entry-point call in an eager pass of the same step, behind a device-side spin "
                                      "so host enqueue latency is excluded, minus the duration of an empty event bracket (graph replay cannot "
                                      "be event-bracketed per kernel)")
            tb = sum(r["algorithmic_bytes"] * r["launches_per_step"] for r in corr)
            tt = sum(r["total_us_per_step"] for r in corr)
            out["roofline"]["all_corr_kernels"] = {"algorithmic_bytes_per_step": int(tb), "us_per_step": round(tt, 1),
                                                   "achieved": round(tb / tt / 1e3, 1), "frac": round(tb / tt / 1e3 / peak, 4)}
            glue = [r for r in table if not r["name"].startswith("pcfa_corr_") and r.get("algorithmic_bytes")]
            if glue:
                out["roofline_glue"] = roof(max(glue, key=lambda r: r["total_us_per_step"]))
        if world == 1:
            # outer L-BFGS steps per second of the attack loop itself (10 closures + one
            # re-prediction + the optimiser's vector algebra per step), pcfa_attack on the same pair
            try:
                import time as _time
                from pcfa_b200.attack import pcfa_attack
                a1, a2 = i1_host.to(device), i2_host.to(device)
                pcfa_attack(net, "RAFT", a1, a2, steps=1, iters=12, keep_best=False)
                r = pcfa_attack(net, "RAFT", a1, a2, steps=5, iters=12, keep_best=False)
                hs = r.history
                dts = sorted(hs[k]["t"] - hs[k - 1]["t"] for k in range(1, len(hs)))
                ncl = sorted(hs[k]["closure_evals"] - hs[k - 1]["closure_evals"] for k in range(1, len(hs)))
                out["attack"] = {"outer_steps_per_s": round(1.0 / dts[len(dts) // 2], 3), "outer_step_ms": round(1e3 * dts[len(dts) // 2], 2),
                                 "closures_per_outer_step": ncl[len(ncl) // 2], "optimizer": "pcfa_b200.lbfgs.DeviceLBFGS (torch.optim.LBFGS semantics, max_iter 10)"}
            except Exception as e:                                   # the headline line must not depend on it
                out["attack"] = {"error": repr(e)[:200]}
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(args, samples=args.cpu_samples)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------ config 5: universal perturbation
def bench_universal(args, net, device, world, rank, barrier):
    """BASELINE.json configs[4] / SURVEY 8(d) config 5: RAFT universal-joint perturbation, `--universal-pairs` pairs per
    rank evaluated as one batch, graph-captured closure, and per closure ONE all-reduce (sum, then / world) of the fused
    fp32 buffer [grad_delta (3 x 440 x 1024 = 1.35 M floats = 5.4 MB) | loss] on the compute stream
    (attack_PCFA.py:455-517 with the batch sharded over ranks; pcfa_b200.dist.pack_reduce_unpack).  Weak scaling:
    8 pairs per GPU at every N, so efficiency(N) = ms_per_closure(1) / ms_per_closure(N) = pairs_per_s(N) / (N * pairs_per_s(1))."""
    import torch.distributed as dist
    from pcfa_b200 import objective as J
    from pcfa_b200.adapter import preprocess_img
    from pcfa_b200.attack import GraphedEvaluate
    from pcfa_b200.dist import pack_reduce_unpack
    from pcfa_b200.networks.weights import synthetic_pair
    P = args.universal_pairs
    pairs = [synthetic_pair(100 + rank * P + i, args.height, args.width) for i in range(P)]
    i1 = torch.cat([p[0] for p in pairs]).to(device) / 255.0
    i2 = torch.cat([p[1] for p in pairs]).to(device) / 255.0
    padder, (a, b) = preprocess_img("RAFT", i1, i2)
    a, b = a.contiguous(), b.contiguous()
    chw = a.shape[1:]
    n = int(torch.Size(chw).numel())
    fo = J.FusedObjective(lambda x, y: net(x, y, iters=12, test_mode=True)[1], a, b,
                          torch.zeros(P, 2, args.height, args.width, device=device), mode=J.BOX_UNIVERSAL, joint=True,
                          pad=padder.top_left, eps_box=EPS_BOX, scale=255.0, delta_bound=DELTA_BOUND, mu=MU, loss="aee")
    flat = torch.zeros(n + 1, device=device)                       # [grad_delta | loss]: gradients are written in place
    delta = (0.002 * torch.randn(chw, device=device, generator=torch.Generator(device=device).manual_seed(7))).contiguous()
    g1 = flat[:n].view(chw)
    ev = GraphedEvaluate(fo, delta, None, use_graph=not args.no_graph, g1=g1)
    ar0 = [torch.cuda.Event(enable_timing=True) for _ in range(64)]
    ar1 = [torch.cuda.Event(enable_timing=True) for _ in range(64)]

    def closure(k=None):
        loss = ev()
        flat[-1:].copy_(loss.reshape(1))
        if world > 1:
            if k is not None and k < 64:
                ar0[k].record()
            dist.all_reduce(flat)
            flat.div_(world)
            if k is not None and k < 64:
                ar1[k].record()
        return flat[-1]

    steps = args.universal_steps or args.steps
    for _ in range(max(3, args.warmup)):
        closure()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for k in range(steps):
        closure(k)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    ar_us = None
    if world > 1:
        us = sorted(x.elapsed_time(y) * 1e3 for x, y in zip(ar0[:min(steps, 64)], ar1[:min(steps, 64)]))
        ar_us = round(us[len(us) // 2], 1)
    loss = float(flat[-1].item())
    gn = float(flat[:n].norm().item())
    corr_rows = None
    if rank == 0:
        # the correlation entry points at batch %d (where they are not launch-latency-bound), same instrumented eager pass as
        # the headline kernel table
        try:
            from pcfa_b200 import profiling
            peak, _ = peaks()
            tab = profiling.kernel_table(lambda: fo.evaluate(delta, None, g1), n_steps=2, B=P, C=256, H=a.shape[2] // 8, W=a.shape[3] // 8,
                                         iters=12, peak_gbs=peak, img_numel=a.numel(), flow_numel=2 * P * a.shape[2] * a.shape[3])
            corr_rows = [{k: r[k] for k in ("name", "launches_per_step", "avg_us", "algorithmic_bytes", "achieved_gbs", "frac_of_hbm_peak") if k in r}
                         for r in tab if r["name"].startswith("pcfa_corr_") and r.get("algorithmic_bytes")]
        except Exception as e:
            corr_rows = {"error": repr(e)[:200]}
    del ev, fo
    torch.cuda.empty_cache()
    return {"workload": "RAFT universal-joint perturbation, clipping, zero target, aee, 12 GRU iters, %dx%d, %d pairs per GPU "
                        "batched, %d pairs per closure in total" % (args.height, args.width, P, P * world),
            "pairs_per_gpu": P, "pairs_total": P * world, "steps": steps, "ms_per_closure": round(ms, 3),
            "closures_per_s": round(1e3 / ms, 3), "pairs_per_s": round(P * world * 1e3 / ms, 2),
            "allreduce": "none (1 rank)" if world == 1 else "NCCL all_reduce(sum) of [grad_delta | loss] = %d bytes per closure on the compute stream, then 1/world" % (4 * (n + 1)),
            "allreduce_bytes": 4 * (n + 1), "allreduce_us_median": ar_us, "scaling": "weak",
            "efficiency_definition": "ms_per_closure(N=1) / ms_per_closure(N), both at %d pairs per GPU" % P,
            "loss": loss, "grad_norm": gn, "cuda_graph": not args.no_graph, "corr_kernels_at_batch": corr_rows}


# ------------------------------------------------------------------------------------ CPU arm
def cpu_closure_factory(args):
    """The reference closure (attack_PCFA.py:175-189) restated with torch CPU ops: oracle port.  The network class is this
    repo's own state-dict-compatible definition (pcfa_b200/networks/raft.py) with the oracle's torch-op CorrBlock injected —
    a port of the reference, not the reference: it skips the reference's dead per-iteration mask-head / up-sampling work
    in test_mode (models/raft/raft.py:128-137), so it is slightly FASTER than the reference itself would be."""
    from oracle import torch_ref as TR
    from pcfa_b200.adapter import InputPadder, build_network
    from pcfa_b200.networks.weights import synthetic_pair
    net = build_network("RAFT", device="cpu", seed=0, ops=TR)
    i1, i2 = synthetic_pair(0, args.height, args.width)
    i1, i2 = i1 / 255.0, i2 / 255.0
    padder = InputPadder(i1.shape)
    i1, i2 = padder.pad(i1, i2)
    target = torch.zeros(1, 2, args.height, args.width)
    w1 = init_vars(i1).requires_grad_(True)
    w2 = init_vars(i2).requires_grad_(True)

    def closure():
        w1.grad = None
        w2.grad = None
        x1 = TR.scaled_input(w1, var_change=True, eps_box=EPS_BOX, make_unit_input=True)
        x2 = TR.scaled_input(w2, var_change=True, eps_box=EPS_BOX, make_unit_input=True)
        flow = padder.unpad(net(x1, x2, iters=12, test_mode=True)[1])
        d1, d2 = TR.extract_deltas(w1, w2, i1, i2, "change_of_variables", eps_box=EPS_BOX)
        loss = TR.loss_delta_constraint(flow, target, d1, d2, None, delta_bound=DELTA_BOUND, mu=MU, f_type="aee")
        loss.backward()
        return float(loss)
    return closure


def use_all_host_cores():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use the box's cores at every N."""
    n = os.cpu_count() or 1
    torch.set_num_threads(n)
    return torch.get_num_threads()


def cpu_baseline(args, samples=3):
    use_all_host_cores()
    closure = cpu_closure_factory(args)
    closure()
    ts = []
    for _ in range(samples):
        t0 = time.perf_counter()
        closure()
        ts.append(time.perf_counter() - t0)
    med = statistics.median(ts)
    return {"value": round(1.0 / med, 4), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": "%d full-size closure evaluations (RAFT 12 iters, %dx%d, fwd+bwd) after 1 warm-up, median; "
                      "torch CPU ops restating the reference closure (oracle/torch_ref.py), nproc=%d"
                      % (samples, args.height, args.width, os.cpu_count())}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    use_all_host_cores()
    closure = cpu_closure_factory(args)
    for _ in range(max(1, args.warmup)):
        closure()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        closure()
    dt = time.perf_counter() - t0
    v = args.steps / dt
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": round(v, 4), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": max(1, args.warmup), "ms_per_step": round(dt / args.steps * 1e3, 2),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, 1),
        "cpu_baseline": {"value": round(v, 4), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "sample": "each step = one full-size closure evaluation on the host cores (oracle/torch_ref.py "
                                   "restating the reference's torch CPU path; /root/reference is not on this box)"},
        "e2e": {"value": round(v, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device — the pcfa_b200 arm has no CPU fallback "
                             "(use --impl reference for the CPU baseline)")
        run_b200(a)
