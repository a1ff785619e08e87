"""Closure time and outer-step time for the other configurations SURVEY.md section 8(d) lists (bench.py is config 2):
  RAFT disjoint 436x1024 | GMA joint 436x1024 (6 iterations) | PWCNet disjoint 375x1242 | FlowNet2 disjoint 375x1242 |
  RAFT universal-joint, 8 pairs batched on one GPU.
Closure: CUDA events around graph replays (median of 20 after 5 warm-ups).  Outer step: wall clock of pcfa_attack /
UniversalAttack over 2 outer L-BFGS steps (torch.optim.LBFGS on the host, max_iter 10) divided by 2."""
import json, statistics, sys, time
import torch
sys.path.insert(0, '.')
from pcfa_b200 import _lib, objective as J
from pcfa_b200.adapter import build_network, preprocess_img, model_takes_unit_input
from pcfa_b200.attack import GraphedEvaluate, UniversalAttack, _net_forward, pcfa_attack, resolve_mu
from pcfa_b200.networks.weights import synthetic_pair

_lib.load()
dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
only = sys.argv[1:] or ["RAFT", "GMA", "GMA-batch8", "PWCNet", "FlowNet2", "RAFT-universal8"]
rows = []


def closure_ms(ev, n=20, warm=5):
    for _ in range(warm):
        ev()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ev(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


for name in only:
    torch.cuda.empty_cache()
    net_name = name.split("-")[0]
    H, W = (436, 1024) if net_name in ("RAFT", "GMA") else (375, 1242)
    model = build_network(net_name, device=dev, seed=0, gain=0.5)
    row = dict(config=name, shape=[H, W])
    try:
        if name.endswith("universal8"):
            pairs = [synthetic_pair(i, H, W) for i in range(8)]
            i1 = torch.cat([p[0] for p in pairs]).to(dev); i2 = torch.cat([p[1] for p in pairs]).to(dev)
            ua = UniversalAttack(model, net_name, (H, W), dev, joint_perturbation=True, use_graph=True)
            ua.run_batch(i1, i2, 1)                                  # builds and warms everything
            def run(steps):
                torch.cuda.synchronize(); t0 = time.perf_counter(); n0 = ua.closure_evals
                ua.run_batch(i1, i2, steps)
                torch.cuda.synchronize()
                return time.perf_counter() - t0, ua.closure_evals - n0
            # every run_batch call rebuilds the objective and re-captures the closure graph (~1 s): difference a long and a
            # short run so that this set-up cancels (with 4 vs 1 steps its jitter could exceed the three steps' time)
            run(2)
            t1, n1 = min((run(1) for _ in range(3)), key=lambda r: r[0])
            t9, n9 = min((run(9) for _ in range(2)), key=lambda r: r[0])
            row.update(outer_step_s=(t9 - t1) / 8, closures_per_outer_step=(n9 - n1) / 8,
                       ms_per_closure_incl_host=1e3 * (t9 - t1) / max(1, n9 - n1), pairs=8)
        elif name.endswith("batch8"):
            # BASELINE config 3 at N = 1: the eight Sintel-shaped pairs of the batch evaluated as ONE batched joint closure
            # (per-pair delta, clipping); closure time only
            pairs = [synthetic_pair(i, H, W) for i in range(8)]
            a = torch.cat([p[0] for p in pairs]).to(dev) / 255.; b = torch.cat([p[1] for p in pairs]).to(dev) / 255.
            padder, (a, b) = preprocess_img(net_name, a, b)
            a, b = a.contiguous(), b.contiguous()
            fo = J.FusedObjective(_net_forward(model, net_name, None), a, b, torch.zeros(8, 2, H, W, device=dev), mode=J.box_mode("clipping", joint=True),
                                  joint=True, pad=padder.top_left, eps_box=1e-7, scale=255.0, delta_bound=0.005,
                                  mu=resolve_mu(-1., 0.005, "zero"), loss="aee")
            ev = GraphedEvaluate(fo, torch.zeros_like(a), None, use_graph=True)
            row.update(closure_ms=closure_ms(ev, n=10, warm=3), pairs=8)
            row["ms_per_pair"] = row["closure_ms"] / 8
            del ev, fo
        else:
            i1, i2 = synthetic_pair(0, H, W)
            i1, i2 = i1.to(dev), i2.to(dev)
            joint = net_name == "GMA"
            box = "clipping" if joint else "change_of_variables"
            iters = 12 if net_name == "RAFT" else None
            # closure alone
            unit = model_takes_unit_input(net_name)
            a, b = (i1, i2) if unit else (i1 / 255., i2 / 255.)
            padder, (a, b) = preprocess_img(net_name, a, b)
            a, b = a.contiguous(), b.contiguous()
            mode = J.box_mode(box, joint=joint)
            fo = J.FusedObjective(_net_forward(model, net_name, iters), a, b, torch.zeros(1, 2, H, W, device=dev), mode=mode,
                                  joint=joint, pad=padder.top_left, eps_box=1e-7, scale=1.0 if unit else 255.0,
                                  delta_bound=0.005, mu=resolve_mu(-1., 0.005, "zero"), loss="aee")
            if joint:
                v1, v2 = torch.zeros_like(a), None
            else:
                v1 = torch.atanh(2. * (1. - 1e-7) * a - (1 - 1e-7)).contiguous(); v2 = torch.atanh(2. * (1. - 1e-7) * b - (1 - 1e-7)).contiguous()
            ev = GraphedEvaluate(fo, v1, v2, use_graph=True)
            row.update(closure_ms=closure_ms(ev), launches_per_closure=None)
            n0 = _lib.launch_count(); fo.evaluate(v1, v2, ev.g1, ev.g2); torch.cuda.synchronize()
            row["pcfa_launches_per_closure"] = _lib.launch_count() - n0
            del ev, fo
            # outer steps (host L-BFGS around graph replays): wall clock between the per-step metric syncs of ONE
            # 5-step run (pcfa_attack stamps every outer step after its D2H of the metrics); the first step is dropped
            pcfa_attack(model, net_name, i1, i2, steps=2, joint_perturbation=joint, boxconstraint=box, iters=iters, keep_best=False)
            r = pcfa_attack(model, net_name, i1, i2, steps=5, joint_perturbation=joint, boxconstraint=box, iters=iters, keep_best=False)
            hs = r.history
            dts = [hs[k]["t"] - hs[k - 1]["t"] for k in range(1, len(hs))]
            ncl = [hs[k]["closure_evals"] - hs[k - 1]["closure_evals"] for k in range(1, len(hs))]
            row.update(outer_step_s=statistics.median(dts), closures_per_outer_step=statistics.median(ncl),
                       host_overhead_frac=round(1.0 - statistics.median(ncl) * row["closure_ms"] * 1e-3 / statistics.median(dts), 3))
    except Exception as e:                                           # keep going: one config must not hide the others
        row["error"] = repr(e)[:300]
    row["max_mem_GB"] = round(torch.cuda.max_memory_allocated() / 2**30, 2)
    row["max_reserved_GB"] = round(torch.cuda.max_memory_reserved() / 2**30, 2)
    torch.cuda.reset_peak_memory_stats()
    print(row, flush=True)
    rows.append(row)
    del model
json.dump(rows, open("gpurun_out/bench_configs.json", "w"), indent=1)
