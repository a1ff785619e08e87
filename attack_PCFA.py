#!/usr/bin/env python
"""attack_PCFA.py — same command line as the reference driver (attack_PCFA.py:705-713,
helper_functions/parsing_file.py), running the B200 closure.

    python attack_PCFA.py --net RAFT --delta_bound 0.005 [--joint_perturbation --boxconstraint clipping]
                          [--universal_perturbation] --target zero --loss aee
    torchrun --nproc-per-node 8 attack_PCFA.py ...      # pairs (or the universal batch) sharded over ranks

Datasets cannot be reached offline, so `--dataset` selects the SHAPE of synthetic pairs (Sintel
436x1024, Kitti15 375x1242 as enforced by helper_functions/datasets.py:185-187).  Perturbations are
written as .npy with the reference's naming (helper_functions/logging.py:265-286) unless --no_save.
"""
from __future__ import annotations

import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))

from pcfa_b200.adapter import build_network  # noqa: E402
from pcfa_b200.attack import UniversalAttack, pcfa_attack, resolve_mu  # noqa: E402
from pcfa_b200.networks.weights import synthetic_pair  # noqa: E402
from pcfa_b200.parsing import create_parser  # noqa: E402

SHAPES = {"Sintel": (436, 1024), "Kitti15": (375, 1242)}


def _dist_setup():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return world, rank, torch.device("cuda", local)


def _save(folder, batch, name, t):
    if folder is not None and t is not None:
        np.save(folder / ("%05d_%s.npy" % (batch, name)), t.detach().cpu().numpy())


def _header(args, mu, folder):
    print("\nStarting Perturbation Constrained Flow Attack (PCFA):\n")
    print("\tModel:                   %s" % args.net)
    print("\tPerturbation universal:  %s" % str(args.universal_perturbation))
    print("\tPerturbation joint:      %s" % str(args.joint_perturbation))
    print("\tPerturbation bound:      %f\n" % args.delta_bound)
    print("\tTarget:                  %s" % args.target)
    print("\tOptimizer steps:         %d" % args.steps)
    print("\tOptimizer boxconstraint: %s" % ("clipping" if args.universal_perturbation else args.boxconstraint))
    print("\tOptimizer mu:            %f\n" % mu)
    print("\tOutputfolder:            %s\n" % folder)


def main(argv=None):
    args = create_parser('training', 'pcfa').parse_args(argv)
    if not torch.cuda.is_available():
        raise SystemExit("attack_PCFA.py needs a CUDA device: pcfa_b200 has no CPU path")
    world, rank, device = _dist_setup()
    mu = resolve_mu(args.mu, args.delta_bound, args.target)
    H, W = SHAPES[args.dataset]
    n_pairs = args.num_pairs or (32 if args.small_run else 4)
    tag = "%s_PCFA_%s_%s" % (args.net, "cd" if args.joint_perturbation else "dd", "u" if args.universal_perturbation else "-")
    folder = None
    if not args.no_save:
        stamp = [time.strftime("%Y-%m-%d_%H:%M:%S")]
        if world > 1:                                                # one folder name for all ranks (rank 0's clock)
            import torch.distributed as dist
            dist.broadcast_object_list(stamp, src=0)
        folder = Path(args.output_folder) / tag / stamp[0] / "patches"
        folder.mkdir(parents=True, exist_ok=True)
    if rank == 0:
        print(args)
        _header(args, mu, folder)
    model = build_network(args.net, device=device, weights=args.weights)
    t0 = time.time()
    if args.universal_perturbation:
        ua = UniversalAttack(model, args.net, (H, W), device, delta_bound=args.delta_bound, mu=args.mu, target=args.target,
                             loss=args.loss, joint_perturbation=args.joint_perturbation,
                             custom_target_path=args.custom_target_path)
        per_rank = max(1, args.batch_size // world)
        n_batches = max(1, n_pairs // (per_rank * world))
        for epoch in range(args.epochs):
            for b in range(n_batches):
                base = (b * world + rank) * per_rank
                pairs = [synthetic_pair(base + i, H, W) for i in range(per_rank)]
                i1 = torch.cat([p[0] for p in pairs]).to(device)
                i2 = torch.cat([p[1] for p in pairs]).to(device)
                stats = ua.run_batch(i1, i2, args.steps)
            l2 = ua.l2_norms()
            if rank == 0:
                print("epoch %d: AEE(f_adv,f_targ)=%f AEE(f_adv,f_init)=%f L2=%f" % (epoch, stats[-1][0], stats[-1][1], l2[2]))
                if folder is not None:
                    _save(folder, epoch, "delta1_e%d" % epoch, ua.delta1)
                    _save(folder, epoch, "delta2_e%d" % epoch, ua.delta1 if ua.delta2 is None else ua.delta2)
        evals = ua.closure_evals
    else:
        sums = np.zeros(3)
        count, evals = 0, 0
        for idx in range(rank, n_pairs, world):                      # pair sharding, no communication
            i1, i2 = synthetic_pair(idx, H, W)
            r = pcfa_attack(model, args.net, i1.to(device), i2.to(device), steps=args.steps, delta_bound=args.delta_bound,
                            mu=args.mu, target=args.target, loss=args.loss, joint_perturbation=args.joint_perturbation,
                            boxconstraint=args.boxconstraint, custom_target_path=args.custom_target_path,
                            use_graph=not args.no_cuda_graph)
            sums += (r.aee_adv_pred_min, r.aee_adv_tgt_min, r.l2_delta12_min)
            count += 1
            evals += r.closure_evals
            if folder is not None and idx % args.save_frequency == 0:
                _save(folder, idx, "delta1_best", r.delta1_best)
                _save(folder, idx, "delta2_best", r.delta2_best)
                _save(folder, idx, "flow_pred_best", r.flow_best)
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor(list(sums) + [count, evals], device=device, dtype=torch.float64)
            dist.all_reduce(t)                                       # final metric gather only
            sums, count, evals = t[:3].cpu().numpy(), int(t[3]), int(t[4])
        if rank == 0:
            print("\nFinished attacking with PCFA. The best achieved values are")
            print("\tAEE(f_adv, f_init)=%f" % (sums[0] / count))
            print("\tAEE(f_adv, f_targ)=%f" % (sums[1] / count))
            print("\tL2(perturbation)  =%f\n" % (sums[2] / count))
    torch.cuda.synchronize()
    if rank == 0:
        dt = time.time() - t0
        print("%d closure evaluations in %.2f s (%.1f closures/s over %d GPU(s))" % (evals, dt, evals / dt, world))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
