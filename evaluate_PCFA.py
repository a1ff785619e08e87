#!/usr/bin/env python
"""Evaluate stored PCFA perturbations with a (possibly different) network — the reference's evaluate_PCFA.py command
line (flags of helper_functions/parsing_file.py: --net --origin_net --perturbation_sourcefolder --joint_perturbation
--universal_perturbation --dataset --batch_size ...) on synthetic pairs of the dataset's shape (no data offline).
Prints, per stored epoch, the reference's closing lines (evaluate_PCFA.py:296-298)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
from pcfa_b200.adapter import build_network  # noqa: E402
from pcfa_b200.evaluate import (convert_perturbationsizes, evaluate_perturbation, extract_epoch_patchlist, l2_metrics,  # noqa: E402
                                load_delta)
from pcfa_b200.networks.weights import synthetic_pair  # noqa: E402
from pcfa_b200.parsing import create_parser  # noqa: E402

SHAPES = {"Sintel": (436, 1024), "Kitti15": (375, 1242)}


def main(argv=None):
    args = create_parser('evaluation', 'pcfa').parse_args(argv)
    if not torch.cuda.is_available():
        raise SystemExit("evaluate_PCFA.py needs a CUDA device: pcfa_b200 has no CPU path")
    if args.origin_net is None:
        raise ValueError("args.origin_net is not allowed to be empty. Please state which network was used to train the "
                         "perturbations via the --origin_net argument.")
    device = torch.device("cuda")
    H, W = SHAPES[args.dataset]
    print("Evaluating a Perturbation Constrained Flow Attack:\n")
    print("\tModel (evaluation, now): %s" % args.net)
    print("\tModel (training):        %s" % args.origin_net)
    print("\tPerturbation universal:  %s" % str(args.universal_perturbation))
    print("\tPerturbation joint:      %s\n" % str(args.joint_perturbation))
    epochs, d1_paths, d2_paths = extract_epoch_patchlist(args.perturbation_sourcefolder)
    model = build_network(args.net, device=device, weights=args.weights)
    n_pairs = args.num_pairs or (32 if args.small_run else 4)
    bs = max(1, args.batch_size)

    def batches():
        for s in range(0, n_pairs, bs):
            pairs = [synthetic_pair(i, H, W) for i in range(s, min(s + bs, n_pairs))]
            yield torch.cat([p[0] for p in pairs]).to(device), torch.cat([p[1] for p in pairs]).to(device)

    for epoch in range(epochs):
        print("Evaluation for perturbation from epoch %d" % epoch)
        delta1 = convert_perturbationsizes(load_delta(d1_paths[epoch], device), (H, W), args.origin_net, args.net)
        if args.universal_perturbation or not d2_paths:
            delta2 = delta1
        else:
            delta2 = convert_perturbationsizes(load_delta(d2_paths[epoch], device), (H, W), args.origin_net, args.net)
        res = evaluate_perturbation(model, args.net, delta1, delta2, batches(), joint=args.joint_perturbation,
                                    boxconstraint=args.boxconstraint)
        l2 = l2_metrics(delta1, delta2)
        print("Finished attacking epoch %d" % epoch)
        print("\tAEE(f_adv, f_init)=%f" % res["aee_adv_pred"])
        print("\tL2(perturbation)  =%f\n" % l2[2])


if __name__ == '__main__':
    main()
