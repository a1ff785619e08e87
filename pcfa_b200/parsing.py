"""Command-line surface of the PCFA drivers — the flags, choices and defaults of the reference's
helper_functions/parsing_file.py:3-98 (stage 'training', attack 'pcfa'), plus a small group of options
that only exist here because this build runs offline (synthetic pairs, weight files, GRU iterations)."""
from __future__ import annotations

import argparse


def create_parser(stage="training", attack_type="pcfa"):
    stage, attack_type = stage.lower(), attack_type.lower()
    if stage not in ['training', 'evaluation']:
        raise ValueError('To create a parser the stage has to be specified. Please choose one of "training" or "evaluation"')
    if attack_type not in ["pcfa"]:
        raise ValueError('This build implements the "pcfa" attack only')
    p = argparse.ArgumentParser(usage='%(prog)s [options (see below)]')
    g = p.add_argument_group(title='network arguments')
    # the reference also lists SpyNet (its default): it has no cost volume and is not part of this build
    g.add_argument('--net', default='RAFT', choices=['RAFT', 'GMA', 'PWCNet', 'FlowNet2'],
                   help="specify the network under attack (SpyNet is not implemented in this build)")
    g = p.add_argument_group(title="dataset arguments")
    g.add_argument('--dataset', default='Kitti15', choices=['Kitti15', 'Sintel'])
    g.add_argument('--dataset_stage', default='evaluation', choices=['training', 'evaluation'])
    g.add_argument('--small_run', action='store_true')
    g.add_argument('--dstype', default='final', choices=['clean', 'final'])
    g = p.add_argument_group(title="data saving arguments")
    g.add_argument('--output_folder', default='experiment_data')
    g.add_argument('--small_save', action='store_true')
    g.add_argument('--save_frequency', type=int, default=1)
    g.add_argument('--no_save', action='store_true')
    g.add_argument('--unregistered_artifacts', action='store_true', default=False)
    g = p.add_argument_group(title="global distortion attack arguments")
    g.add_argument('--joint_perturbation', action='store_true', default=False)
    g.add_argument('--steps', default=20, type=int)
    g = p.add_argument_group(title="pcfa arguments")
    g.add_argument('--universal_perturbation', action='store_true', default=False)
    g.add_argument('--boxconstraint', default='change_of_variables', choices=['clipping', 'change_of_variables'])
    g.add_argument('--batch_size', default=4, type=int)
    if stage == "training":
        g.add_argument('--delta_bound', default=0.005, type=float)
        g.add_argument('--mu', default=-1, type=float)
        g.add_argument('--epochs', default=25, type=int)
        t = p.add_argument_group(title="training arguments")
        t.add_argument('--target', default='zero', choices=['zero', 'neg_flow', 'custom'])
        t.add_argument('--custom_target_path', default='')
        t.add_argument('--loss', default='aee', choices=['aee', 'mse', 'cosim'])
    else:
        g.add_argument('--perturbation_sourcefolder')
        g.add_argument('--origin_net')
    o = p.add_argument_group(title="offline-build arguments (not in the reference)")
    o.add_argument('--num_pairs', type=int, default=None,
                   help="number of synthetic pairs of the dataset's shape (default 4; 32 with --small_run)")
    o.add_argument('--weights', default=None, help="checkpoint file; default: deterministic synthetic weights")
    o.add_argument('--no_cuda_graph', action='store_true', help="launch the closure eagerly")
    return p
