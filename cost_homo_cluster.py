#!/usr/bin/env python3
"""Drop-in CLI for the reference's cost_homo_cluster.py (same flags, ``rank, cost, plan`` table).
The reference's own __main__ cannot run as shipped (cost_homo_cluster.py:44,49: an assert on
bandwidth units and two names that do not exist); this one builds the same objects without them."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from metis_b200.api import GPTActivationAndParam, HomoCostEstimator, cost_homo_cluster  # noqa: E402
from metis_b200.arguments import parse_args  # noqa: E402
from metis_b200.data_loader import ProfileDataLoader  # noqa: E402
from metis_b200.gpu_cluster import GPUCluster  # noqa: E402
from metis_b200.utils import ModelConfig  # noqa: E402


def main(argv=None):
    args = parse_args(argv)
    gpu_cluster = GPUCluster(hostfile_path=args.hostfile_path, clusterfile_path=args.clusterfile_path)
    profile_data, device_types = ProfileDataLoader(args.profile_data_path).load_profile_data_all()
    if len(profile_data.keys()) > 0:
        print('\nProfiled data has been loaded.')
    assert len(profile_data.keys()) > 0, 'There is no profiled data at the specified path.'

    model_config = ModelConfig(model_name=args.model_name, num_layers=args.num_layers,
                               sequence_length=args.sequence_length, vocab_size=args.vocab_size,
                               hidden_size=args.hidden_size, attention_head_size=args.attention_head_size)
    model_volume = GPTActivationAndParam(model_config, profile_data['model']['parameters'])
    cost_estimator = HomoCostEstimator(profile_data, model_config, model_volume, gpu_cluster)

    estimate_costs = cost_homo_cluster(args, gpu_cluster, cost_estimator, device_types[0])
    ranked = sorted(estimate_costs, key=lambda kv: kv[1])
    print('rank, cost, plan')
    for idx, r in enumerate(ranked):
        print(f'{idx + 1}, {r[1]}, {r[0]}')
    return ranked


if __name__ == '__main__':
    main()
