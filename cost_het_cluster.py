#!/usr/bin/env python3
"""Drop-in CLI for the reference's cost_het_cluster.py (same flags, same ranked stdout table);
the search itself runs on the GPU (metis_b200.api.cost_het_cluster)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from metis_b200.api import (GPTActivationAndParam, HeteroCostEstimator, LayerLoadBalancer,  # noqa: E402
                            cost_het_cluster)
from metis_b200.arguments import parse_args  # noqa: E402
from metis_b200.data_loader import ProfileDataLoader  # noqa: E402
from metis_b200.gpu_cluster import GPUCluster  # noqa: E402
from metis_b200.utils import ModelConfig  # noqa: E402


def main(argv=None):
    args = parse_args(argv)
    gpu_cluster = GPUCluster(hostfile_path=args.hostfile_path, clusterfile_path=args.clusterfile_path)
    profile_data, _ = ProfileDataLoader(args.profile_data_path).load_profile_data_all()
    print(profile_data)
    assert len(profile_data.keys()) > 0, 'There is no profiled data at the specified path.'

    model_config = ModelConfig(model_name=args.model_name, num_layers=args.num_layers,
                               sequence_length=args.sequence_length, vocab_size=args.vocab_size,
                               hidden_size=args.hidden_size, attention_head_size=args.attention_head_size)
    model_volume = GPTActivationAndParam(model_config, profile_data['model']['parameters'])
    cost_estimator = HeteroCostEstimator(profile_data, model_config, model_volume, gpu_cluster)
    layer_load_balancer = LayerLoadBalancer(gpu_cluster, profile_data, model_config, args.gbs)

    start_time = time.time()
    estimate_costs = cost_het_cluster(args, gpu_cluster, profile_data, model_config, cost_estimator,
                                      layer_load_balancer)
    print(f'search_time: {time.time() - start_time}s')
    print(f'len(costs): {len(estimate_costs)}')
    ranked = estimate_costs.ranked()     # = sorted(estimate_costs, key=lambda kv: kv[6]), order from the device sort
    print('rank, cost, node_sequence, device_groups, strategies(dp_deg, tp_deg), batches(number of batch), '
          'layer_partition')
    for idx, r in enumerate(ranked):
        print(f'{idx + 1}, {r[6]}, {r[0]}, {r[1]}, {r[2]}, {r[3]}, {r[4]}')
    return ranked


if __name__ == '__main__':
    main()
