#!/usr/bin/env python3
"""Developer tool: one search with the default schedule, with the bulk round forced (MetisShard.reserved = 1) and with
the chain kernel alone (reserved = 2^31 - 1), for the shard of rank 0 of 1 / 2 / 4 / 8 ranks; CUDA-event times.
python tools/mode_compare.py workload [workload ...]"""
import itertools, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from metis_b200 import flatten, search
from metis_b200.data_loader import ProfileDataLoader
from metis_b200.gpu_cluster import GPUCluster
from metis_b200.utils import ModelConfig
from metis_b200.workloads import WORKLOADS, materialize, profile_file_order

for name in sys.argv[1:] or ['c3_homo64_mpl6', 'c4_het128']:
    w = WORKLOADS[name]
    tmp = tempfile.mkdtemp(); materialize(w, tmp)
    cluster = GPUCluster(tmp + '/hostfile', tmp + '/clusterfile.json')
    profile, _ = ProfileDataLoader(tmp + '/profile', profile_file_order(w)).load_profile_data_all()
    cfg = ModelConfig('SYN', w.num_layers, w.sequence_length, w.vocab_size, w.hidden_size, 32)
    seqs = list(itertools.permutations(w.device_types()))
    problem = flatten.build_problem(profile, cluster, cfg, w.gbs, w.max_tp, w.max_bs, seqs)
    space = flatten.build_plan_space(len(seqs), cluster.get_total_num_devices(), w.gbs, w.num_layers, w.variance,
                                     w.max_permute_len, device_rows=True)
    dp = search.DeviceProblem(problem, space, 'cuda:0')
    for world in (1, 2, 4, 8):
        for label, reserved in (('default', 0), ('bulk round forced', 1), ('chain kernel only', 2 ** 31 - 1)):
            s = search.HetSearcher(dp, 0, world, want_records=False)
            s.shard.reserved = reserved
            for _ in range(2):
                s.launch()
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(3):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); s.launch(); b.record(); torch.cuda.synchronize()
                best = min(best, a.elapsed_time(b))
            sm = s.summary()
            print(f'{name:18s} world {world}  {label:18s} {best:9.3f} ms   admitted {int(sm.reserved[0])} '
                  f'chained {int(sm.reserved[1])} records {int(sm.num_records)}')
