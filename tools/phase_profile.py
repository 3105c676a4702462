#!/usr/bin/env python3
"""Developer tool: run one search with the phase-clock build (libmetis_b200_prof.so, compiled with
-DMETIS_PROFILE_PHASES) and print the share of warp-cycles per phase of search_loop."""
import itertools
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from metis_b200 import flatten, native, search  # noqa: E402
from metis_b200.data_loader import ProfileDataLoader  # noqa: E402
from metis_b200.gpu_cluster import GPUCluster  # noqa: E402
from metis_b200.utils import ModelConfig  # noqa: E402
from metis_b200.workloads import WORKLOADS, materialize, profile_file_order  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else 'c3_homo64_mpl6'
prof = os.path.join(os.path.dirname(native.LIB_PATH), os.environ.get('METIS_LIB', 'libmetis_b200_prof.so'))
native._lib = native.load_library(prof)
w = WORKLOADS[name]
tmp = tempfile.mkdtemp()
materialize(w, tmp)
cluster = GPUCluster(tmp + '/hostfile', tmp + '/clusterfile.json')
profile, _ = ProfileDataLoader(tmp + '/profile', profile_file_order(w)).load_profile_data_all()
cfg = ModelConfig('SYN', w.num_layers, w.sequence_length, w.vocab_size, w.hidden_size, 32)
seqs = list(itertools.permutations(w.device_types()))
problem = flatten.build_problem(profile, cluster, cfg, w.gbs, w.max_tp, w.max_bs, seqs)
space = flatten.build_plan_space(len(seqs), cluster.get_total_num_devices(), w.gbs, w.num_layers, w.variance,
                                 w.max_permute_len, native._lib)
dp = search.DeviceProblem(problem, space, 'cuda:0')
dp.lib = native._lib
s = search.HetSearcher(dp, want_records=False)
s.shard.reserved = int(os.environ.get('METIS_BULK_MIN', '0'))
for _ in range(3):
    s.launch()
torch.cuda.synchronize()
import ctypes
marks = (ctypes.c_longlong * 64)()
native._lib.metis_debug_marks(None, 1)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); s.launch(); b.record(); torch.cuda.synchronize()
sm = s.summary()
native._lib.metis_debug_marks(marks, 0)
names = {0: 'fetch+decode', 1: 'begin', 2: 'P perf', 10: 'R forward', 11: 'R backward', 12: 'R leftovers', 13: 'R vote',
         14: 'R cnt+capa', 15: 'R adjust', 16: 'R part', 20: 'M demand', 21: 'M reweight', 22: 'C stage terms',
         23: 'C sums+emit', 24: 'chain advance'}
mt = sum(marks[:32]) or 1
print(f'{name}: {a.elapsed_time(b):.2f} ms (profiling build), plans {space.num_plans}, admitted {sm.reserved[0]}, chained '
      f'{sm.reserved[1]}, B {sm.num_partition_calls}, runs {sm.num_balancer_runs}, C {sm.num_records}')
print('  chain kernel, leader-lane cycles per phase: ' + ', '.join(
    f'{names.get(i, i)} {100.0 * marks[i] / mt:.1f}%' for i in range(32) if marks[i]))
print(f'  total {mt / 1e6:.1f} Mcycles over all chain warps')
