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
s.shard.reserved = int(os.environ.get('METIS_COOP', '0'))
for _ in range(3):
    s.launch()
torch.cuda.synchronize()
import ctypes
marks = (ctypes.c_longlong * 64)()
native._lib.metis_debug_marks(None, 1)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); s.launch(); b.record(); torch.cuda.synchronize()
sm = s.summary()
cyc = list(sm.reserved)[:5]
tot = sum(cyc) or 1
native._lib.metis_debug_marks(marks, 0)
names = {0: 'between tasks', 1: 'restore', 2: 'P perf', 10: 'R fwd', 11: 'R bwd', 12: 'R leftovers', 13: 'R vote', 14: 'R cnt+capa', 15: 'R adjust', 16: 'R part', 20: 'M demand', 21: 'M reweight', 22: 'C cost', 23: 'chain', 24: 'save'}
mt = sum(marks[:32]) or 1
print('  cooperative-mode marks (leader-lane cycles): ' + ', '.join(f'{names.get(i, i)} {100.0 * marks[i] / mt:.1f}%' for i in range(32) if marks[i]))
print(f'{name}: {a.elapsed_time(b):.2f} ms, plans {space.num_plans}, B {sm.num_partition_calls}, runs {sm.num_balancer_runs}, C {sm.num_records}')
import numpy as np  # noqa: E402
base = (s.workspace.data_ptr() + 127) & ~127
off = base - s.workspace.data_ptr()
trace = s.workspace[off + 4096: off + 4096 + 512 * 16].cpu().numpy().view(np.uint64).reshape(-1, 2)
rows = [(int(n), int(t)) for n, t in trace if t]
if rows:
    t0 = rows[0][1]
    print('  round: tasks  start_us  (duration_us)')
    for i, (n, t) in enumerate(rows):
        dur = (rows[i + 1][1] - t) / 1e3 if i + 1 < len(rows) else 0.0
        print(f'  {i + 1:3d}: {n:7d} {(t - t0) / 1e3:9.1f}  ({dur:8.1f})')
for k, n in zip(cyc, ['F fetch/advance', 'P performance', 'R balance_run', 'M memory/adjust', 'C cost/emit']):
    print(f'  {n:18s} {100.0 * k / tot:6.2f} %   {k / 1e6:10.1f} Mcycles (summed over warps)')
if any(marks[32:]):
    print('  largest lane skew (cycles) at marks: ' + ', '.join(f'{names.get(i, i)} {marks[32 + i]}' for i in range(32) if marks[32 + i]))
elif os.environ.get('METIS_LIB', '').find('skew') >= 0:
    print('  lane skew: 0 cycles at every mark (all lanes read the same clock value)')
