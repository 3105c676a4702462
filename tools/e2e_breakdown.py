#!/usr/bin/env python3
"""Developer tool: where the end-to-end time of one search goes (host enumeration, staging, H2D, kernels,
record sort, D2H), each stage closed by a device synchronisation.  python tools/e2e_breakdown.py [workload]"""
import itertools
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from metis_b200 import flatten, native, search  # noqa: E402
from metis_b200.data_loader import ProfileDataLoader  # noqa: E402
from metis_b200.gpu_cluster import GPUCluster  # noqa: E402
from metis_b200.utils import ModelConfig  # noqa: E402
from metis_b200.workloads import WORKLOADS, materialize, profile_file_order  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else 'c3_homo64_mpl6'
w = WORKLOADS[name]
tmp = tempfile.mkdtemp()
materialize(w, tmp)
cluster = GPUCluster(tmp + '/hostfile', tmp + '/clusterfile.json')
profile, _ = ProfileDataLoader(tmp + '/profile', profile_file_order(w)).load_profile_data_all()
cfg = ModelConfig('SYN', w.num_layers, w.sequence_length, w.vocab_size, w.hidden_size, 32)
seqs = list(itertools.permutations(w.device_types()))
problem = flatten.build_problem(profile, cluster, cfg, w.gbs, w.max_tp, w.max_bs, seqs)
ndev = cluster.get_total_num_devices()
DEVICE_ROWS = os.environ.get('METIS_HOST_ROWS', '') in ('', '0')      # default: the GPU writes the rows (8(f)-1)
space = flatten.build_plan_space(len(seqs), ndev, w.gbs, w.num_layers, w.variance, w.max_permute_len,
                                 device_rows=DEVICE_ROWS)
dp = search.DeviceProblem(problem, space, 'cuda:0')
full = search.HetSearcher(dp, want_records=True)
stream = torch.cuda.current_stream()
full.run(stream)
acc = {}
N = 6
for it in range(N):
    torch.cuda.synchronize()
    marks = [('start', time.perf_counter())]

    def mark(label):
        torch.cuda.synchronize()
        marks.append((label, time.perf_counter()))
    sp2 = flatten.build_plan_space(len(seqs), ndev, w.gbs, w.num_layers, w.variance, w.max_permute_len,
                                   rows_out=None if DEVICE_ROWS else dp.staging('rows'), device_rows=DEVICE_ROWS)
    mark('host enumeration')
    dp.restage_space(sp2)
    mark('restage')
    dp.upload(stream)
    mark('H2D + row kernel' if DEVICE_ROWS else 'H2D')
    full.launch(stream)
    mark('search kernels')
    n = int(full.summary().num_records)
    full.sort_records(n, native.SORT_POSITION, stream)
    mark('record sort kernel')
    rec = full.records[:2 * n].cpu().numpy()
    mark('D2H of all records')
    if it:
        for (_, a), (lbl, b) in zip(marks, marks[1:]):
            acc[lbl] = acc.get(lbl, 0.0) + (b - a)
for k, v in acc.items():
    print(f'{k:28s} {1e3 * v / (N - 1):8.3f} ms')
print(f'{"sum":28s} {1e3 * sum(acc.values()) / (N - 1):8.3f} ms   ({name}, {space.num_plans} plans, {n} records)')
