"""CPU tests: the C-ABI library loads and exports every symbol the header declares, host mirror
classes behave like the reference's, and the multi-rank reduction logic works (gloo, world 2)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import C1_DIR, REPO, load_golden


def test_library_exports_every_declared_symbol():
    from metis_b200 import build, native
    build.build_library()                       # nvcc cross-compiles without a GPU
    lib = native.load_library()
    header = open(os.path.join(REPO, 'include', 'metis_b200.h')).read()
    declared = set(re.findall(r'\b(metis_[a-z_0-9]+)\s*\(', header))
    assert declared == set(native.SYMBOLS)
    for sym in declared:
        assert getattr(lib, sym) is not None
    assert lib.metis_abi_version() == 2


def test_struct_layouts_match_header():
    import ctypes as C
    from metis_b200 import native
    assert C.sizeof(native.MetisRecord) == 16
    assert C.sizeof(native.MetisPlanBlock) == 32
    assert C.sizeof(native.MetisShard) == 16
    assert C.sizeof(native.MetisSearchSummary) == 8 * 5 + 8 + 16 + 48
    assert np.dtype(native.RECORD_DTYPE).itemsize == 16 and np.dtype(native.BLOCK_DTYPE).itemsize == 32
    # compile a probe with the real header and compare sizes / offsets
    probe = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "metis_b200.h"
    int main(){printf("%zu %zu %zu %zu %zu %zu\n", sizeof(MetisProblem), offsetof(MetisProblem, key_index),
      offsetof(MetisProblem, optimizer_time), sizeof(MetisPlanSpace), sizeof(MetisSearchSummary),
      offsetof(MetisSearchSummary, best));return 0;}'''
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, 'p.c')
        open(src, 'w').write(probe)
        exe = os.path.join(tmp, 'p')
        subprocess.check_call(['gcc', '-I', os.path.join(REPO, 'include'), '-o', exe, src])
        got = [int(x) for x in subprocess.check_output([exe]).split()]
    P, S = native.MetisProblem, native.MetisSearchSummary
    assert got == [C.sizeof(P), P.key_index.offset, P.optimizer_time.offset, C.sizeof(native.MetisPlanSpace),
                   C.sizeof(S), S.best.offset]


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    from metis_b200 import native, search
    with pytest.raises(native.MetisNativeError):
        search.layer_balance([[0.5, 0.5]], [0.5, 0.5], 2)


def test_product_never_imports_oracle():
    pkg = os.path.join(REPO, 'metis_b200')
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.cpp')):
                text = open(os.path.join(root, f)).read()
                assert 'import oracle' not in text and 'from oracle' not in text and 'hostsim' not in text.replace(
                    'tests/hostsim', ''), f
    for f in ('cost_het_cluster.py', 'cost_homo_cluster.py', 'arguments.py'):
        assert 'oracle' not in open(os.path.join(REPO, f)).read()


def test_arguments_same_flags_as_reference():
    from metis_b200.arguments import parse_args
    a = parse_args(['--model_name', 'GPT', '--model_size', '1.5B', '--num_layers', '10', '--gbs', '128',
                    '--hidden_size', '4096', '--sequence_length', '1024', '--vocab_size', '51200',
                    '--attention_head_size', '32', '--hostfile_path', 'h', '--clusterfile_path', 'c',
                    '--log_path', 'l', '--home_dir', 'd', '--profile_data_path', 'p',
                    '--max_profiled_tp_degree', '4', '--max_profiled_batch_size', '4',
                    '--min_group_scale_variance', '1', '--max_permute_len', '4'])
    assert (a.gbs, a.num_layers, a.min_group_scale_variance, a.max_permute_len, a.model_size) == (128, 10, 1, 4, '1.5B')
    assert len(vars(a)) == 17
    with pytest.raises(SystemExit):
        parse_args(['--min_group_scale_variance', '0.5'])        # type=int in the reference (quirk Q12)


def test_loader_cluster_and_generators_mirror_reference():
    from metis_b200 import api
    from metis_b200.data_loader import ProfileDataLoader
    from metis_b200.gpu_cluster import GPUCluster
    from metis_b200.utils import DeviceType
    from oracle import metis_oracle as orc
    meta, _ = load_golden('c1_het')
    cl = GPUCluster(os.path.join(C1_DIR, 'hostfile'), os.path.join(C1_DIR, 'clusterfile.json'))
    assert cl.get_total_num_devices() == 64 and cl.get_num_nodes() == 8 and cl.get_num_devices_per_node() == 8
    assert cl.get_device_types() == [DeviceType.A100] * 8
    assert cl.get_inter_bandwidth(0) == cl.get_intra_bandwidth(0) == 5312500000.0     # quirk Q2
    assert cl.get_device_memory_for_device_type('A100') == 80 * 1024
    prof, types = ProfileDataLoader(os.path.join(C1_DIR, 'profile_data_samples'), meta['file_order']).load_profile_data_all()
    oprof, otypes = orc.load_profile_dir(os.path.join(C1_DIR, 'profile_data_samples'), meta['file_order'])
    assert prof == oprof and types == otypes == ['A100']
    assert prof['model']['optimizer_time'] == 2 * 20.301532745361328          # first listed file: tp2_bs2 (Q3)
    with pytest.raises(ValueError):
        DeviceType.from_string('tpu')
    got = [(p.dp, p.pp, p.tp, p.mbs, p.gbs) for p in api.UniformPlanGenerator(64, 4, 128)]
    assert got == list(orc.uniform_plans(64, 4, 128)) and len(got) == 345
    gen = api.InterStagePlanGenerator({DeviceType.A100}, 64, 128, 10, 1, 4)
    mine = [(p.ns_idx, p.dg_idx, p.device_groups, p.num_stage, p.batches) for p in gen]
    want = [(p['ns_idx'], p['dg_idx'], p['device_groups'], p['num_stage'], p['batches'])
            for p in orc.inter_stage_plans([('A100',)], 64, 128, 10, 1, 4)]
    assert mine == want and len(mine) == 32


def test_plan_space_matches_oracle_order_including_q1():
    from metis_b200 import flatten
    from oracle import metis_oracle as orc
    for ndev, ntypes, gbs, L, var, mpl in [(16, 2, 32, 10, 1, 4), (8, 3, 12, 6, 0, 3), (32, 2, 64, 24, 1, 6)]:
        seqs = [('A',) * 1, ('B',)][:1] if ntypes == 1 else None
        import itertools
        names = ['A100', 'V100', 'T4'][:ntypes]
        seqs = list(itertools.permutations(names))
        space = flatten.build_plan_space(len(seqs), ndev, gbs, L, var, mpl)
        want = list(orc.inter_stage_plans(seqs, ndev, gbs, L, var, mpl))
        assert space.num_plans == len(want)
        for o in list(range(0, len(want), 7)) + [len(want) - 1]:
            ns, label, row, batches, codes = space.locate(o)
            w = want[o]
            assert (ns, label, row, batches, [1 << int(c) for c in codes]) == \
                (w['ns_idx'], w['num_stage'], w['dg_idx'], w['batches'], w['device_groups'])


WORKER = r'''
import os, sys
sys.path.insert(0, os.environ['REPO']); sys.path.insert(0, os.path.join(os.environ['REPO'], 'tests'))
import torch, torch.distributed as dist
dist.init_process_group('gloo', init_method='tcp://127.0.0.1:' + os.environ['PORT'],
                        rank=int(os.environ['RANK']), world_size=2)
import numpy as np
import hostsim_util as hs
from conftest import load_golden
from metis_b200 import flatten, search
from metis_b200.workloads import WORKLOADS, materialize
import tempfile
meta, arr = load_golden('c2_v100')
w = WORKLOADS['c2_v100']
root = tempfile.mkdtemp(); materialize(w, root)
cluster, profile, _, cfg = hs.load_inputs(root, 'profile', meta['file_order'], w.num_layers, w.hidden_size, w.sequence_length, w.vocab_size)
seqs = [tuple(s) for s in meta['node_sequences']]
problem = flatten.build_problem(profile, cluster, cfg, w.gbs, w.max_tp, w.max_bs, seqs)
space = flatten.build_plan_space(len(seqs), 16, w.gbs, w.num_layers, w.variance, w.max_permute_len)
rank = dist.get_rank()
rec, det, sm = hs.host_het_search(problem, space, rank=rank, world=2, tile=64)   # shard compute: test shim
local_best = (sm.best.cost, sm.best.ordinal, sm.best.step, sm.best.num_repartition, sm.best.num_stage) if sm.num_records else None
best = search.global_best(local_best, 'cpu')                                     # product reduction logic
counters = search.global_counters(dict(num_records=int(sm.num_records), num_partition_calls=int(sm.num_partition_calls),
    num_balancer_runs=int(sm.num_balancer_runs), num_keyerror=int(sm.num_keyerror), fatal_ordinal=int(sm.fatal_ordinal)), 'cpu')
i = int(np.lexsort((arr['step'], arr['ordinal'], arr['cost']))[0])
assert best[:3] == (float(arr['cost'][i]), int(arr['ordinal'][i]), int(arr['step'][i])), best
c = meta['counters']
assert (counters['num_records'], counters['num_partition_calls'], counters['num_balancer_runs']) == (c['C'], c['B'], c['runs']), counters
assert 0 < sm.num_records < c['C']
# the API path's single collective gives the same counters / winner, the records of every rank, and spreads an error
both, best2 = search.global_exchange(dict(num_records=int(sm.num_records), num_partition_calls=int(sm.num_partition_calls),
    num_balancer_runs=int(sm.num_balancer_runs), num_keyerror=int(sm.num_keyerror), fatal_ordinal=int(sm.fatal_ordinal)),
    local_best, 'cpu', local_error=int(rank == 1))
assert best2 == best and both['num_records'] == c['C'] and both['any_rank_failed'] == 1
assert sum(both['records_per_rank']) == c['C'] and both['records_per_rank'][rank] == int(sm.num_records)
assert both['global_fatal_ordinal'] == 2 ** 62
fatal = search.global_counters(dict(fatal_ordinal=1000 + rank, fatal_code=3 + rank, fatal_aux=70 + rank), 'cpu')
assert (fatal['global_fatal_ordinal'], fatal['global_fatal_code'], fatal['global_fatal_aux']) == (1000, 3, 70)
dist.barrier(); dist.destroy_process_group()
print('rank', rank, 'ok')
'''


def test_two_rank_reduction_gloo(tmp_path):
    """world_size 2 over gloo: each rank evaluates its interleaved shard, then the product's
    collective logic (search.global_best / global_counters) yields the golden winner and counters."""
    load_golden('c2_v100')
    import hostsim_util
    hostsim_util.hostsim()                      # build the test shim once, before the ranks start
    script = tmp_path / 'worker.py'
    script.write_text(WORKER)
    import socket
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    procs = []
    for rank in range(2):
        env = dict(os.environ, REPO=REPO, RANK=str(rank), PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), '\n'.join(outs)


def test_python_constants_match_header():
    """The numeric constants the ctypes layer uses are the ones include/metis_b200.h defines."""
    import re
    from metis_b200 import native
    text = open(os.path.join(REPO, 'include', 'metis_b200.h')).read()
    defs = {m.group(1): int(m.group(2)) for m in re.finditer(r'#define\s+(METIS_[A-Z_0-9]+)\s+(-?\d+)\b', text)}
    assert (defs['METIS_SORT_POSITION'], defs['METIS_SORT_RANKED'], defs['METIS_SORT_BY_COST_STABLE']) == \
        (native.SORT_POSITION, native.SORT_RANKED, native.SORT_BY_COST_STABLE)
    for code, name in native.FATAL_NAMES.items():
        assert defs['METIS_FATAL_' + name] == code
    assert native.DETAIL_STRIDE == 3 * defs['METIS_MAX_STAGES'] + 1


def test_enumeration_into_a_caller_buffer():
    """build_plan_space(rows_out=...) (the pinned staging buffer in production) yields the same space as the
    allocating call, in place, and falls back to its own buffer when the caller's is too small."""
    from metis_b200 import flatten
    base = flatten.build_plan_space(2, 32, 64, 24, 1, 4)
    buf = np.full(base.rows.size + 4096, 0xAB, dtype=np.uint8)
    inplace = flatten.build_plan_space(2, 32, 64, 24, 1, 4, rows_out=buf)
    assert np.shares_memory(inplace.rows, buf)
    assert inplace.num_plans == base.num_plans and (inplace.blocks == base.blocks).all()
    assert (inplace.rows[:base.rows.size] == base.rows).all()
    small = np.zeros(16, dtype=np.uint8)
    other = flatten.build_plan_space(2, 32, 64, 24, 1, 4, rows_out=small)
    assert not np.shares_memory(other.rows, small) and (other.rows == base.rows).all()


def _enumerate_in_child(queue):
    from metis_b200 import flatten
    queue.put(int(flatten.build_plan_space(1, 64, 512, 96, 1, 4).num_plans))


def test_enumerator_worker_pool_survives_fork():
    """The C++ enumerator keeps persistent worker threads; a forked child (where they do not exist) must build
    its own pool instead of waiting for the parent's."""
    import multiprocessing as mp
    from metis_b200 import flatten
    want = int(flatten.build_plan_space(1, 64, 512, 96, 1, 4).num_plans)      # creates the pool in this process
    ctx = mp.get_context('fork')
    q = ctx.Queue()
    p = ctx.Process(target=_enumerate_in_child, args=(q,))
    p.start()
    p.join(60)
    assert not p.is_alive(), 'enumeration in the forked child did not finish'
    assert p.exitcode == 0 and q.get(timeout=5) == want == 82520
    assert int(flatten.build_plan_space(1, 64, 512, 96, 1, 4).num_plans) == want   # the parent's pool still works


def _write_profile(path, layers=4, **overrides):
    import json
    raw = {'model': {'parameters': {'parameters_per_layer_bytes': [10] * layers}},
           'execution_time': {'layer_compute_total_ms': [1.0] * layers, 'forward_backward_time_ms': 5.0,
                              'optimizer_time_ms': 2.0, 'batch_generator_time_ms': 0.5},
           'execution_memory': {'layer_memory_total_mb': [100.0] * layers}}
    for dotted, value in overrides.items():
        node = raw
        keys = dotted.split('__')
        for k in keys[:-1]:
            node = node[k]
        if value is None:
            del node[keys[-1]]
        else:
            node[keys[-1]] = value
    with open(path, 'w') as fh:
        json.dump(raw, fh)


def test_opt_in_profile_validation_and_sorted_listing(tmp_path):
    """schema validation and a deterministic listing are opt-in; the default loader behaves like
    the reference (no checks, os.listdir order)."""
    from metis_b200.data_loader import ProfileDataLoader, ProfileSchemaError
    d = tmp_path / 'p'
    d.mkdir()
    _write_profile(d / 'DeviceType.H100_tp1_bs1.json')
    _write_profile(d / 'DeviceType.A100_tp2_bs1.json')
    data, types = ProfileDataLoader(str(d), sort_files=True, validate=True).load_profile_data_all()
    assert types == ['A100', 'H100'] and set(data) == {'model', 'DeviceType.A100', 'DeviceType.H100'}
    assert data['DeviceType.A100']['tp2_bs1']['time']['fb_sync'] == 1.0
    # the same files through the default path give the same dict (order of types aside)
    plain, _ = ProfileDataLoader(str(d)).load_profile_data_all()
    assert plain['DeviceType.H100'] == data['DeviceType.H100']
    for bad, text in [({'execution_time__layer_compute_total_ms': None}, 'missing key execution_time/layer_compute_total_ms'),
                      ({'execution_memory__layer_memory_total_mb': [1.0, 'x', 3.0, 4.0]}, 'non-numeric'),
                      ({'execution_memory__layer_memory_total_mb': [1.0, 2.0]}, 'differ in length'),
                      ({'execution_time__optimizer_time_ms': '3'}, 'must be a number')]:
        _write_profile(d / 'DeviceType.H100_tp1_bs1.json', **bad)
        with pytest.raises(ProfileSchemaError, match=text):
            ProfileDataLoader(str(d), sort_files=True, validate=True).load_profile_data_all()
    _write_profile(d / 'DeviceType.H100_tp1_bs1.json', layers=6)
    with pytest.raises(ProfileSchemaError, match='6 layers, other profiles have 4'):
        ProfileDataLoader(str(d), sort_files=True, validate=True).load_profile_data_all()
    _write_profile(d / 'DeviceType.H100_tp1_bs1.json')
    _write_profile(d / 'notes.json')
    with pytest.raises(ProfileSchemaError, match='file name must look like'):
        ProfileDataLoader(str(d), sort_files=True, validate=True).load_profile_data_all()


def test_opt_in_strict_cluster_files(tmp_path):
    """Multi-digit `slots=` counts and named errors are opt-in; the default keeps quirk Q10 (`slots=16` reads 1)."""
    import json
    from metis_b200.gpu_cluster import GPUCluster
    from metis_b200.utils import parse_hostfile
    host = tmp_path / 'hostfile'
    host.write_text('n0 slots=16\nn1 slots=16\n\n')
    nodes = {f'n{i}': {'instance_type': 'B200', 'inter_bandwidth': 1e9, 'intra_bandwidth': 9e11, 'memory': 180}
             for i in range(2)}
    cl = tmp_path / 'cluster.json'
    cl.write_text(json.dumps(nodes))
    assert [e['num_device'] for e in parse_hostfile(str(host), strict=True).values()] == [16, 16]
    loose = tmp_path / 'hostfile_q10'
    loose.write_text('n0 slots=16\nn1 slots=16\n')
    assert [e['num_device'] for e in parse_hostfile(str(loose)).values()] == [1, 1]      # the reference's reading
    cluster = GPUCluster(str(host), str(cl), strict=True)
    assert cluster.get_total_num_devices() == 32 and cluster.get_num_devices_per_node() == 16
    assert cluster.get_device_memory(0) == 180 * 1024
    host.write_text('n0 slots=16\nn7 slots=16\n')
    with pytest.raises(ValueError, match="host 'n7'"):
        GPUCluster(str(host), str(cl), strict=True)
    host.write_text('n0 slots=16\nn1 gpus=16\n')
    with pytest.raises(ValueError, match=':2: expected'):
        GPUCluster(str(host), str(cl), strict=True)
    host.write_text('n0 slots=16\nn1 slots=16\n')
    nodes['n1']['instance_type'] = 'MI300'
    cl.write_text(json.dumps(nodes))
    with pytest.raises(ValueError, match="unknown instance_type 'MI300'"):
        GPUCluster(str(host), str(cl), strict=True)


@pytest.mark.parametrize('ndev,var,mpl', [(8, 0.5, 4), (16, 1, 6), (32, 0.5, 6), (32, 0, 4), (64, 1, 4), (64, 0.5, 6),
                                          (16, 0.5, 2), (128, 0, 4)])
def test_row_generator_writes_the_host_enumerators_rows(ndev, var, mpl):
    """compositions listed by the host + the per-composition prefix-shift walk of metis_rows.cuh
    (the code of het_rows_kernel, built for the host) = the row blob of the host enumerator, byte for byte."""
    import ctypes as C
    import hostsim_util as hs
    from metis_b200 import flatten
    L = 24
    host = flatten.build_plan_space(2, ndev, 64, L, var, mpl)
    dev = flatten.build_plan_space(2, ndev, 64, L, var, mpl, device_rows=True)
    assert dev.comp_recs is not None and dev.rows.size == 0
    assert dev.num_plans == host.num_plans
    assert dev.blocks.tobytes() == host.blocks.tobytes()
    rows = np.full(dev.rows_total_bytes + 16, 0xEE, dtype=np.uint8)
    hs.hostsim().hostsim_generate_rows(C.c_void_p(dev.comp_recs.ctypes.data), C.c_int64(len(dev.comp_recs)),
                                       C.c_void_p(dev.comp_pool.ctypes.data), C.c_void_p(rows.ctypes.data))
    used = max(int(b['rows_offset']) + int(b['num_rows']) * int(b['num_stage']) for b in host.blocks)
    assert dev.rows_total_bytes >= used
    assert rows[:dev.rows_total_bytes].tobytes() == host.rows[:dev.rows_total_bytes].tobytes()
    assert (rows[dev.rows_total_bytes:] == 0xEE).all()
    # the lazily filled host tables of a device-rows space describe the same rows
    for b in host.blocks:
        st = int(b['num_stage'])
        assert dev.tables[st][0] == host.tables[st][0] and np.array_equal(dev.tables[st][1], host.tables[st][1])
