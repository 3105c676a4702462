"""Test-only helpers: build and drive tests/hostsim (the g++ build of the device evaluator).

Never imported by the metis_b200 package.  Lets the CPU suite validate the flattening, the
plan-space enumeration and the device-side algorithms against the oracle / golden files.
"""
import ctypes as C
import itertools
import os
import subprocess

import numpy as np

from metis_b200 import flatten, native
from metis_b200.data_loader import ProfileDataLoader
from metis_b200.gpu_cluster import GPUCluster
from metis_b200.utils import ModelConfig

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'hostsim', 'hostsim.cpp')
OUT = os.path.join(HERE, 'hostsim', '_build', 'libhostsim.so')
_lib = None


def hostsim():
    global _lib
    if _lib is None:
        deps = [SRC, os.path.join(HERE, '..', 'metis_b200', 'csrc', 'metis_eval.cuh'),
                os.path.join(HERE, '..', 'metis_b200', 'csrc', 'metis_coop.cuh'),
                os.path.join(HERE, '..', 'metis_b200', 'csrc', 'metis_trace.cuh'),
                os.path.join(HERE, '..', 'metis_b200', 'csrc', 'metis_rows.cuh'),
                os.path.join(HERE, '..', 'include', 'metis_b200.h')]
        if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
            os.makedirs(os.path.dirname(OUT), exist_ok=True)
            tmp = f'{OUT}.{os.getpid()}.tmp'
            subprocess.check_call(['g++', '-O2', '-std=c++17', '-ffp-contract=off', '-fPIC', '-shared',
                                   '-o', tmp, SRC])
            os.replace(tmp, OUT)                             # atomic: concurrent test processes may race
        _lib = C.CDLL(OUT)
    return _lib


def load_inputs(root, profile_sub, file_order, num_layers, hidden, seq, vocab):
    cluster = GPUCluster(os.path.join(root, 'hostfile'), os.path.join(root, 'clusterfile.json'))
    profile, types = ProfileDataLoader(os.path.join(root, profile_sub), file_order).load_profile_data_all()
    cfg = ModelConfig(model_name='t', num_layers=num_layers, sequence_length=seq, vocab_size=vocab,
                      hidden_size=hidden, attention_head_size=32)
    return cluster, profile, types, cfg


def host_het_search(problem: flatten.FlatProblem, space: flatten.FlatPlanSpace, rank=0, world=1, tile=32,
                    capacity=None, want_detail=True, mode=1):
    """Runs the host-compiled evaluator; returns (records, detail, summary)."""
    lib = hostsim()
    keep = dict(problem.arrays)
    keep.update(blocks=space.blocks, batches=space.batches, rows=space.rows)
    p = problem.as_struct(lambda n: keep[n].ctypes.data)
    s = space.as_struct(lambda n: keep[n].ctypes.data)
    shard = native.MetisShard(rank, world, tile, 0)
    if capacity is None:
        capacity = max(1024, space.num_plans * 4)
    records = np.zeros(capacity, dtype=native.RECORD_DTYPE)
    detail = np.zeros((capacity, native.DETAIL_STRIDE), dtype=np.uint8) if want_detail else None
    summary = native.MetisSearchSummary()
    rc = lib.hostsim_het_search(C.byref(p), C.byref(s), C.byref(shard), C.c_void_p(records.ctypes.data),
                                C.c_int64(capacity), C.c_void_p(detail.ctypes.data if want_detail else 0),
                                C.c_int32(native.DETAIL_STRIDE), C.byref(summary), C.c_int32(mode))
    assert rc == 0
    n = min(int(summary.num_records), capacity)
    return records[:n], (detail[:n] if want_detail else None), summary


def unpack_candidates(records, detail, space: flatten.FlatPlanSpace):
    """-> list of (ordinal, step, ns_idx, groups, strategies, batches, partition, nrep, cost)."""
    order = np.lexsort((records['step'], records['ordinal']))
    out = []
    for i in order:
        r = records[i]
        ns, _label, _row, batches, row = space.locate(int(r['ordinal']))
        S = int(r['num_stage'])
        d = detail[i]
        groups = [1 << int(c) for c in row]
        strategies = [(1 << int(d[s]), 1 << int(d[S + s])) for s in range(S)]
        part = [int(x) for x in d[2 * S:3 * S + 1]]
        out.append((int(r['ordinal']), int(r['step']), ns, groups, strategies, batches, part,
                    int(r['num_repartition']), float(r['cost'])))
    return out


def host_layer_balance(capa_rows, lc, num_layers):
    lib = hostsim()
    n = len(capa_rows)
    stride = max(len(c) for c in capa_rows)
    capa = np.zeros((n, stride))
    ns = np.zeros(n, dtype=np.int32)
    for i, c in enumerate(capa_rows):
        capa[i, :len(c)] = c
        ns[i] = len(c)
    lc = np.asarray(lc, dtype=np.float64)
    out = np.zeros((n, stride + 1), dtype=np.uint16)
    rc = lib.hostsim_layer_balance(C.c_void_p(capa.ctypes.data), C.c_void_p(ns.ctypes.data), C.c_int64(n),
                                   C.c_int32(stride), C.c_void_p(lc.ctypes.data), C.c_int32(len(lc)),
                                   C.c_int32(num_layers), C.c_void_p(out.ctypes.data))
    assert rc == 0
    return [out[i, :ns[i] + 1].tolist() for i in range(n)]


def host_homo_cost(problem: flatten.FlatProblem, type_id, plans):
    lib = hostsim()
    keep = dict(problem.arrays)
    p = problem.as_struct(lambda n: keep[n].ctypes.data)
    plans = np.ascontiguousarray(plans, dtype=np.int32)
    cost = np.zeros(len(plans))
    status = np.zeros(len(plans), dtype=np.int32)
    rc = lib.hostsim_homo_cost(C.byref(p), C.c_int32(type_id), C.c_void_p(plans.ctypes.data), C.c_int64(len(plans)),
                               C.c_void_p(cost.ctypes.data), C.c_void_p(status.ctypes.data))
    assert rc == 0
    return cost, status
