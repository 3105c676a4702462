"""Test helper: the pinned CPU oracle over WHOLE plan spaces, block-parallel on the host cores.

Each work item is a run of device-group rows of one (node sequence, stage count) block of the plan space.  The
worker rebuilds the block's rows with the oracle's own enumerator (oracle.metis_oracle.device_group_rows), builds the
inter-stage plans of the item exactly as InterStagePlanGenerator emits them (search_space/plan.py:153-175: row-major,
batches descending; `num_stage` = the block's label, which is 1 for the mislabelled Q1 blocks) and evaluates them with
oracle.het_evaluate_plan.  The block list (first ordinal, label, stage count, row count) comes from the product's
host-side enumeration, which the CPU suite checks against the oracle's generator; everything numeric is the oracle's.
Only tests import this module.
"""
import itertools
import os
import time

import numpy as np

_S = {}


def _init(root, workload_name, file_order, node_sequences):
    from metis_b200.workloads import WORKLOADS
    from oracle import metis_oracle as orc
    w = WORKLOADS[workload_name]
    cluster = orc.OracleCluster(os.path.join(root, 'hostfile'), os.path.join(root, 'clusterfile.json'))
    profile, _ = orc.load_profile_dir(os.path.join(root, 'profile'), file_order)
    model = orc.OracleModel(w.num_layers, w.hidden_size, w.sequence_length, w.vocab_size,
                            profile['model']['parameters'])
    _S.update(orc=orc, w=w, cluster=cluster, profile=profile, model=model, norm=orc.norm_layer_duration(profile),
              seqs=[tuple(s) for s in node_sequences], rows={})


def _rows(stages):
    if stages not in _S['rows']:
        orc, w = _S['orc'], _S['w']
        _S['rows'][stages] = orc.device_group_rows(stages, _S['cluster'].total_devices, w.variance, w.max_permute_len)
    return _S['rows'][stages]


def _work(item):
    """item = (first_ordinal of the block, ns_idx, label, stages, row_lo, row_hi) -> packed candidates."""
    first, ns_idx, label, stages, row_lo, row_hi = item
    orc, w = _S['orc'], _S['w']
    rows = _rows(stages)
    batches = [b for b in range(w.gbs, 0, -1) if w.gbs % b == 0]
    counters = {'A': 0, 'B': 0, 'C': 0, 'runs': 0, 'keyerr': 0}
    out = []
    for row in range(row_lo, row_hi):
        for d, b in enumerate(batches):
            ordinal = first + row * len(batches) + d
            plan = {'ns_idx': ns_idx, 'node_sequence': _S['seqs'][ns_idx], 'dg_idx': row,
                    'device_groups': list(rows[row]), 'num_stage': label, 'batches': b, 'gbs': w.gbs}
            counters['A'] += 1
            orc.het_evaluate_plan(_S['profile'], _S['cluster'], _S['model'], _S['norm'], plan, ordinal, w.num_layers,
                                  w.max_tp, w.max_bs, counters, out)
    n = len(out)
    packed = {
        'ordinal': np.array([c[0] for c in out], dtype=np.int64), 'step': np.array([c[1] for c in out], dtype=np.int64),
        'nrep': np.array([c[7] for c in out], dtype=np.int64), 'cost': np.array([c[8] for c in out], dtype=np.float64),
        'batches': np.array([c[5] for c in out], dtype=np.int64),
        'dp': np.zeros((n, stages), dtype=np.int64), 'tp': np.zeros((n, stages), dtype=np.int64),
        'part': np.zeros((n, stages + 1), dtype=np.int64), 'groups': np.zeros((n, stages), dtype=np.int64),
    }
    for i, c in enumerate(out):
        packed['groups'][i] = c[3]
        packed['dp'][i] = [d for d, _ in c[4]]
        packed['tp'][i] = [t for _, t in c[4]]
        packed['part'][i] = c[6]
    return item, packed, counters


def usable_cores():
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    try:
        txt = open('/sys/fs/cgroup/cpu.max').read().split()
        if txt[0] != 'max':
            n = max(1, min(n, int(float(txt[0]) / float(txt[1]))))
    except Exception:
        pass
    return n


def work_items(space, rows_per_item=64):
    items = []
    for blk in space.blocks:
        first, rows = int(blk['first_ordinal']), int(blk['num_rows'])
        for lo in range(0, rows, rows_per_item):
            items.append((first, int(blk['ns_idx']), int(blk['label_stage']), int(blk['num_stage']), lo,
                          min(rows, lo + rows_per_item)))
    return items


def run(root, workload_name, file_order, node_sequences, items, budget_s, procs=None):
    """Yields (item, packed candidates, counters) as the pool finishes them; stops handing out work after budget_s."""
    import multiprocessing as mp
    procs = procs or usable_cores()
    # long items first would need their cost; a fixed shuffle spreads the expensive stage counts over the run
    order = list(range(len(items)))
    np.random.default_rng(7).shuffle(order)
    t0 = time.time()
    with mp.get_context('spawn').Pool(procs, initializer=_init,
                                      initargs=(root, workload_name, list(file_order), [list(s) for s in node_sequences])) as pool:
        for res in pool.imap_unordered(_work, (items[i] for i in order), chunksize=1):
            yield res
            if time.time() - t0 > budget_s:
                pool.terminate()
                return
