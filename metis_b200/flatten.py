"""Flatten the reference's nested inputs into the dense arrays of ``MetisProblem`` and
enumerate the candidate space into ``MetisPlanSpace`` (include/metis_b200.h).

Host-side logic only (no arithmetic of the search itself happens here, except the
load-time constants the reference also computes once on the host: ``sum(layer-computes)``
per profile key, ``norm_layer_duration`` (model/load_balancer.py:22-27) and the per-type
bandwidth/memory lookups of gpu_cluster.py).
"""
from __future__ import annotations

import ctypes as C
import math
import sys
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import native


def py312_sum(values) -> float:
    """CPython >= 3.12 ``sum`` (Neumaier-compensated for floats, Python/bltinmodule.c) so that
    load-time constants do not depend on the interpreter version running the host code."""
    if sys.version_info >= (3, 12):
        return sum(values)                                  # the interpreter's own sum IS this algorithm
    it = iter(values)
    acc = 0
    for x in it:
        if isinstance(x, int) and not isinstance(x, bool):
            acc += x
            continue
        acc = acc + x
        break
    else:
        return acc
    f, c = float(acc), 0.0
    for x in it:
        if isinstance(x, float):
            t = f + x
            c += ((f - t) + x) if abs(f) >= abs(x) else ((x - t) + f)
            f = t
        else:
            f += float(x)
    if c and math.isfinite(c):
        f += c
    return f


def _numeric_list(values, what: str) -> np.ndarray:
    kinds = {type(v) for v in values}
    if not kinds <= {int, float}:
        raise TypeError(f'{what}: only int/float entries are supported')
    if len(kinds) == 2:
        # CPython's sum() leaves ints un-compensated inside a float list; not representable here
        raise NotImplementedError(f'{what}: mixed int/float arrays are not supported')
    return np.asarray(values, dtype=np.float64)


def _type_name(t) -> str:
    return t.name if hasattr(t, 'name') else str(t)


@dataclass
class FlatProblem:
    """Numpy twin of MetisProblem; ``as_struct`` binds pointers (host or device)."""
    scalars: Dict[str, object]
    arrays: Dict[str, np.ndarray]
    type_names: List[str]
    key_names: List[Tuple[str, int, int]]            # (type, tp, bs) per key id
    node_sequences: List[Tuple[str, ...]]

    def as_struct(self, ptr_of: Callable[[str], int]) -> native.MetisProblem:
        p = native.MetisProblem()
        for k, v in self.scalars.items():
            setattr(p, k, v)
        for name in ('key_index', 'layer_compute', 'layer_memory', 'exec_full', 'fb_sync', 'norm_lc',
                     'type_memory', 'type_bw_first', 'type_bw_min', 'ns_run_type', 'ns_run_end', 'ns_q10_end'):
            setattr(p, name, ptr_of(name))
        return p


def norm_layer_duration(profile_data: Dict) -> List[float]:
    """model/load_balancer.py:22-27: weights from the first-listed type's tp1_bs1 (quirk Q3)."""
    first = next(iter(profile_data))
    durations = profile_data[first]['tp1_bs1']['time']['layer-computes']
    total = py312_sum(durations)
    return [d / total for d in durations]


def build_problem(profile_data: Dict, gpu_cluster, model_config, gbs: int, max_tp: int, max_bs: int,
                  node_sequences: Sequence[Sequence], norm_lc: Optional[Sequence[float]] = None,
                  corrected: Sequence[str] = ()) -> FlatProblem:
    """``corrected`` (opt-in): 'Q2' fills the between-node bandwidth table from the
    clusterfile's ``inter_bandwidth`` instead of reproducing gpu_cluster.py:56-58, which returns the intra value;
    'Q5' / 'Q6' set the METIS_FIX_* bits evaluated on the device (vote without dropped layers, memory demand from
    the stage's own device type)."""
    nodes = [gpu_cluster.nodes[i] for i in gpu_cluster.nodes.keys()]
    per_node = nodes[0].num_devices                         # gpu_cluster.py:25-26: node 0's count stands for all (Q10)
    type_names: List[str] = []
    for n in nodes:
        if _type_name(n.device_type) not in type_names:
            type_names.append(_type_name(n.device_type))
    if len(type_names) > native.METIS_MAX_TYPES:
        raise NotImplementedError(f'more than {native.METIS_MAX_TYPES} device types')
    if len(node_sequences) > 256:                            # ns_idx travels in 8 bits of the list entries
        raise NotImplementedError('more than 256 node sequences')
    num_layers = model_config.num_layers
    if num_layers > min(native.METIS_MAX_LAYERS, 255):       # layer_partition entries travel as one byte
        raise NotImplementedError('--num_layers > 255')

    num_tp = max(1, int(math.floor(math.log2(max_tp))) + 1) if max_tp >= 1 else 1
    profiled_bs = [1]
    for name in type_names:
        for key in profile_data.get(f'DeviceType.{name}', {}):
            profiled_bs.append(int(key.split('_bs')[1]))
    num_bs = min(max(max(profiled_bs), max_bs, 1), 4096)

    key_index = np.full((len(type_names), num_tp, num_bs), -1, dtype=np.int16)
    key_names: List[Tuple[str, int, int]] = []
    lc_rows, mem_rows, exec_full, fb_sync = [], [], [], []
    for ti, name in enumerate(type_names):
        for key, entry in profile_data.get(f'DeviceType.{name}', {}).items():
            tp = int(key[2:].split('_bs')[0])
            bs = int(key.split('_bs')[1])
            if tp < 1 or tp & (tp - 1) or tp > (1 << (num_tp - 1)) or bs < 1 or bs > num_bs:
                continue                                    # never addressed by the het search
            lc = entry['time']['layer-computes']
            key_index[ti, int(math.log2(tp)), bs - 1] = len(key_names)
            key_names.append((name, tp, bs))
            lc_rows.append(_numeric_list(lc, f'{name} {key} layer_compute_total_ms'))
            mem_rows.append(_numeric_list(entry['memory'], f'{name} {key} layer_memory_total_mb'))
            exec_full.append(float(py312_sum(lc)))
            fb = entry['time'].get('fb_sync')
            fb_sync.append(float(fb) if fb else 0.0)         # falsy -> KeyError in the reference (Q9)
    if not key_names:
        raise KeyError('no profile data for the device types of the cluster')
    if norm_lc is None:
        norm_lc = norm_layer_duration(profile_data)
    lpad = max([num_layers] + [len(r) for r in lc_rows] + [len(r) for r in mem_rows])
    lpad += lpad & 1
    layer_compute = np.zeros((len(key_names), lpad), dtype=np.float64)
    layer_memory = np.zeros((len(key_names), lpad), dtype=np.float64)
    for i, (lc, mem) in enumerate(zip(lc_rows, mem_rows)):
        layer_compute[i, :len(lc)] = lc
        layer_memory[i, :len(mem)] = mem

    type_memory, bw_first, bw_min = [], [], []
    node_ids = list(gpu_cluster.nodes.keys())
    for name in type_names:
        mem = gpu_cluster.get_device_memory_for_device_type(name)
        if mem is None:
            raise TypeError("unsupported operand type(s) for *: 'NoneType' and 'int'")   # device_group.py:99-100
        type_memory.append(float(mem))
        mine = [i for i in node_ids if _type_name(gpu_cluster.nodes[i].device_type) == name]
        bw_first.append(float(gpu_cluster.get_intra_bandwidth(mine[0])))     # cluster_bandwidth.py:49-54
        if 'Q2' in corrected:
            bw_min.append(float(min(gpu_cluster.nodes_info[gpu_cluster.host_entries[i]['ip']]['inter_bandwidth']
                                    for i in mine)))
        else:
            bw_min.append(float(min(gpu_cluster.get_inter_bandwidth(i) for i in mine)))   # :56-68 (Q2)
    uniform_bw = int(len(set(bw_first + bw_min)) == 1)

    seqs = [tuple(_type_name(t) for t in seq) for seq in node_sequences]
    run_type = np.zeros((len(seqs), len(type_names)), dtype=np.uint8)
    run_end = np.zeros((len(seqs), len(type_names)), dtype=np.int32)
    q10_end = np.zeros((len(seqs), len(type_names)), dtype=np.int32)
    nodes_of = {name: sum(1 for n in nodes if _type_name(n.device_type) == name) for name in type_names}
    for si, seq in enumerate(seqs):
        if sorted(seq) != sorted(type_names):
            raise ValueError('node sequence is not a permutation of the cluster device types')
        total = total_q10 = 0
        for k, name in enumerate(seq):
            total += gpu_cluster.get_num_nodes_by_device_type(name)      # devices of the type: model/device_group.py:22-32
            total_q10 += nodes_of[name] * per_node                       # load_balancer.py:109-119 (Q10)
            run_type[si, k] = type_names.index(name)
            run_end[si, k] = total
            q10_end[si, k] = total_q10

    params = profile_data['model']['parameters']
    scalars = dict(
        num_types=len(type_names), num_tp=num_tp, num_bs=num_bs, num_keys=len(key_names), lpad=lpad,
        num_layers=num_layers, norm_len=len(norm_lc), gbs=gbs, max_tp=max_tp, max_bs=max_bs,
        num_nodes=len(nodes), devices_per_node=per_node, total_devices=int(sum(n.num_devices for n in nodes)),
        num_node_sequences=len(seqs), uniform_bw=uniform_bw, q10_devices=per_node * len(nodes),
        corrected=(1 if 'Q5' in corrected else 0) | (2 if 'Q6' in corrected else 0), reserved1=0,
        sequence_length=int(model_config.sequence_length), hidden_size=int(model_config.hidden_size),
        vocab_size=int(model_config.vocab_size),
        optimizer_time=float(profile_data['model']['optimizer_time']),
        batch_generator=float(profile_data['model']['batch_generator']),
        input_params=float(params[0]), transformer_params=float(params[1]), output_params=float(params[-1]),
        node0_bandwidth=float(gpu_cluster.get_intra_bandwidth(0)),
        node0_memory=float(gpu_cluster.get_device_memory(0)),
    )
    arrays = dict(
        key_index=np.ascontiguousarray(key_index), layer_compute=layer_compute, layer_memory=layer_memory,
        exec_full=np.asarray(exec_full, dtype=np.float64), fb_sync=np.asarray(fb_sync, dtype=np.float64),
        norm_lc=np.asarray(norm_lc, dtype=np.float64),
        type_memory=np.asarray(type_memory, dtype=np.float64),
        type_bw_first=np.asarray(bw_first, dtype=np.float64), type_bw_min=np.asarray(bw_min, dtype=np.float64),
        ns_run_type=run_type, ns_run_end=run_end, ns_q10_end=q10_end,
    )
    return FlatProblem(scalars, arrays, type_names, key_names, seqs)


# ---------------------------------------------------------------------------------------------
# candidate space
# ---------------------------------------------------------------------------------------------
def enumerate_device_groups(num_stages: int, num_gpus: int, variance, max_permute_len: int,
                            lib=None) -> np.ndarray:
    """Rows of gen_dgroups_for_stages_with_variance (search_space/device_group.py:93-107) as
    log2 codes, shape [rows, num_stages]; enumerated by the library's C++ host enumerator."""
    lib = lib or native.load_library()
    n = lib.metis_enum_device_groups(num_stages, num_gpus, float(variance), max_permute_len, None, 0)
    if n < 0:
        raise native.MetisNativeError(f'metis_enum_device_groups failed ({n})')
    out = np.empty((n, num_stages), dtype=np.uint8)
    if n:
        got = lib.metis_enum_device_groups(num_stages, num_gpus, float(variance), max_permute_len,
                                           out.ctypes.data, n)
        if got != n:
            raise native.MetisNativeError('metis_enum_device_groups: inconsistent row count')
    return out


def enumerate_device_group_tables(first_stage: int, last_stage: int, num_gpus: int, variance, max_permute_len: int,
                                  lib=None, out: Optional[np.ndarray] = None) -> Dict[int, np.ndarray]:
    """All row tables for stage counts first_stage..last_stage in one threaded library call.  ``out`` (uint8,
    e.g. the pinned staging buffer of a DeviceProblem) receives the tables when it is large enough."""
    lib = lib or native.load_library()
    n = last_stage - first_stage + 1
    counts = np.zeros(n, dtype=np.int64)
    total = lib.metis_enum_device_group_tables(first_stage, last_stage, num_gpus, float(variance), max_permute_len,
                                               counts.ctypes.data, None, 0)
    if total < 0:
        raise native.MetisNativeError(f'metis_enum_device_group_tables failed ({total})')
    padded = ((max(int(total), 1) + 15) // 16) * 16                              # 16 B multiple for the device copy
    if out is not None and out.dtype == np.uint8 and out.ndim == 1 and out.size >= padded and out.flags.c_contiguous:
        blob = out[:padded]
        blob[int(total):] = 0
    else:
        blob = np.zeros(padded, dtype=np.uint8)
    got = lib.metis_enum_device_group_tables(first_stage, last_stage, num_gpus, float(variance), max_permute_len,
                                             counts.ctypes.data, blob.ctypes.data, int(total))
    if got != total:
        raise native.MetisNativeError('metis_enum_device_group_tables: inconsistent size')
    out: Dict[int, np.ndarray] = {}
    off = 0
    for i in range(n):
        stages = first_stage + i
        size = int(counts[i]) * stages
        out[stages] = blob[off:off + size].reshape(int(counts[i]), stages)   # views into the one blob
        off += size
    out[0] = blob                                                            # the blob itself (offset 0 = first table)
    return out


@dataclass
class FlatPlanSpace:
    """Numpy twin of MetisPlanSpace."""
    num_plans: int
    blocks: np.ndarray            # structured, native.BLOCK_DTYPE
    batches: np.ndarray           # int32, divisors of gbs descending
    rows: np.ndarray              # uint8 blob of all row tables
    tables: Dict[int, Tuple[int, np.ndarray]] = field(default_factory=dict)   # S -> (byte offset, rows)
    rows_total_bytes: int = -1    # device_rows spaces: size of the row blob the GPU writes (rows stays empty)
    comp_recs: Optional[np.ndarray] = None     # native.COMP_DTYPE, device_rows spaces
    comp_pool: Optional[np.ndarray] = None

    def as_struct(self, ptr_of: Callable[[str], int]) -> native.MetisPlanSpace:
        s = native.MetisPlanSpace()
        s.num_plans = self.num_plans
        s.rows_bytes = int(self.rows_total_bytes if self.rows_total_bytes >= 0 else self.rows.size)
        s.num_blocks = len(self.blocks)
        s.num_div = len(self.batches)
        s.max_stage = int(self.blocks['num_stage'].max()) if len(self.blocks) else 1
        s.blocks = ptr_of('blocks')
        s.batches = ptr_of('batches')
        s.rows = ptr_of('rows')
        return s

    def host_rows(self) -> np.ndarray:
        """The row blob on the host; a device_rows space enumerates it on first use (debug / test paths only)."""
        if self.rows.size == 0 and self.comp_recs is not None:
            self.tables._fill()
            return self.tables.blob
        return self.rows

    def locate(self, ordinal: int) -> Tuple[int, int, int, int, np.ndarray]:
        """ordinal -> (ns_idx, label_stage, dg_idx, batches, device_groups) like InterStagePlan."""
        firsts = self.blocks['first_ordinal']
        b = int(np.searchsorted(firsts, ordinal, side='right')) - 1
        blk = self.blocks[b]
        rel = ordinal - int(blk['first_ordinal'])
        row, div = divmod(rel, len(self.batches))
        _, table = self.tables[int(blk['num_stage'])]
        return int(blk['ns_idx']), int(blk['label_stage']), row, int(self.batches[div]), table[row]


def _walk_blocks(num_node_sequences: int, cap: int, nrows_of: Callable[[int], int],
                 corrected: Sequence[str]) -> List[Tuple[int, int, int]]:
    """The (ns_idx, label, stage count) blocks in the order of InterStagePlanGenerator.__next__
    (search_space/plan.py:153-175), including the mislabelled num_stage=1 block of every later node sequence
    (quirk Q1; with 'Q1' in ``corrected`` every node sequence starts with the real one-stage rows)."""
    def next_stage(start: int) -> int:                         # plan.py:130-142
        s = start
        while nrows_of(s) == 0 and s <= cap:
            s += 1
        return s

    if nrows_of(1) == 0:
        raise IndexError('list index out of range')            # plan.py:117-118
    out: List[Tuple[int, int, int]] = []
    ns, label, stages = 0, 1, 1
    while True:
        out.append((ns, label, stages))
        s = next_stage(label + 1)
        if s > cap:
            ns += 1
            if ns >= num_node_sequences:
                break
            if 'Q1' in corrected:
                label, stages = 1, 1
            else:
                label, stages = 1, next_stage(2)               # plan.py:144-148 (Q1)
            if nrows_of(stages) == 0:
                raise IndexError('list index out of range')    # plan.py:173
        else:
            label, stages = s, s
    return out


def _blocks_array(plan_blocks, nrows_of, offset_of, ndiv: int) -> Tuple[np.ndarray, int]:
    blocks = np.zeros(len(plan_blocks), dtype=native.BLOCK_DTYPE)
    ordinal = 0
    for i, (ns_idx, label, stages) in enumerate(plan_blocks):
        blocks[i]['first_ordinal'] = ordinal
        blocks[i]['rows_offset'] = offset_of(stages)
        blocks[i]['num_rows'] = nrows_of(stages)
        blocks[i]['ns_idx'] = ns_idx
        blocks[i]['label_stage'] = label
        blocks[i]['num_stage'] = stages
        ordinal += nrows_of(stages) * ndiv
    return blocks, ordinal


def build_plan_space(num_node_sequences: int, num_devices: int, gbs: int, num_layers: int, variance,
                     max_permute_len: int, lib=None, rows_out: Optional[np.ndarray] = None,
                     corrected: Sequence[str] = (), device_rows: bool = False) -> FlatPlanSpace:
    """The candidate space of one search.  ``device_rows``: the host only lists the compositions
    (``comp_recs`` / ``comp_pool``) and the GPU writes the rows (metis_generate_rows); ``rows`` stays empty and
    ``tables`` is filled by the host enumerator only if somebody asks for it."""
    cap = min(num_devices, num_layers)
    lib = lib or native.load_library()
    batches = [b for b in range(gbs, 0, -1) if gbs % b == 0]   # plan.py:120-124
    if device_rows:
        counts, recs, pool, most = enumerate_compositions(1, cap + 1, num_devices, variance, max_permute_len, lib)
        if most <= native.METIS_MAX_PERMUTE_GROUPS:
            offsets, off = {}, 0
            for stages in range(1, cap + 2):
                offsets[stages] = off
                off += int(counts[stages - 1]) * stages
            nrows_of = lambda st: int(counts[st - 1]) if 1 <= st <= cap + 1 else 0     # noqa: E731
            plan_blocks = _walk_blocks(num_node_sequences, cap, nrows_of, corrected)
            blocks, total = _blocks_array(plan_blocks, nrows_of, lambda st: offsets[st], len(batches))
            if off > 0xFFFFFFFF or total > 0xFFFFFFF0:
                raise NotImplementedError('device-group tables of 4 GiB or more / more than 2^32 plans are not supported')
            space = FlatPlanSpace(total, blocks, np.asarray(batches, dtype=np.int32), np.zeros(0, dtype=np.uint8))
            space.rows_total_bytes = off
            space.comp_recs, space.comp_pool = recs, pool
            space.tables = _LazyTables(cap, num_devices, variance, max_permute_len)
            return space
        # a composition with more merged groups than the device kernel handles: enumerate on the host
    cache: Dict[int, np.ndarray] = enumerate_device_group_tables(1, cap + 1, num_devices, variance, max_permute_len, lib,
                                                                 rows_out)
    blob = cache.pop(0)
    base_addr = blob.__array_interface__['data'][0]
    nrows_of = lambda st: len(cache[st]) if st in cache else 0                           # noqa: E731
    offset_of = lambda st: cache[st].__array_interface__['data'][0] - base_addr         # noqa: E731
    plan_blocks = _walk_blocks(num_node_sequences, cap, nrows_of, corrected)
    blocks, total = _blocks_array(plan_blocks, nrows_of, offset_of, len(batches))
    tables = {st: (offset_of(st), cache[st]) for _, _, st in plan_blocks}
    if blob.size > 0xFFFFFFFF:                               # list entries address a row with 32 bits
        raise NotImplementedError('device-group tables of 4 GiB or more are not supported')
    if total > 0xFFFFFFF0:
        raise NotImplementedError('more than 2^32 inter-stage plans')
    return FlatPlanSpace(total, blocks, np.asarray(batches, dtype=np.int32), blob, tables)


class _LazyTables(dict):
    """S -> (byte offset, rows): filled by the host enumerator on first use (device_rows spaces)."""

    def __init__(self, cap, num_devices, variance, max_permute_len):
        super().__init__()
        self._args = (cap, num_devices, variance, max_permute_len)
        self._done = False
        self.blob = None

    def _fill(self):
        if not self._done:
            cap, num_devices, variance, mpl = self._args
            cache = enumerate_device_group_tables(1, cap + 1, num_devices, variance, mpl)
            blob = self.blob = cache.pop(0)
            base = blob.__array_interface__['data'][0]
            for st, rows in cache.items():
                dict.__setitem__(self, st, (rows.__array_interface__['data'][0] - base, rows))
            self._done = True

    def __getitem__(self, key):
        self._fill()
        return dict.__getitem__(self, key)

    def __contains__(self, key):
        self._fill()
        return dict.__contains__(self, key)


def enumerate_compositions(first_stage: int, last_stage: int, num_gpus: int, variance, max_permute_len: int, lib=None):
    """metis_enum_compositions: (rows per stage count, MetisCompRec array, pool bytes, largest number of merged groups)."""
    import ctypes as C
    lib = lib or native.load_library()
    n = last_stage - first_stage + 1
    counts = np.zeros(n, dtype=np.int64)
    pool_bytes, most = C.c_int64(0), C.c_int32(0)
    ncomp = lib.metis_enum_compositions(first_stage, last_stage, num_gpus, float(variance), max_permute_len,
                                        counts.ctypes.data, None, 0, None, 0, C.byref(pool_bytes), C.byref(most))
    if ncomp < 0:
        raise native.MetisNativeError(f'metis_enum_compositions failed ({ncomp})')
    recs = np.zeros(max(int(ncomp), 1), dtype=native.COMP_DTYPE)
    pool = np.zeros(max(int(pool_bytes.value), 16), dtype=np.uint8)
    got = lib.metis_enum_compositions(first_stage, last_stage, num_gpus, float(variance), max_permute_len,
                                      counts.ctypes.data, recs.ctypes.data, int(ncomp), pool.ctypes.data,
                                      int(pool_bytes.value), C.byref(pool_bytes), C.byref(most))
    if got != ncomp:
        raise native.MetisNativeError('metis_enum_compositions: inconsistent count')
    return counts, recs[:int(ncomp)], pool, int(most.value)
