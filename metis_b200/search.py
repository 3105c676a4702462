"""Device-side plan search: the B200 replacement of the loops in cost_het_cluster.py:21-50
and cost_homo_cluster.py:21-37 of the reference.

PyTorch is used only for device buffers, streams and (multi-GPU) torch.distributed; all
search arithmetic runs in libmetis_b200.so (hand-written sm_100a CUDA) behind the C ABI of
include/metis_b200.h.  There is no CPU path: without CUDA these functions raise.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import flatten, native


def _require_cuda(device) -> torch.device:
    if not torch.cuda.is_available():
        raise native.MetisNativeError('CUDA device required: metis_b200 has no CPU fallback')
    return torch.device(device if device is not None else f'cuda:{torch.cuda.current_device()}')


class DeviceProblem:
    """A flattened problem + candidate space resident in HBM (one upload, many searches)."""

    def __init__(self, problem: flatten.FlatProblem, space: flatten.FlatPlanSpace, device=None,
                 pinned: bool = True):
        self.device = _require_cuda(device)
        self.lib = native.load_library()
        self.problem = problem
        self.space = space
        self._host: Dict[str, torch.Tensor] = {}
        self._dev: Dict[str, torch.Tensor] = {}
        arrays = dict(problem.arrays)
        arrays.update(blocks=space.blocks.view(np.uint8).reshape(-1), batches=space.batches, rows=space.rows)
        self.h2d_bytes = 0
        for name, arr in arrays.items():
            flat = np.ascontiguousarray(arr).view(np.uint8).reshape(-1)
            host = torch.from_numpy(flat.copy() if flat.size else np.zeros(16, dtype=np.uint8))
            if pinned:
                host = host.pin_memory()
            self._host[name] = host
            self.h2d_bytes += host.numel()
        self.upload()

    def staging(self, name: str) -> np.ndarray:
        """The pinned host copy of one table (numpy view): fill it in place, then upload()."""
        return self._host[name].numpy()

    def restage_space(self, space: flatten.FlatPlanSpace) -> None:
        """Put a freshly enumerated space of the same shape into the staging buffers; tables that
        build_plan_space(rows_out=staging('rows')) already wrote in place are not copied again."""
        rows, blocks = self.staging('rows'), self.staging('blocks')
        if space.rows.size > rows.size or space.blocks.nbytes != blocks.size:
            raise ValueError('space does not fit the staging buffers of this DeviceProblem')
        if space.rows.size and not np.shares_memory(space.rows, rows):
            rows[:space.rows.size] = space.rows
        blocks[:] = space.blocks.view(np.uint8).reshape(-1)
        self.space = space

    def upload(self, stream: Optional[torch.cuda.Stream] = None) -> None:
        """Host -> HBM copy of every table (part of the end-to-end timed region)."""
        with torch.cuda.device(self.device), torch.cuda.stream(stream or torch.cuda.current_stream(self.device)):
            for name, host in self._host.items():
                dev = self._dev.get(name)
                if dev is None:
                    self._dev[name] = host.to(self.device, non_blocking=True)
                else:
                    dev.copy_(host, non_blocking=True)
        self.p_struct = self.problem.as_struct(lambda n: self._dev[n].data_ptr())
        self.s_struct = self.space.as_struct(lambda n: self._dev[n].data_ptr())

    def workspace_bytes(self, num_plans: int) -> int:
        n = self.lib.metis_het_workspace_bytes(C.byref(self.p_struct), num_plans, self.s_struct.max_stage)
        if n < 0:
            native.check(int(n), 'metis_het_workspace_bytes')
        return int(n)


@dataclass
class HetSearchOutput:
    """Result of one shard's search (numpy, host)."""
    summary: Dict[str, int]
    best: Optional[Tuple[float, int, int, int, int]]      # cost, ordinal, step, num_repartition, num_stage
    records: Optional[np.ndarray]                         # native.RECORD_DTYPE sorted by (ordinal, step)
    detail: Optional[np.ndarray]                          # uint8 [n, DETAIL_STRIDE] aligned with records
    d2h_bytes: int = 0
    rank_order: Optional[np.ndarray] = None               # uint32: records[rank_order] = sorted(..., key=cost), stable


class HetSearcher:
    """Owns the output buffers for repeated searches over one DeviceProblem."""

    def __init__(self, dp: DeviceProblem, rank: int = 0, world: int = 1, tile: int = 128,
                 want_records: bool = True, want_detail: bool = False, capacity: Optional[int] = None,
                 want_ranking: bool = False):
        self.dp = dp
        self.want_ranking = want_ranking and want_records
        self._sort_ws = None
        self.shard = native.MetisShard(rank, world, tile, 0)
        self.want_records = want_records
        self.want_detail = want_detail and want_records
        rounds = -(-dp.space.num_plans // (tile * world))
        self.shard_plans = rounds * tile
        self.capacity = 0
        dev = dp.device
        self.workspace = torch.empty(dp.workspace_bytes(self.shard_plans), dtype=torch.uint8, device=dev)
        self.summary_host = torch.zeros(C.sizeof(native.MetisSearchSummary), dtype=torch.uint8).pin_memory()
        self.records = self.detail = None
        if want_records:
            self._alloc(capacity if capacity is not None else self.shard_plans + 4096)

    def _alloc(self, capacity: int) -> None:
        dev = self.dp.device
        self.capacity = capacity
        self.records = torch.empty(capacity * 2, dtype=torch.int64, device=dev)
        self.detail = (torch.empty((capacity, native.DETAIL_STRIDE), dtype=torch.uint8, device=dev)
                       if self.want_detail else None)

    def launch(self, stream: Optional[torch.cuda.Stream] = None) -> None:
        """Enqueue pack + search + finalize + summary copy on ``stream`` (asynchronous)."""
        dp = self.dp
        s = stream or torch.cuda.current_stream(dp.device)
        rc = dp.lib.metis_het_search(
            C.byref(dp.p_struct), C.byref(dp.s_struct), C.byref(self.shard),
            C.c_void_p(self.records.data_ptr() if self.records is not None else 0), C.c_int64(self.capacity),
            C.c_void_p(self.detail.data_ptr() if self.detail is not None else 0), C.c_int32(native.DETAIL_STRIDE),
            C.c_void_p(self.workspace.data_ptr()), C.c_int64(self.workspace.numel()),
            C.c_void_p(self.summary_host.data_ptr()), C.c_void_p(s.cuda_stream))
        native.check(rc, 'metis_het_search')

    def summary(self) -> native.MetisSearchSummary:
        return native.MetisSearchSummary.from_buffer_copy(self.summary_host.numpy().tobytes())

    def run(self, stream: Optional[torch.cuda.Stream] = None) -> HetSearchOutput:
        """launch + synchronise + bring results to the host; grows the record buffer if needed."""
        dp = self.dp
        s = stream or torch.cuda.current_stream(dp.device)
        with torch.cuda.device(dp.device):
            self.launch(s)
            s.synchronize()
            sm = self.summary()
            if sm.fatal_code == native.FATAL_SCHEDULER:
                raise native.MetisNativeError('device task queue watchdog fired (loop, ticket, slot value, head, tail, '
                                              f'alive) = {[int(v) for v in sm.reserved]}')
            if self.want_records and sm.num_records > self.capacity:
                self._alloc(int(sm.num_records) + 1024)
                self.launch(s)
                s.synchronize()
                sm = self.summary()
            out_summary = dict(num_records=int(sm.num_records), num_partition_calls=int(sm.num_partition_calls),
                               num_balancer_runs=int(sm.num_balancer_runs), num_keyerror=int(sm.num_keyerror),
                               fatal_ordinal=int(sm.fatal_ordinal), fatal_code=int(sm.fatal_code),
                               fatal_aux=int(sm.fatal_aux))
            best = None
            if sm.num_records > 0:
                b = sm.best
                best = (float(b.cost), int(b.ordinal), int(b.step), int(b.num_repartition), int(b.num_stage))
            records = detail = rank_order = None
            d2h = C.sizeof(native.MetisSearchSummary)
            if self.want_records:
                n = int(sm.num_records)
                # estimate_costs order = (ordinal, step): rank_records_kernel, in place
                order = self.sort_records(n, native.SORT_POSITION, s, want_perm=self.want_detail)
                records = self.records[:2 * n].cpu().numpy().view(np.uint8).reshape(-1).view(native.RECORD_DTYPE)
                d2h += n * 16
                if self.want_detail:
                    detail = self.detail[:n].index_select(0, order.long()).cpu().numpy()
                    d2h += n * native.DETAIL_STRIDE
                if self.want_ranking:
                    # sorted(estimate_costs, key=cost): stable by cost on a copy of the ordered records
                    by_cost = self.records[:2 * n].clone()
                    rank_order = self.sort_records(n, native.SORT_BY_COST_STABLE, s, want_perm=True,
                                                   buf=by_cost).cpu().numpy().view(np.uint32)
                    d2h += n * 4
        return HetSearchOutput(out_summary, best, records, detail, d2h, rank_order)

    def sort_records(self, n: int, mode: int, stream: torch.cuda.Stream, want_perm: bool = False, buf=None):
        """metis_sort_records on the first n records (device, in place); returns the permutation tensor
        (int32 view of the uint32 indices) when asked."""
        dp = self.dp
        need = int(dp.lib.metis_sort_workspace_bytes(C.c_int64(n)))
        if self._sort_ws is None or self._sort_ws.numel() < need:
            self._sort_ws = torch.empty(need, dtype=torch.uint8, device=dp.device)
        perm = torch.empty(max(n, 1), dtype=torch.int32, device=dp.device) if want_perm else None
        buf = self.records if buf is None else buf
        rc = dp.lib.metis_sort_records(C.c_void_p(buf.data_ptr()), C.c_int64(n), C.c_int32(mode),
                                       C.c_void_p(perm.data_ptr() if perm is not None else 0),
                                       C.c_void_p(self._sort_ws.data_ptr()), C.c_int64(self._sort_ws.numel()),
                                       C.c_void_p(stream.cuda_stream))
        native.check(rc, 'metis_sort_records')
        stream.synchronize()
        return perm[:n] if perm is not None else None

    def detail_for(self, picks: np.ndarray, stream: Optional[torch.cuda.Stream] = None) -> np.ndarray:
        """Strategies and partition of chosen records (metis_het_detail replay)."""
        dp = self.dp
        s = stream or torch.cuda.current_stream(dp.device)
        n = len(picks)
        with torch.cuda.device(dp.device):
            raw = torch.from_numpy(np.ascontiguousarray(picks).view(np.uint8).reshape(-1).copy()).to(dp.device)
            out = torch.zeros((max(n, 1), native.DETAIL_STRIDE), dtype=torch.uint8, device=dp.device)
            rc = dp.lib.metis_het_detail(C.byref(dp.p_struct), C.byref(dp.s_struct), C.c_void_p(raw.data_ptr()),
                                         C.c_int64(n), C.c_void_p(out.data_ptr()), C.c_int32(native.DETAIL_STRIDE),
                                         C.c_void_p(self.workspace.data_ptr()), C.c_int64(self.workspace.numel()),
                                         C.c_void_p(s.cuda_stream))
            native.check(rc, 'metis_het_detail')
            s.synchronize()
            return out[:n].cpu().numpy()


def raise_fatal(summary: Dict[str, int], problem: flatten.FlatProblem) -> None:
    """Re-raise what the reference would have raised at the first failing plan (quirk Q8)."""
    code = summary['fatal_code']
    if summary['fatal_ordinal'] == 2 ** 64 - 1 or code == 0:
        return
    aux = summary['fatal_aux']
    tp, bs = 1 << ((aux >> 16) & 0xFF), aux & 0xFFFF
    if code in (1, 2):
        raise KeyError(f'tp{tp}_bs{bs}')
    if code == 3:
        raise IndexError('list index out of range')
    if code == 6:
        raise ZeroDivisionError('float division by zero')
    raise RuntimeError(f'plan {summary["fatal_ordinal"]}: {native.FATAL_NAMES.get(code, code)} '
                       f'(the reference does not complete this search either)')


def materialize(records: np.ndarray, detail: np.ndarray, space: flatten.FlatPlanSpace,
                node_sequences: Sequence[Tuple]) -> List[Tuple]:
    """Records -> the reference's 7-tuples (cost_het_cluster.py:44-46)."""
    out = []
    ndiv = len(space.batches)
    firsts = space.blocks['first_ordinal']
    blk_of = np.searchsorted(firsts, records['ordinal'].astype(np.int64), side='right') - 1
    for i in range(len(records)):
        blk = space.blocks[blk_of[i]]
        rel = int(records['ordinal'][i]) - int(blk['first_ordinal'])
        row, div = divmod(rel, ndiv)
        S = int(blk['num_stage'])
        table = space.tables[S][1]
        d = detail[i]
        groups = [1 << int(c) for c in table[row]]
        strategies = [(1 << int(d[s]), 1 << int(d[S + s])) for s in range(S)]
        part = [int(x) for x in d[2 * S:3 * S + 1]]
        out.append((node_sequences[int(blk['ns_idx'])], groups, strategies, int(space.batches[div]), part,
                    int(records['num_repartition'][i]), float(records['cost'][i])))
    return out


# ---------------------------------------------------------------------------------------------
# multi-GPU: shard by plan ordinal, one collective at the end
# ---------------------------------------------------------------------------------------------
def global_best(local_best: Optional[Tuple[float, int, int, int, int]], device) -> Optional[Tuple]:
    """all_gather of one 16-byte (cost, ordinal/step) record per rank, then exact lexicographic min."""
    import torch.distributed as dist
    world = dist.get_world_size()
    mine = torch.zeros(4, dtype=torch.float64, device=device)
    if local_best is not None:
        cost, ordinal, step, nrep, nstage = local_best
        mine[0], mine[1], mine[2], mine[3] = cost, float(ordinal), float(step), float(nrep * 256 + nstage)
    else:
        mine[0], mine[1] = float('inf'), float(2 ** 40)
    allb = torch.empty(4 * world, dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(allb, mine)
    rows = allb.view(world, 4).cpu().tolist()
    rows = [r for r in rows if r[1] < 2 ** 40]
    if not rows:
        return None
    c, o, s, m = min(rows, key=lambda r: (r[0], r[1], r[2]))
    return (c, int(o), int(s), int(m) // 256, int(m) % 256)


def global_counters(summary: Dict[str, int], device) -> Dict[str, int]:
    import torch.distributed as dist
    keys = ['num_records', 'num_partition_calls', 'num_balancer_runs', 'num_keyerror']
    t = torch.tensor([summary[k] for k in keys], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    out = dict(summary)
    out.update({k: int(v) for k, v in zip(keys, t.cpu().tolist())})
    f = torch.tensor([min(summary['fatal_ordinal'], 2 ** 62)], dtype=torch.int64, device=device)
    dist.all_reduce(f, op=dist.ReduceOp.MIN)
    out['global_fatal_ordinal'] = int(f.item())
    return out


# ---------------------------------------------------------------------------------------------
# homogeneous path
# ---------------------------------------------------------------------------------------------
def homo_costs(problem: flatten.FlatProblem, type_id: int, plans: np.ndarray, device=None
               ) -> Tuple[np.ndarray, np.ndarray]:
    """HomoCostEstimator.get_cost for every row (dp, pp, tp, mbs, gbs) of ``plans`` on the GPU."""
    dev = _require_cuda(device)
    lib = native.load_library()
    with torch.cuda.device(dev):
        tens = {k: torch.from_numpy(np.ascontiguousarray(v).view(np.uint8).reshape(-1).copy()).to(dev)
                for k, v in problem.arrays.items()}
        p = problem.as_struct(lambda n: tens[n].data_ptr())
        ws = torch.empty(int(lib.metis_het_workspace_bytes(C.byref(p), 0, 1)), dtype=torch.uint8, device=dev)
        n = len(plans)
        d_plans = torch.from_numpy(np.ascontiguousarray(plans, dtype=np.int32).reshape(-1)).to(dev)
        cost = torch.zeros(max(n, 1), dtype=torch.float64, device=dev)
        status = torch.zeros(max(n, 1), dtype=torch.int32, device=dev)
        s = torch.cuda.current_stream(dev)
        rc = lib.metis_homo_cost(C.byref(p), C.c_int32(type_id), C.c_void_p(d_plans.data_ptr()), C.c_int64(n),
                                 C.c_void_p(cost.data_ptr()), C.c_void_p(status.data_ptr()),
                                 C.c_void_p(ws.data_ptr()), C.c_int64(ws.numel()), C.c_void_p(s.cuda_stream))
        native.check(rc, 'metis_homo_cost')
        s.synchronize()
        return cost[:n].cpu().numpy(), status[:n].cpu().numpy()


def layer_balance(capa_rows: Sequence[Sequence[float]], lc: Sequence[float], num_layers: int, device=None
                  ) -> List[List[int]]:
    """LayerComputeBalancer.run for many capacity vectors on the GPU (unit-level entry point)."""
    dev = _require_cuda(device)
    lib = native.load_library()
    n = len(capa_rows)
    stride = max(len(c) for c in capa_rows)
    capa = np.zeros((n, stride))
    ns = np.zeros(n, dtype=np.int32)
    for i, c in enumerate(capa_rows):
        capa[i, :len(c)] = c
        ns[i] = len(c)
    with torch.cuda.device(dev):
        d_capa = torch.from_numpy(capa).to(dev)
        d_ns = torch.from_numpy(ns).to(dev)
        d_lc = torch.tensor(list(lc), dtype=torch.float64, device=dev)
        out = torch.zeros((n, stride + 1), dtype=torch.int16, device=dev)
        ws = torch.empty(len(lc) * 8 + 512, dtype=torch.uint8, device=dev)
        s = torch.cuda.current_stream(dev)
        rc = lib.metis_layer_balance(C.c_void_p(d_capa.data_ptr()), C.c_void_p(d_ns.data_ptr()), C.c_int64(n),
                                     C.c_int32(stride), C.c_void_p(d_lc.data_ptr()), C.c_int32(len(lc)),
                                     C.c_int32(num_layers), C.c_void_p(out.data_ptr()), C.c_void_p(ws.data_ptr()),
                                     C.c_int64(ws.numel()), C.c_void_p(s.cuda_stream))
        native.check(rc, 'metis_layer_balance')
        s.synchronize()
        res = out.cpu().numpy().view(np.uint16)
    return [res[i, :ns[i] + 1].tolist() for i in range(n)]
