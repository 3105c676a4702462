#!/usr/bin/env python3
"""bench.py - candidate plans evaluated / second on B200 (BASELINE.json metric).

A "step" is one full search of the workload's candidate space (every inter-stage plan enumerated by
InterStagePlanGenerator, its intra-stage chain, load balancer and cost model) by libmetis_b200.so.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

Workload: BASELINE.json configs[2] ("homo 64-GPU cluster, 96-layer GPT-3, gbs=512 (~10^6 candidates)
- 1xB200, HBM-roofline capture"), i.e. c3_homo64_mpl6 = 771 750 inter-stage plans, the largest
single-GPU configuration; configs[1] (16 GPUs, 1 752 plans) finishes in one wave of blocks and is a
parity-test case (tests/test_gpu_parity.py).  N > 1 shards the same space by plan ordinal over the
ranks (strong scaling) with one NCCL all_gather of 16-byte best records per step.

Timed regions
  value : tables + plan space already resident in HBM; per step CUDA events on the launching stream
          around pack + search + finalize (+ the NCCL exchange when N > 1); a 256 MiB write between
          steps flushes L2; ms_per_step = mean, max over ranks.
  e2e   : per step, from host inputs: host enumeration of the plan space (C++), H2D of every table
          from pinned memory, the same kernels, device sort + D2H of all 16-byte records, summary and
          the winner's strategies/partition; wall clock between device synchronisations.
  cpu_baseline / --impl reference : the oracle (Python port of the reference, oracle/metis_oracle.py)
          on a bounded random sample of the same plans, one process per host core.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402

METRIC = 'candidate plans evaluated/sec'
DEFAULT_WORKLOAD = 'c3_homo64_mpl6'


# ---------------------------------------------------------------------------------------------
# CPU side: oracle port on a bounded sample (cpu_baseline and --impl reference)
# ---------------------------------------------------------------------------------------------
_W = {}


def _cpu_worker_init(workload_name, root, share_seed, nproc, per_step, nsteps):
    """Each worker loads the inputs through the oracle and collects ITS sampled plans (untimed)."""
    import itertools
    import random
    from metis_b200.workloads import WORKLOADS, profile_file_order
    from oracle import metis_oracle as orc
    w = WORKLOADS[workload_name]
    cluster = orc.OracleCluster(os.path.join(root, 'hostfile'), os.path.join(root, 'clusterfile.json'))
    profile, _ = orc.load_profile_dir(os.path.join(root, 'profile'), profile_file_order(w))
    model = orc.OracleModel(w.num_layers, w.hidden_size, w.sequence_length, w.vocab_size,
                            profile['model']['parameters'])
    seqs = list(itertools.permutations(w.device_types()))
    _W.update(orc=orc, w=w, cluster=cluster, profile=profile, model=model,
              norm=orc.norm_layer_duration(profile), plans={}, seqs=seqs, per_step=per_step)
    _W['share'] = (share_seed, nproc, nsteps)


def _cpu_worker_collect(args):
    worker, total = args
    orc, w = _W['orc'], _W['w']
    seed, nproc, nsteps = _W['share']
    import random
    rng = random.Random(seed)
    want = rng.sample(range(total), min(total, nproc * nsteps * _W['per_step']))
    mine = {}
    for i, o in enumerate(want):
        if i % nproc == worker:
            mine[o] = (i // nproc) % nsteps                    # ordinal -> step index
    plans = {}
    for ordinal, plan in enumerate(orc.inter_stage_plans(_W['seqs'], _W['cluster'].total_devices, w.gbs,
                                                         w.num_layers, w.variance, w.max_permute_len)):
        if ordinal in mine:
            plans.setdefault(mine[ordinal], []).append((ordinal, dict(plan, device_groups=list(plan['device_groups']))))
    _W['plans'] = plans
    return sum(len(v) for v in plans.values())


def _cpu_worker_step(step):
    orc, w = _W['orc'], _W['w']
    counters = {'A': 0, 'B': 0, 'C': 0, 'runs': 0, 'keyerr': 0}
    out = []
    t0 = time.perf_counter()
    for ordinal, plan in _W['plans'].get(step, []):
        counters['A'] += 1
        orc.het_evaluate_plan(_W['profile'], _W['cluster'], _W['model'], _W['norm'], plan, ordinal, w.num_layers,
                              w.max_tp, w.max_bs, counters, out)
    return counters['A'], counters['C'], time.perf_counter() - t0


class CpuPort:
    """Pool of oracle workers over a bounded sample of the workload's plans."""

    def __init__(self, workload_name, total_plans, per_step_per_core, nsteps):
        import multiprocessing as mp
        from metis_b200.workloads import WORKLOADS, materialize
        self.cores = os.cpu_count() or 1
        self.tmp = tempfile.TemporaryDirectory()
        materialize(WORKLOADS[workload_name], self.tmp.name)
        self.pool = mp.get_context('spawn').Pool(self.cores, initializer=_cpu_worker_init,
                                                 initargs=(workload_name, self.tmp.name, 20240921, self.cores,
                                                           per_step_per_core, nsteps))
        self.collected = sum(self.pool.map(_cpu_worker_collect, [(k, total_plans) for k in range(self.cores)], 1))

    def step(self, idx):
        t0 = time.perf_counter()
        res = self.pool.map(_cpu_worker_step, [idx] * self.cores, 1)
        wall = time.perf_counter() - t0
        return sum(r[0] for r in res), sum(r[1] for r in res), wall

    def close(self):
        self.pool.close()
        self.pool.join()
        self.tmp.cleanup()


def count_plans_host(workload_name):
    """A (number of inter-stage plans) via the library's host enumerator - no GPU needed."""
    from metis_b200 import flatten
    from metis_b200.workloads import WORKLOADS
    import math
    w = WORKLOADS[workload_name]
    nseq = math.factorial(len(w.device_types()))
    ndev = sum(n for _, n in w.nodes)
    return flatten.build_plan_space(nseq, ndev, w.gbs, w.num_layers, w.variance, w.max_permute_len)


def run_reference_arm(ns):
    """--impl reference: the reference's CPU implementation of the path = the oracle port (the
    reference is pure Python and /root/reference does not exist on the GPU box), all host cores."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    space = count_plans_host(ns.workload)
    steps, warm = ns.steps, ns.warmup
    budget_s = 150.0
    per_step = max(40, int(budget_s / (steps + warm) * 350))           # ~350 plans/s/core in CPython
    port = CpuPort(ns.workload, space.num_plans, per_step, steps + warm)
    for i in range(warm):
        port.step(i)
    plans = costed = 0
    wall = 0.0
    for i in range(warm, warm + steps):
        a, c, t = port.step(i)
        plans, costed, wall = plans + a, costed + c, wall + t
    port.close()
    value = plans / wall
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'plans/s', 'n_gpus': ns.gpus,
        'steps': steps, 'warmup': warm, 'ms_per_step': 1e3 * wall / steps, 'higher_is_better': True,
        'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(ns.workload, space.num_plans),
        'cpu_baseline': {'value': value, 'unit': 'plans/s', 'cores': port.cores, 'kind': 'port',
                         'sample': f'{plans} uniformly sampled inter-stage plans of the {space.num_plans}-plan '
                                   f'space ({per_step} per core per step), oracle/metis_oracle.py, '
                                   f'{port.cores} processes'},
        'e2e': {'value': value, 'unit': 'plans/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'costed_per_s': costed / wall,
    }
    emit_result(line)


def workload_config(name, num_plans):
    from metis_b200.workloads import WORKLOADS
    w = WORKLOADS[name]
    return {'workload': f'{name}: {len(w.nodes)} nodes x {w.nodes[0][1]} GPUs ({"+".join(w.device_types())}), '
                        f'{w.num_layers} layers, gbs {w.gbs}, variance {w.variance}, max_permute_len '
                        f'{w.max_permute_len}, tp<= {w.max_tp}, bs<= {w.max_bs} (BASELINE.json configs[2])',
            'inter_stage_plans': int(num_plans), 'l2': 'flushed between timed steps (256 MiB write)',
            'parallelism': 'plans sharded by ordinal, interleaved 128-plan tiles'}


# ---------------------------------------------------------------------------------------------
# GPU side
# ---------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Streams `nvidia-smi -lms 50` for one GPU while the timed region runs (B200_PROFILING.md clocks line)."""

    QUERY = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.proc = None
        self.armed = threading.Event()

    def run(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.QUERY}',
                                          '--format=csv,noheader,nounits', '-lms', '50'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.armed.is_set() and line.strip():
                    self.rows.append([x.strip() for x in line.strip().split(',')])
        except Exception:
            pass

    def stop(self):
        self.armed.clear()
        if self.proc is not None:
            self.proc.terminate()

    def summary(self):
        if not self.rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        sm = [float(r[0]) for r in self.rows if r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith('active') for r in self.rows)]
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(self.rows)}


def run_ours(ns):
    import torch
    import torch.distributed as dist
    from metis_b200 import flatten, native, search
    from metis_b200.data_loader import ProfileDataLoader
    from metis_b200.gpu_cluster import GPUCluster
    from metis_b200.utils import ModelConfig
    from metis_b200.workloads import WORKLOADS, materialize, profile_file_order
    import itertools

    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm')
    native.load_library()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device(f'cuda:{local}')
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    assert world == ns.gpus or world == 1 and ns.gpus == 1, f'--gpus {ns.gpus} but WORLD_SIZE {world}'

    w = WORKLOADS[ns.workload]
    tmp = tempfile.TemporaryDirectory()
    materialize(w, tmp.name)
    cluster = GPUCluster(os.path.join(tmp.name, 'hostfile'), os.path.join(tmp.name, 'clusterfile.json'))
    profile, _ = ProfileDataLoader(os.path.join(tmp.name, 'profile'), profile_file_order(w)).load_profile_data_all()
    cfg = ModelConfig(model_name='SYN', num_layers=w.num_layers, sequence_length=w.sequence_length,
                      vocab_size=w.vocab_size, hidden_size=w.hidden_size, attention_head_size=32)
    seqs = list(itertools.permutations(w.device_types()))
    ndev = cluster.get_total_num_devices()

    def enumerate_space(rows_out=None):
        return flatten.build_plan_space(len(seqs), ndev, w.gbs, w.num_layers, w.variance, w.max_permute_len,
                                        rows_out=rows_out)

    t0 = time.perf_counter()
    problem = flatten.build_problem(profile, cluster, cfg, w.gbs, w.max_tp, w.max_bs, seqs)
    space = enumerate_space()
    host_prep_s = time.perf_counter() - t0
    dp = search.DeviceProblem(problem, space, dev)
    tile = 128
    fast = search.HetSearcher(dp, rank, world, tile, want_records=False)          # `value`: best only
    full = search.HetSearcher(dp, rank, world, tile, want_records=True, want_detail=False)   # e2e: all records
    stream = torch.cuda.current_stream(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def exchange(best):
        return search.global_best(best, dev) if world > 1 else best

    def gpu_step():
        fast.launch(stream)
        if world > 1:
            # the summary lands in pinned memory after the stream sync; the collective itself is tiny
            stream.synchronize()
            sm = fast.summary()
            b = sm.best
            lb = (b.cost, b.ordinal, b.step, b.num_repartition, b.num_stage) if sm.num_records else None
            return exchange(lb)
        return None

    # ---- warm-up (also proves the search result before timing) -----------------------------------
    for _ in range(max(ns.warmup, 3)):
        gpu_step()
    stream.synchronize()
    ref = full.run(stream)
    if ref.summary['fatal_ordinal'] != 2 ** 64 - 1:
        raise SystemExit(f'fatal plan {ref.summary}')
    if world > 1:
        counters = search.global_counters(ref.summary, dev)
        gbest = exchange(ref.best)
    else:
        counters, gbest = ref.summary, ref.best

    # ---- timed: K steps, device events, L2 flushed between steps ---------------------------------
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()                  # nvidia-smi is already streaming when the timed region starts
        time.sleep(0.3)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    if sampler:
        sampler.armed.set()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(ns.steps)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(ns.steps)]
    for a, b in kev:
        a.record(stream)
        b.record(stream)           # creates the handles; the library re-records them around the kernel
    wall0 = time.perf_counter()
    for i in range(ns.steps):
        flush.fill_(i & 0xFF)
        ev[i][0].record(stream)
        dp.lib.metis_set_profile_events(kev[i][0].cuda_event, kev[i][1].cuda_event)
        gpu_step()
        ev[i][1].record(stream)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    wall = time.perf_counter() - wall0
    if sampler:
        sampler.armed.clear()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    kern_ms = [a.elapsed_time(b) for a, b in kev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    kmean = torch.tensor([sum(kern_ms) / len(kern_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(kmean, op=dist.ReduceOp.MAX)
    ms_per_step = float(total_ms.item()) / ns.steps
    kernel_ms = float(kmean.item())

    # ---- e2e: host inputs -> result on the host, every step ---------------------------------------
    e2e_steps = max(3, min(ns.steps, 5))
    e2e_wall = []
    enum_ms = []
    d2h = 0
    for i in range(e2e_steps + 1):
        flush.fill_(i)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        sp2 = enumerate_space(dp.staging('rows'))             # host enumeration (C++) straight into pinned staging
        t1 = time.perf_counter()
        dp.restage_space(sp2)
        dp.upload(stream)                                     # H2D of every table (pinned -> HBM)
        out = full.run(stream)                                # kernels + device sort + D2H of all records
        best = exchange(out.best)
        if rank == 0 and best is not None and world == 1:
            picks = np.zeros(1, dtype=native.RECORD_DTYPE)
            picks['ordinal'], picks['step'] = best[1], best[2]
            full.detail_for(picks, stream)                    # winner's strategies + partition to the host
        torch.cuda.synchronize(dev)
        if i > 0:                                             # first iteration = warm-up
            e2e_wall.append(time.perf_counter() - t0)
            enum_ms.append(1e3 * (t1 - t0))
        d2h = out.d2h_bytes + native.DETAIL_STRIDE
        assert out.best == ref.best
    e2e_t = torch.tensor([statistics.mean(e2e_wall)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_t.item())

    if sampler:
        sampler.stop()
        sampler.join(timeout=3)

    A = space.num_plans
    if rank == 0:
        nst = space.blocks['num_stage'].astype(np.int64)
        plans_per_block = space.blocks['num_rows'].astype(np.int64) * len(space.batches)
        alg_bytes = int((plans_per_block * (nst + 16)).sum() + 16 * counters['num_records'])
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        peak = float(peaks.get('hbm_gbs', 6650.0))
        achieved = alg_bytes / world / (kernel_ms * 1e-3) / 1e9
        traffic = None                                        # measured for the 1-GPU launch only
        tpath = os.path.join(REPO, 'profiles', 'r01_traffic.json')
        if os.path.exists(tpath) and world == 1:
            traffic = json.load(open(tpath)).get('dram_bytes_per_launch')
        line = {
            'metric': METRIC, 'value': A / (ms_per_step * 1e-3), 'unit': 'plans/s', 'n_gpus': world,
            'steps': ns.steps, 'warmup': max(ns.warmup, 3), 'ms_per_step': ms_per_step, 'higher_is_better': True,
            'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': workload_config(ns.workload, A),
            'counters': {'A_inter_stage_plans': A, 'B_partition_layer_calls': counters['num_partition_calls'],
                         'balancer_runs': counters['num_balancer_runs'], 'C_costed': counters['num_records'],
                         'keyerror': counters['num_keyerror']},
            'best_plan': {'cost': gbest[0], 'ordinal': gbest[1], 'step': gbest[2]} if gbest else None,
            'costed_per_s': counters['num_records'] / (ms_per_step * 1e-3),
            'time_to_best_ms': {'gpu_resident': ms_per_step, 'end_to_end': 1e3 * e2e_s,
                                'host_enumeration': statistics.mean(enum_ms), 'host_flatten_once': 1e3 * host_prep_s},
            'e2e': {'value': A / e2e_s, 'unit': 'plans/s', 'h2d_bytes_per_step': int(dp.h2d_bytes),
                    'd2h_bytes_per_step': int(d2h), 'ms_per_step': 1e3 * e2e_s, 'steps': e2e_steps,
                    'timing': 'wall clock between device synchronisations, max over ranks; includes host '
                              'enumeration of the plan space'},
            'gpu_launches': 3 * ns.steps,
            'kernel_ms': {'het_search_kernel_mean': kernel_ms, 'step_mean': ms_per_step,
                          'step_min': min(step_ms), 'step_max': max(step_ms)},
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                         'frac': achieved / peak, 'traffic': traffic,
                         'peak_source': 'MEASURED_PEAKS.json hbm_gbs (measured)' if 'hbm_gbs' in peaks else 'fallback',
                         'algorithmic_bytes_per_launch': alg_bytes // world,
                         'note': 'S+16 B read per inter-stage plan + 16 B written per costed candidate '
                                 '(SURVEY.md 8d); the path is fp64-latency / divergence bound, not HBM bound'},
            'clocks': sampler.summary() if sampler else None,
            'wall_s_timed_region': wall,
        }
        if world == 1 and not ns.no_cpu:
            port = CpuPort(ns.workload, A, ns.cpu_sample, 1)
            a, c, t = port.step(0)
            port.close()
            line['cpu_baseline'] = {'value': a / t, 'unit': 'plans/s', 'cores': port.cores, 'kind': 'port',
                                    'sample': f'{a} uniformly sampled inter-stage plans of the same {A}-plan space '
                                              f'({c} costed), oracle/metis_oracle.py (Python port of the pure-Python '
                                              f'reference), {port.cores} processes, {t:.1f} s'}
        emit_result(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    tmp.cleanup()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default=DEFAULT_WORKLOAD)
    ap.add_argument('--cpu-sample', type=int, default=3000, help='plans per host core for cpu_baseline')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ns = ap.parse_args()
    # stdout carries exactly one JSON line: libraries that write to fd 1 (NCCL prints its version there when
    # NCCL_DEBUG=VERSION) are sent to stderr for the duration of the run
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    if ns.impl == 'reference':
        run_reference_arm(ns)
    else:
        run_ours(ns)


_RESULT_FD = None


def emit_result(line):
    data = (json.dumps(line) + '\n').encode()
    sys.stdout.flush()
    if _RESULT_FD is None:
        os.write(1, data)
    else:
        os.write(_RESULT_FD, data)


if __name__ == '__main__':
    main()
