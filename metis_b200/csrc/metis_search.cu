// metis_search.cu - sm_100a kernels + C ABI of libmetis_b200.so (see include/metis_b200.h).
//
// Kernel map:
//   pack_tables_kernel    flattens the profile tables into one 16 B-aligned blob (+ norm_lc/7)
//   het_search_kernel     a2..a16: cooperative persistent kernel over task lists (a task = one partition
//                         attempt of one plan): admission, counting sort of the first list, lockstep rounds
//                         (32 tasks per warp) and a barrier-free queue (one task per warp) - DESIGN.md 4.1;
//                         tables staged into shared memory by one TMA bulk copy (cp.async.bulk + mbarrier)
//                         per block; 16 B record per costed candidate; warp-shuffle + block argmin
//   het_finalize_kernel   grid argmin over the per-block bests, counters -> summary
//   het_detail_kernel     replays chosen (ordinal, step) candidates to materialise strategies/partition
//   homo_cost_kernel      a17: one thread per UniformPlan
//   layer_balance_kernel  a10 alone, for unit parity
//   (rank_records_kernel, the stable record sort, lives in metis_rank.cu)
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false (no FMA contraction: parity).
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "metis_eval.cuh"

namespace cg = cooperative_groups;

namespace metis {

// block shape measured on B200 (profiles/): 256 threads, >= 3 blocks/SM (80 registers) beat 128 x 6 and 256 x 4
#ifndef METIS_THREADS
#define METIS_THREADS 256
#endif
#ifndef METIS_MIN_BLOCKS
#define METIS_MIN_BLOCKS 3
#endif
constexpr int kThreads = METIS_THREADS;
constexpr int kMaxS = METIS_MAX_STAGES;
constexpr int kMaxL = METIS_MAX_LAYERS;
constexpr int kSmemBlobMax = 160 * 1024;

static thread_local char g_err[256] = "";
static thread_local cudaEvent_t g_ev_before = nullptr, g_ev_after = nullptr;

static int cuda_fail(cudaError_t e, const char *what) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return METIS_E_CUDA;
}
static int arg_fail(const char *what) {
    snprintf(g_err, sizeof(g_err), "%s", what);
    return METIS_E_ARG;
}
int fail_cuda(cudaError_t e, const char *what) { return cuda_fail(e, what); }   // metis_internal.h
int fail_arg(const char *what) { return arg_fail(what); }

struct BlobLayout {
    uint32_t total;
    uint32_t key, lc, mem, exec_full, fb, norm, derived, tmem, bwf, bwm, runt, rune;
};

static uint32_t align16(uint32_t v) { return (v + 15u) & ~15u; }

static BlobLayout make_layout(const MetisProblem &p) {
    BlobLayout l;
    uint32_t o = 0;
    const uint32_t nkey = (uint32_t)p.num_types * p.num_tp * p.num_bs;
    l.key = o;       o = align16(o + nkey * 2);
    l.lc = o;        o = align16(o + (uint32_t)p.num_keys * p.lpad * 8);
    l.mem = o;       o = align16(o + (uint32_t)p.num_keys * p.lpad * 8);
    l.exec_full = o; o = align16(o + (uint32_t)p.num_keys * 8);
    l.fb = o;        o = align16(o + (uint32_t)p.num_keys * 8);
    l.norm = o;      o = align16(o + (uint32_t)p.norm_len * 8);
    l.derived = o;   o = align16(o + (uint32_t)derived_layout(p).total * 8);
    l.tmem = o;      o = align16(o + (uint32_t)p.num_types * 8);
    l.bwf = o;       o = align16(o + (uint32_t)p.num_types * 8);
    l.bwm = o;       o = align16(o + (uint32_t)p.num_types * 8);
    l.runt = o;      o = align16(o + (uint32_t)p.num_node_sequences * p.num_types);
    l.rune = o;      o = align16(o + (uint32_t)p.num_node_sequences * p.num_types * 4);
    l.total = o;
    return l;
}

__device__ __forceinline__ Tables make_tables(const MetisProblem &p, const BlobLayout &l, const uint8_t *base) {
    Tables T;
    T.p = p;
    T.key_index = reinterpret_cast<const int16_t *>(base + l.key);
    T.lc = reinterpret_cast<const double *>(base + l.lc);
    T.mem = reinterpret_cast<const double *>(base + l.mem);
    T.exec_full = reinterpret_cast<const double *>(base + l.exec_full);
    T.fb_sync = reinterpret_cast<const double *>(base + l.fb);
    T.norm_lc = reinterpret_cast<const double *>(base + l.norm);
    bind_derived(T, reinterpret_cast<const double *>(base + l.derived));
    T.type_memory = reinterpret_cast<const double *>(base + l.tmem);
    T.bw_first = reinterpret_cast<const double *>(base + l.bwf);
    T.bw_min = reinterpret_cast<const double *>(base + l.bwm);
    T.run_type = base + l.runt;
    T.run_end = reinterpret_cast<const int32_t *>(base + l.rune);
    return T;
}

__device__ __forceinline__ void copy_bytes(uint8_t *dst, const void *src, uint32_t n, uint32_t tid, uint32_t nthr) {
    const uint8_t *s = static_cast<const uint8_t *>(src);
    for (uint32_t i = tid; i < n; i += nthr) dst[i] = s[i];
}

__global__ void pack_tables_kernel(MetisProblem p, BlobLayout l, uint8_t *blob) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    const uint32_t nkey = (uint32_t)p.num_types * p.num_tp * p.num_bs;
    copy_bytes(blob + l.key, p.key_index, nkey * 2, tid, nthr);
    copy_bytes(blob + l.lc, p.layer_compute, (uint32_t)p.num_keys * p.lpad * 8, tid, nthr);
    copy_bytes(blob + l.mem, p.layer_memory, (uint32_t)p.num_keys * p.lpad * 8, tid, nthr);
    copy_bytes(blob + l.exec_full, p.exec_full, (uint32_t)p.num_keys * 8, tid, nthr);
    copy_bytes(blob + l.fb, p.fb_sync, (uint32_t)p.num_keys * 8, tid, nthr);
    copy_bytes(blob + l.norm, p.norm_lc, (uint32_t)p.norm_len * 8, tid, nthr);
    copy_bytes(blob + l.tmem, p.type_memory, (uint32_t)p.num_types * 8, tid, nthr);
    copy_bytes(blob + l.bwf, p.type_bw_first, (uint32_t)p.num_types * 8, tid, nthr);
    copy_bytes(blob + l.bwm, p.type_bw_min, (uint32_t)p.num_types * 8, tid, nthr);
    copy_bytes(blob + l.runt, p.ns_run_type, (uint32_t)p.num_node_sequences * p.num_types, tid, nthr);
    copy_bytes(blob + l.rune, p.ns_run_end, (uint32_t)p.num_node_sequences * p.num_types * 4, tid, nthr);
    // derived tables: each entry is one IEEE operation of the reference, evaluated once per launch
    double *derived = reinterpret_cast<double *>(blob + l.derived);
    const DerivedLayout d = derived_layout(p);
    for (uint32_t i = tid; i < (uint32_t)d.total; i += nthr)
        derived[i] = derive_entry(p, d, p.norm_lc, p.exec_full, p.type_bw_first, (int)i);
}

// ---- TMA 1-D bulk copy global -> shared, completion on an mbarrier --------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void stage_blob_tma(uint8_t *smem, const uint8_t *blob, uint32_t bytes, uint64_t *mbar) {
    const uint32_t bar = smem_u32(mbar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        constexpr uint32_t kChunk = 32768;
        for (uint32_t off = 0; off < bytes; off += kChunk) {
            const uint32_t n = (bytes - off < kChunk) ? bytes - off : kChunk;
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                ::"r"(smem_u32(smem + off)), "l"(blob + off), "r"(n), "r"(bar)
                : "memory");
        }
    }
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar)
            : "memory");
    }
}

// ---- ordinal -> plan ---------------------------------------------------------------------------
__device__ __forceinline__ int find_block(const MetisPlanSpace &sp, int64_t ordinal) {
    int lo = 0, hi = sp.num_blocks - 1;
    while (lo < hi) {                                         // last block with first_ordinal <= ordinal
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(&sp.blocks[mid].first_ordinal) <= ordinal) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// `hint` >= 0: a block known to start at or before `ordinal` (the warp's first plan); blocks are walked
// forward from it, so the 32 consecutive plans of a warp cost one binary search instead of 32.
__device__ __forceinline__ bool decode_plan(const MetisPlanSpace &sp, int64_t ordinal, PlanDesc &pd, int hint = -1) {
    if (ordinal < 0 || ordinal >= sp.num_plans) return false;
    int lo;
    if (hint >= 0) {
        lo = hint;
        while (lo + 1 < sp.num_blocks && __ldg(&sp.blocks[lo + 1].first_ordinal) <= ordinal) ++lo;
    } else {
        lo = find_block(sp, ordinal);
    }
    const MetisPlanBlock b = sp.blocks[lo];
    const int64_t rel = ordinal - b.first_ordinal;
    const int64_t row = rel / sp.num_div;
    const int div = (int)(rel - row * sp.num_div);
    pd.ordinal = (uint32_t)ordinal;
    pd.ns = b.ns_idx;
    pd.S = b.num_stage;
    pd.label = b.label_stage;
    pd.batches = __ldg(&sp.batches[div]);
    const int64_t off = b.rows_offset + row * b.num_stage;
    pd.row = sp.rows + off;
    pd.geo = pack_geo(off, b.num_stage, b.label_stage, b.ns_idx, div);
    return true;
}

// task list entry -> plan (no block search: the geometry word was stored at admission)
__device__ __forceinline__ void decode_task(const MetisPlanSpace &sp, uint64_t hdr, uint64_t geo, PlanDesc &pd) {
    pd.ordinal = (uint32_t)hdr;
    pd.geo = geo;
    pd.row = sp.rows + (geo & 0xFFFFFFFFULL);
    pd.S = (int)((geo >> 32) & 0xFF) + 1;
    pd.label = (int)((geo >> 40) & 0xFF) + 1;
    pd.ns = (int)((geo >> 48) & 0xFF);
    pd.batches = __ldg(&sp.batches[(geo >> 56) & 0xFF]);
}

__device__ __forceinline__ bool rec_less(double c0, uint32_t o0, uint32_t s0, double c1, uint32_t o1, uint32_t s1) {
    if (c0 < c1) return true;
    if (c0 > c1) return false;
    if (o0 != o1) return o0 < o1;
    return s0 < s1;
}

struct DeviceOut {
    MetisRecord *records;
    long long capacity;
    uint8_t *detail;
    int detail_stride;
    unsigned long long *counters;   // [0] records [1] partition calls [2] balancer runs [3] keyerrors [4] fatal key
    MetisRecord *block_best;
};

struct DeviceSink {
    const DeviceOut &o;
    unsigned int n_part, n_run, n_key;
    double best_cost;
    uint32_t best_ord, best_step, best_meta;
    bool leader;                    // cooperative mode: only lane 0 of the warp produces side effects
    __device__ DeviceSink(const DeviceOut &out)
        : o(out), n_part(0), n_run(0), n_key(0), best_cost(INFINITY), best_ord(0xFFFFFFFFu), best_step(0xFFFFu),
          best_meta(0), leader(true) {}
#ifdef METIS_PROFILE_PHASES
    long long t_last = 0; int cur = -1; long long acc[6] = {0, 0, 0, 0, 0, 0};
    __device__ void phase(int k) {
        if (!leader) return;
        const long long now = clock64();
        if (cur >= 0) acc[cur] += now - t_last;
        cur = k; t_last = now;
    }
#else
    __device__ void phase(int) {}
#endif
    __device__ void partition_call() { n_part += leader ? 1u : 0u; }
    __device__ void balancer_run() { n_run += leader ? 1u : 0u; }
    __device__ void keyerror() { n_key += leader ? 1u : 0u; }
    __device__ void fatal(uint32_t ordinal, int code, uint32_t aux) {
        if (!leader) return;
        const unsigned long long key = ((unsigned long long)ordinal << 32) | ((unsigned long long)(code & 0xFF) << 24) |
                                       (unsigned long long)(((aux >> 16) & 0xFF) << 16) | (aux & 0xFFFF);
        atomicMin(&o.counters[4], key);
    }
    __device__ void emit(const PlanDesc &pd, int step, int nrep, double cost, const uint8_t *tpc, const uint16_t *part) {
        if (!leader) return;
        const unsigned long long slot = atomicAdd(&o.counters[0], 1ULL);
        if ((long long)slot < o.capacity) {
            MetisRecord r;
            r.cost = cost; r.ordinal = pd.ordinal; r.step = (uint16_t)step;
            r.num_repartition = (uint8_t)nrep; r.num_stage = (uint8_t)pd.S;
            o.records[slot] = r;
            if (o.detail) {
                uint8_t *d = o.detail + (size_t)slot * o.detail_stride;
                for (int s = 0; s < pd.S; ++s) { d[s] = (uint8_t)(pd.row[s] - tpc[s]); d[pd.S + s] = tpc[s]; }
                for (int s = 0; s <= pd.S; ++s) d[2 * pd.S + s] = (uint8_t)part[s];
            }
        }
        if (rec_less(cost, pd.ordinal, (uint32_t)step, best_cost, best_ord, best_step)) {
            best_cost = cost; best_ord = pd.ordinal; best_step = (uint32_t)step;
            best_meta = ((uint32_t)nrep << 8) | (uint32_t)pd.S;
        }
    }
};

// Warp-level services of begin_task / run_task: ballot + one aggregated atomicAdd per warp hands
// out consecutive slots of the next round's task list (so the state stores are coalesced).
struct DeviceWarp {
    unsigned int *counter;
    __device__ explicit DeviceWarp(unsigned int *c) : counter(c) {}
    __device__ int64_t append(bool want) const {
        const unsigned full = 0xFFFFFFFFu;
        const unsigned m = __ballot_sync(full, want);
        if (m == 0) return -1;
        const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
        unsigned int base = 0;
        if (lane == leader) base = atomicAdd(counter, (unsigned int)__popc(m));
        base = __shfl_sync(full, base, leader);
        return (int64_t)base + __popc(m & ((1u << lane) - 1u));
    }
    __device__ void consumed(int64_t) const {}
    __device__ void publish(int64_t, bool) const {}
};

// Lane policy of the cooperative (latency) mode, see metis_eval.cuh `Serial`.
#ifdef METIS_PROFILE_PHASES
__device__ long long g_mark_acc[64];   // [0,32) cycles per phase, [32,64) largest lane skew seen at each mark
#endif
struct WarpLanes {
    static constexpr bool kUniform = true;
#ifdef METIS_PROFILE_PHASES
    mutable long long t_last = 0;
    mutable int cur = 0;
    __device__ void mark(int id) const {
#ifdef METIS_PROBE_SKEW
        {   // do the lanes of the warp reach this point in the same cycle?  (redundant execution relies on it)
            const long long t = clock64();
            const long long t0 = __shfl_sync(0xFFFFFFFFu, t, 0);
            const long long d = t > t0 ? t - t0 : t0 - t;
            if (d) atomicMax((unsigned long long *)&g_mark_acc[32 + (id & 31)], (unsigned long long)d);
        }
#endif
        if ((threadIdx.x & 31) != 0) return;
        const long long now = clock64();
        if (t_last) atomicAdd((unsigned long long *)&g_mark_acc[cur & 31], (unsigned long long)(now - t_last));
        cur = id; t_last = now;
    }
#else
    __device__ void mark(int) const {}
#endif
    __device__ int lane() const { return threadIdx.x & 31; }
    __device__ int width() const { return 32; }
    __device__ void sync() const { __syncwarp(); }
    __device__ bool any(bool p) const { return __any_sync(0xFFFFFFFFu, p); }
    __device__ void argmax_first(double &v, int &i) const {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const double v2 = __shfl_xor_sync(0xFFFFFFFFu, v, d);
            const int i2 = __shfl_xor_sync(0xFFFFFFFFu, i, d);
            if (v2 > v || (v2 == v && i2 < i)) { v = v2; i = i2; }
        }
    }
    __device__ double max_all(double v) const {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const double v2 = __shfl_xor_sync(0xFFFFFFFFu, v, d);
            if (v2 > v) v = v2;
        }
        return v;
    }
};

// Continuous latency mode: once a list is short enough to run one task per warp, the grid stops meeting at
// barriers.  The remaining tasks live in a bounded multi-producer / multi-consumer ring over the slots of one
// task buffer (sequence number per slot: == ticket -> free for that ticket, == ticket + 1 -> published); a
// warp pops a ticket, runs the task, pushes its successor (if the plan continues) and pops again, so nobody
// idles while any chain still has a step to run.  At most one task per chain is alive, so with ring >= number of
// tasks at the switch a producer never finds its slot occupied; `alive` (pushed - retired) reaching zero tells
// the poppers that no ticket will ever be published again.
__device__ __forceinline__ unsigned int ld_relaxed_u32(const unsigned int *p) {     // polling: no L1 invalidation per try
    unsigned int v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u32(unsigned int *p, unsigned int v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Watchdog of the queue's spin loops: a waiter that sees neither a push nor a retirement for 2 s declares the
// scheduler dead (ctl[3] = 1 releases every waiter, the search reports METIS_FATAL_SCHEDULER instead of hanging).
// All waiting below is written warp-uniformly: every lane runs the same loop and takes the same branches, lane 0
// performs the memory operation (a single predicated instruction) and its value is broadcast.  A branchy
// `if (lane == 0) { spin }` lets the compiler keep lane 0 and lanes 1-31 as two separately scheduled groups for
// the rest of the task, and the redundant-execution sections of the latency mode then see
// each other's half-finished updates.
struct QueueWatch {
    unsigned int *ctl;
    unsigned long long *diag;          // counters[16..]: which loop, ticket, slot value, head, tail, alive
    unsigned int polls = 0, last_tail = 0, last_alive = 0;
    unsigned long long since = 0;
    __device__ QueueWatch(unsigned int *c, unsigned long long *d) : ctl(c), diag(d) {}
    // true = give up (uniform: every lane evaluates the same broadcast values)
    __device__ bool expired(int which, unsigned int ticket, unsigned int seen) {
        if ((++polls & 255u) != 0) return false;
        unsigned int tail = 0, alive = 0, fired = 0;
        unsigned long long now = 0;
        if ((threadIdx.x & 31) == 0) {
            tail = *(volatile unsigned int *)&ctl[1]; alive = *(volatile unsigned int *)&ctl[2];
            fired = *(volatile unsigned int *)&ctl[3]; now = global_ns();
        }
        tail = __shfl_sync(0xFFFFFFFFu, tail, 0); alive = __shfl_sync(0xFFFFFFFFu, alive, 0);
        fired = __shfl_sync(0xFFFFFFFFu, fired, 0); now = __shfl_sync(0xFFFFFFFFu, now, 0);
        if (fired) return true;
        if (since == 0 || tail != last_tail || alive != last_alive) { since = now; last_tail = tail; last_alive = alive; return false; }
        if (now - since < 2000000000ULL) return false;
        if ((threadIdx.x & 31) == 0 && atomicExch(&ctl[3], 1u) == 0u) {
            diag[0] = (unsigned long long)which; diag[1] = ticket; diag[2] = seen;
            diag[3] = *(volatile unsigned int *)&ctl[0]; diag[4] = tail; diag[5] = alive;
        }
        return true;
    }
};
struct QueueWarp {
    unsigned int *seq;         // [ring]
    unsigned int *ctl;         // [0] head (pop tickets) [1] tail (push tickets) [2] alive [3] watchdog fired
    unsigned long long *diag;
    unsigned int ring;
    unsigned int taken;        // ticket of the task being run
    unsigned int pushed;       // ticket of the successor being written
    __device__ int64_t append(bool want) {
        if (!want) return -1;                                // uniform
        const bool lead = (threadIdx.x & 31) == 0;
        unsigned int t = 0;
        if (lead) t = atomicAdd(&ctl[1], 1u);
        if (lead) atomicAdd(&ctl[2], 1u);                    // alive before the parent retires
        t = __shfl_sync(0xFFFFFFFFu, t, 0);
        QueueWatch watch(ctl, diag);
        for (;;) {                                           // slot free for this ticket? (never waits when ring >= tasks alive)
            unsigned int v = 0;
            if (lead) v = ld_relaxed_u32(&seq[t % ring]);
            v = __shfl_sync(0xFFFFFFFFu, v, 0);
            if (v == t || watch.expired(1, t, v)) break;
            __nanosleep(64);
        }
        __threadfence();                                     // every lane: acquire side of the slot hand-over
        pushed = t;
        return (int64_t)(t % ring);
    }
    __device__ void publish(int64_t slot, bool want) const {
        if (!want) return;
        __threadfence();                                     // every lane's stores to the slot are visible ...
        __syncwarp();
        if ((threadIdx.x & 31) == 0) st_relaxed_u32(&seq[slot], pushed + 1u);   // ... before the flag
    }
    __device__ void consumed(int64_t slot) const {
        __threadfence();                                     // every lane's loads from the slot are complete
        __syncwarp();
        if ((threadIdx.x & 31) == 0) st_relaxed_u32(&seq[slot], taken + ring);
    }
};

struct RoundBuffers {
    TaskBuffers buf[2];
    unsigned int *counts;      // [3] rotating task counters
    unsigned int *seq;         // [wave] slot sequence numbers of the continuous latency mode
    long long wave;            // plans admitted per wave (= capacity of the task lists)
    long long coop_below;      // rounds with fewer pending tasks run one task per warp (latency mode)
    unsigned long long *trace; // profiling builds: (tasks, globaltimer ns) per round, 512 entries
};

template <int MAXS, int MAXL>
__global__ void __launch_bounds__(kThreads, (MAXS <= 64 ? METIS_MIN_BLOCKS : (METIS_MIN_BLOCKS * 2 + 2) / 3))
het_search_kernel(const __grid_constant__ MetisProblem p, const __grid_constant__ MetisPlanSpace sp,
                  const MetisShard sh, const __grid_constant__ BlobLayout lay, const uint8_t *__restrict__ blob,
                  const int use_smem, const unsigned int scratch_off, const long long slots,
                  const __grid_constant__ DeviceOut out, const __grid_constant__ RoundBuffers rb) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t mbar;
    __shared__ double s_cost[kThreads / 32];
    __shared__ uint32_t s_ord[kThreads / 32], s_step[kThreads / 32], s_meta[kThreads / 32];
    cg::grid_group grid = cg::this_grid();

    const uint8_t *base = blob;
    if (use_smem) {
        stage_blob_tma(smem, blob, lay.total, &mbar);
        base = smem;
    }
    const Tables T = make_tables(p, lay, base);

    DeviceSink sink(out);
    {
        Scratch<MAXS, MAXL> w;
        Scratch<MAXS, MAXL> *wsh = reinterpret_cast<Scratch<MAXS, MAXL> *>(smem + scratch_off) + (threadIdx.x >> 5);
#ifndef METIS_NO_OPAQUE
        asm volatile("" : "+l"(wsh));        // opaque: keep the pointer in a register instead of re-deriving it at every use
#endif
        const int lane = threadIdx.x & 31;
        const long long gwarp = (long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
        const long long nwarps = (long long)gridDim.x * (kThreads / 32);
        unsigned int round = 0;                              // global round counter (rotates the 3 counters)
        for (long long wave0 = 0; wave0 < slots; wave0 += rb.wave) {
            const long long wave1 = wave0 + rb.wave < slots ? wave0 + rb.wave : slots;
            // ---- admission: every plan of the wave -> first strategy that can be valid ----------
            sink.phase(0);
            {
                DeviceWarp warp(&rb.counts[(round + 1) % 3]);
                if (blockIdx.x == 0 && threadIdx.x == 0) { rb.counts[(round + 2) % 3] = 0; rb.counts[4 + (round + 2) % 3] = 0; rb.counts[4 + (round + 1) % 3] = 0; }
                for (long long b0 = wave0 + gwarp * 32; b0 < wave1; b0 += nwarps * 32) {
                    const long long i = b0 + lane;
                    PlanDesc pd;
                    bool has = false;
                    // tiles are multiples of 32, so a warp's 32 plans are consecutive ordinals
                    const long long first = ((b0 / sh.tile) * sh.world + sh.rank) * sh.tile + (b0 % sh.tile);
                    int hint = 0;
                    if (lane == 0 && first < sp.num_plans) hint = find_block(sp, first);
                    hint = __shfl_sync(0xFFFFFFFFu, hint, 0);
                    if (i < wave1) {
                        const long long ordinal = first + lane;
                        has = decode_plan(sp, ordinal, pd, hint);
                    }
                    begin_task<MAXS, MAXL>(T, w, sink, warp, rb.buf[(round + 1) & 1], has, pd);
                }
            }
            grid.sync();
            ++round;
            // ---- order the first task list by stage count (counting sort): the throughput mode runs 32
            //      tasks per warp in lockstep and is ~20 % faster when they share loop trip counts ----------
            {
                const unsigned int n0 = *(volatile unsigned int *)&rb.counts[round % 3];
                if ((long long)n0 >= rb.coop_below) {
                    __shared__ unsigned int s_base[METIS_MAX_STAGES + 1];
                    unsigned int *hist = rb.counts + 16, *cursor = rb.counts + 16 + 160;
                    const TaskBuffers &src = rb.buf[round & 1], &dst = rb.buf[(round + 1) & 1];
                    // (both passes aggregate per warp: neighbours in the admission order mostly share a count)
                    const long long span = (long long)gridDim.x * kThreads;
                    const long long upto = (((long long)n0 + 31) / 32) * 32;
                    for (long long pos = (long long)blockIdx.x * kThreads + threadIdx.x; pos < upto; pos += span) {
                        const int k = pos < (long long)n0 ? (int)((src.geo[pos] >> 32) & 0xFF) : -1;
                        const unsigned int peers = __match_any_sync(0xFFFFFFFFu, k);
                        if (k >= 0 && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&hist[k], (unsigned int)__popc(peers));
                    }
                    grid.sync();
                    if (threadIdx.x == 0) {
                        unsigned int acc = 0;
                        for (int k = METIS_MAX_STAGES - 1; k >= 0; --k) { s_base[k] = acc; acc += *(volatile unsigned int *)&hist[k]; }   // longest first
                    }
                    __syncthreads();
                    for (long long pos = (long long)blockIdx.x * kThreads + threadIdx.x; pos < upto; pos += span) {
                        const bool live = pos < (long long)n0;
                        const uint64_t g = live ? src.geo[pos] : 0;
                        const int k = live ? (int)((g >> 32) & 0xFF) : -1;
                        const unsigned int peers = __match_any_sync(0xFFFFFFFFu, k);
                        const int leader = __ffs(peers) - 1, me = threadIdx.x & 31;
                        unsigned int at = 0;
                        if (live && me == leader) at = atomicAdd(&cursor[k], (unsigned int)__popc(peers));
                        at = __shfl_sync(0xFFFFFFFFu, at, leader);
                        if (!live) continue;
                        const long long to = (long long)s_base[k] + at + __popc(peers & ((1u << me) - 1u));
                        dst.hdr[to] = src.hdr[pos];
                        dst.geo[to] = g;
                        for (int st = 0; st <= k; ++st) dst.tpc[(long long)st * dst.cap + to] = src.tpc[(long long)st * src.cap + pos];
                    }
                    if (blockIdx.x == 0 && threadIdx.x == 0) { rb.counts[(round + 1) % 3] = n0; rb.counts[(round + 2) % 3] = 0; rb.counts[4 + (round + 2) % 3] = 0; }
                    grid.sync();
                    if (blockIdx.x == 0 && threadIdx.x < METIS_MAX_STAGES) { hist[threadIdx.x] = 0; cursor[threadIdx.x] = 0; }
                    ++round;
                }
            }
            // ---- rounds: one partition attempt per pending plan --------------------------------
            for (;;) {
                const unsigned int n = *(volatile unsigned int *)&rb.counts[round % 3];
#ifdef METIS_PROFILE_PHASES
                if (blockIdx.x == 0 && threadIdx.x == 0 && round < 512) {
                    unsigned long long t;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                    rb.trace[2 * round] = n;
                    rb.trace[2 * round + 1] = t;
                }
#endif
                if (n == 0) break;
                DeviceWarp warp(&rb.counts[(round + 1) % 3]);
                if (blockIdx.x == 0 && threadIdx.x == 0) { rb.counts[(round + 2) % 3] = 0; rb.counts[4 + (round + 2) % 3] = 0; }
                const TaskBuffers &in = rb.buf[round & 1], &nxt = rb.buf[(round + 1) & 1];
                if ((long long)n >= rb.coop_below) {
                    // throughput mode: one task per lane, 32 tasks per warp in lockstep
                    // (batches are handed out dynamically: the list is ordered longest first)
                    for (;;) {
                        unsigned int fetched = 0;
                        if (lane == 0) fetched = atomicAdd(&rb.counts[4 + round % 3], 1u);
                        const long long b0 = 32LL * __shfl_sync(0xFFFFFFFFu, fetched, 0);
                        if (b0 >= (long long)n) break;
                        const long long pos = b0 + lane;
                        PlanDesc pd;
                        bool has = false;
                        if (pos < (long long)n) { decode_task(sp, list_load(&in.hdr[pos]), list_load(&in.geo[pos]), pd); has = true; }
                        run_task<MAXS, MAXL>(T, w, Serial(), sink, warp, in, nxt, has, pos, pd);
                    }
                } else {
                    // latency mode: one task per warp on shared-memory scratch, no more barriers (QueueWarp)
                    unsigned int *ctl = rb.counts + 8;
                    const long long room = in.cap < (long long)n + 64 ? in.cap : (long long)n + 64;
                    const unsigned int ring = (unsigned int)room;
                    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < room; i += (long long)gridDim.x * kThreads)
                        rb.seq[i] = i < (long long)n ? (unsigned int)i + 1u : (unsigned int)i;
                    if (blockIdx.x == 0 && threadIdx.x == 0) { ctl[0] = 0; ctl[1] = n; ctl[2] = n; ctl[3] = 0; }
                    grid.sync();
                    QueueWarp qwarp{rb.seq, ctl, out.counters + 16, ring, 0u, 0u};
                    WarpLanes lanes_coop;
                    sink.leader = lane == 0;
                    for (;;) {
                        unsigned int h = 0;
                        if (lane == 0) h = atomicAdd(&ctl[0], 1u);
                        h = __shfl_sync(0xFFFFFFFFu, h, 0);
                        // every chain is finite (tp only grows, <= 3 attempts per strategy): far more pops than that = runaway
                        if (h > 4096u * METIS_MAX_STAGES + 64u * n) {
                            if (lane == 0 && atomicExch(&ctl[3], 1u) == 0u) { out.counters[16] = 3; out.counters[17] = h; out.counters[20] = ctl[1]; out.counters[21] = ctl[2]; }
                            break;
                        }
                        int ready = 0;
                        unsigned int nap = 128;                   // back off to 2 us: idle warps share issue slots with working ones
                        QueueWatch watch(ctl, out.counters + 16);
                        for (;;) {                                // warp-uniform wait: lane 0 loads, everyone decides
                            unsigned int v = 0, alive = 0;
                            if (lane == 0) v = ld_relaxed_u32(&rb.seq[h % ring]);
                            if (lane == 0) alive = ld_relaxed_u32(&ctl[2]);
                            v = __shfl_sync(0xFFFFFFFFu, v, 0);
                            alive = __shfl_sync(0xFFFFFFFFu, alive, 0);
                            if (v == h + 1u) { ready = 1; break; }
                            if (alive == 0u) break;               // nothing alive: ticket h will never exist
                            if (watch.expired(2, h, v)) break;
                            __nanosleep(nap);
                            if (nap < 2048) nap <<= 1;
                        }
                        if (!ready) break;
                        __threadfence();                          // acquire: the slot's words were written before its flag
                        qwarp.taken = h;
                        const long long pos = (long long)(h % ring);
                        PlanDesc pd;
                        decode_task(sp, list_load(&in.hdr[pos]), list_load(&in.geo[pos]), pd);
                        run_task<MAXS, MAXL>(T, *wsh, lanes_coop, sink, qwarp, in, in, true, pos, pd);
                        __threadfence();                          // the successor (if any) is published before the parent retires
                        if (lane == 0) atomicSub(&ctl[2], 1u);
                    }
                    sink.leader = true;
                    grid.sync();
                    if (blockIdx.x == 0 && threadIdx.x == 0) {
                        rb.counts[round % 3] = 0;                 // this wave is finished
                        if (ctl[3]) atomicMin(&out.counters[4], (unsigned long long)METIS_FATAL_SCHEDULER << 24);
                    }
                    grid.sync();
                    break;
                }
                grid.sync();
                ++round;
            }
            // counts[round % 3] is 0 here and becomes this wave's successor "previous" counter; the
            // admission of the next wave appends to counts[(round+1) % 3], which was zeroed one round ago
        }
    }

#ifdef METIS_PROFILE_PHASES
    if ((threadIdx.x & 31) == 0)
        for (int k = 0; k < 5; ++k) atomicAdd(&out.counters[8 + k], (unsigned long long)sink.acc[k]);
#endif
    // counters: warp reduce, one atomic per warp
    const unsigned full = 0xFFFFFFFFu;
    const unsigned np = __reduce_add_sync(full, sink.n_part);
    const unsigned nr = __reduce_add_sync(full, sink.n_run);
    const unsigned nk = __reduce_add_sync(full, sink.n_key);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        if (np) atomicAdd(&out.counters[1], (unsigned long long)np);
        if (nr) atomicAdd(&out.counters[2], (unsigned long long)nr);
        if (nk) atomicAdd(&out.counters[3], (unsigned long long)nk);
    }
    // argmin (cost, ordinal, step): __shfl_sync butterfly inside the warp, then across warps
    double c = sink.best_cost;
    uint32_t o = sink.best_ord, st = sink.best_step, mt = sink.best_meta;
    for (int d = 16; d > 0; d >>= 1) {
        const double c2 = __shfl_xor_sync(full, c, d);
        const uint32_t o2 = __shfl_xor_sync(full, o, d), s2 = __shfl_xor_sync(full, st, d),
                       m2 = __shfl_xor_sync(full, mt, d);
        if (rec_less(c2, o2, s2, c, o, st)) { c = c2; o = o2; st = s2; mt = m2; }
    }
    if (lane == 0) { s_cost[warp] = c; s_ord[warp] = o; s_step[warp] = st; s_meta[warp] = mt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int wi = 1; wi < kThreads / 32; ++wi)
            if (rec_less(s_cost[wi], s_ord[wi], s_step[wi], c, o, st)) { c = s_cost[wi]; o = s_ord[wi]; st = s_step[wi]; mt = s_meta[wi]; }
        MetisRecord r;
        r.cost = c; r.ordinal = o; r.step = (uint16_t)st; r.num_repartition = (uint8_t)(mt >> 8); r.num_stage = (uint8_t)mt;
        out.block_best[blockIdx.x] = r;
    }
}

__global__ void het_finalize_kernel(const MetisRecord *block_best, int nblocks, const unsigned long long *counters,
                                    MetisSearchSummary *summary) {
    __shared__ double s_cost[32];
    __shared__ uint32_t s_ord[32], s_step[32], s_meta[32];
    double c = INFINITY;
    uint32_t o = 0xFFFFFFFFu, st = 0xFFFFu, mt = 0;
    for (int i = threadIdx.x; i < nblocks; i += blockDim.x) {
        const MetisRecord r = block_best[i];
        if (rec_less(r.cost, r.ordinal, r.step, c, o, st)) {
            c = r.cost; o = r.ordinal; st = r.step; mt = ((uint32_t)r.num_repartition << 8) | r.num_stage;
        }
    }
    const unsigned full = 0xFFFFFFFFu;
    for (int d = 16; d > 0; d >>= 1) {
        const double c2 = __shfl_xor_sync(full, c, d);
        const uint32_t o2 = __shfl_xor_sync(full, o, d), s2 = __shfl_xor_sync(full, st, d),
                       m2 = __shfl_xor_sync(full, mt, d);
        if (rec_less(c2, o2, s2, c, o, st)) { c = c2; o = o2; st = s2; mt = m2; }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { s_cost[warp] = c; s_ord[warp] = o; s_step[warp] = st; s_meta[warp] = mt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nw = (blockDim.x + 31) / 32;
        for (int wi = 1; wi < nw; ++wi)
            if (rec_less(s_cost[wi], s_ord[wi], s_step[wi], c, o, st)) { c = s_cost[wi]; o = s_ord[wi]; st = s_step[wi]; mt = s_meta[wi]; }
        MetisSearchSummary s;
        memset(&s, 0, sizeof(s));
        s.num_records = counters[0];
        s.num_partition_calls = counters[1];
        s.num_balancer_runs = counters[2];
        s.num_keyerror = counters[3];
        const unsigned long long fk = counters[4];
        if (fk == 0xFFFFFFFFFFFFFFFFULL) { s.fatal_ordinal = 0xFFFFFFFFFFFFFFFFULL; s.fatal_code = 0; s.fatal_aux = 0; }
        else { s.fatal_ordinal = fk >> 32; s.fatal_code = (uint32_t)((fk >> 24) & 0xFF); s.fatal_aux = (uint32_t)(fk & 0xFFFFFF); }
        s.best.cost = c; s.best.ordinal = o; s.best.step = (uint16_t)st;
        s.best.num_repartition = (uint8_t)(mt >> 8); s.best.num_stage = (uint8_t)mt;
        s.reserved[0] = counters[8]; s.reserved[1] = counters[9];        // profiling builds: phase clocks
        s.reserved[2] = counters[10]; s.reserved[3] = counters[11]; s.reserved[4] = counters[12];
        if (s.fatal_code == METIS_FATAL_SCHEDULER)                       // watchdog diagnostics (QueueWatch)
            for (int k = 0; k < 6; ++k) s.reserved[k] = counters[16 + k];
        *summary = s;
    }
}

// replays single candidates: writes dp code, tp code, partition for (ordinal, step)
struct DetailSink {
    uint8_t *dst;
    int want_step;
    __device__ void phase(int) {}
    __device__ void partition_call() {}
    __device__ void balancer_run() {}
    __device__ void keyerror() {}
    __device__ void fatal(uint32_t, int, uint32_t) {}
    __device__ void emit(const PlanDesc &pd, int step, int, double, const uint8_t *tpc, const uint16_t *part) {
        if (step != want_step) return;
        for (int s = 0; s < pd.S; ++s) { dst[s] = (uint8_t)(pd.row[s] - tpc[s]); dst[pd.S + s] = tpc[s]; }
        for (int s = 0; s <= pd.S; ++s) dst[2 * pd.S + s] = (uint8_t)part[s];
    }
};

template <int MAXS, int MAXL>
__global__ void __launch_bounds__(kThreads)
het_detail_kernel(const __grid_constant__ MetisProblem p, const __grid_constant__ MetisPlanSpace sp,
                  const __grid_constant__ BlobLayout lay, const uint8_t *__restrict__ blob,
                  const MetisRecord *__restrict__ picks, long long n, uint8_t *detail, int stride) {
    const Tables T = make_tables(p, lay, blob);
    const long long i = (long long)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    PlanDesc pd;
    if (!decode_plan(sp, picks[i].ordinal, pd)) return;
    DetailSink sink{detail + (size_t)i * stride, (int)picks[i].step};
    Scratch<MAXS, MAXL> w;
    PlanEvaluator<MAXS, MAXL> ev(T, w);
    ev.run(pd, sink, (int)picks[i].step);
}

__global__ void __launch_bounds__(kThreads)
homo_cost_kernel(const __grid_constant__ MetisProblem p, const __grid_constant__ BlobLayout lay,
                 const uint8_t *__restrict__ blob, int type_id, const int32_t *__restrict__ plans, long long n,
                 double *cost, int32_t *status) {
    const Tables T = make_tables(p, lay, blob);
    const long long i = (long long)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    const int32_t *q = plans + i * 5;
    double c = 0.0;
    int oom = 0;
    const int rc = homo_cost(T, type_id, q[0], q[1], q[2], q[3], q[4], c, oom);
    cost[i] = rc ? NAN : c;
    status[i] = rc ? 1 : (oom ? 2 : 0);
}

template <int MAXS, int MAXL>
__global__ void __launch_bounds__(kThreads)
layer_balance_kernel(const double *__restrict__ capa, const int32_t *__restrict__ num_stage, long long n, int stride,
                     const double *__restrict__ lc, const double *__restrict__ dlay, int norm_len, int num_layers,
                     uint16_t *partition) {
    const long long i = (long long)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    Tables T;
    memset(&T, 0, sizeof(T));
    T.p.num_layers = num_layers;
    T.p.norm_len = norm_len;
    T.norm_lc = lc;
    T.dlay = dlay;
    Scratch<MAXS, MAXL> w;
    const int S = num_stage[i];
    uint16_t *out = partition + i * (stride + 1);
    if (S < 1 || S > MAXS || num_layers > MAXL) { out[0] = 0xFFFF; return; }
    for (int s = 0; s < S; ++s) w.perf[s] = capa[i * stride + s];
    const int rc = balance_run<MAXS, MAXL>(T, S, w, Serial());
    if (rc) { out[0] = 0xFFFF; return; }
    for (int s = 0; s <= S; ++s) out[s] = w.part[s];
}

__global__ void divide_by_seven_kernel(const double *lc, int n, double *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = lc[i] / 7.0;
}

static int check_problem(const MetisProblem *p) {
    if (!p) return arg_fail("problem is NULL");
    if (p->num_types < 1 || p->num_types > METIS_MAX_TYPES) return arg_fail("num_types out of range");
    if (p->num_layers < 1 || p->num_layers > METIS_MAX_LAYERS) return arg_fail("num_layers out of range (METIS_MAX_LAYERS)");
    if (p->lpad < p->num_layers) return arg_fail("lpad < num_layers");
    if (p->num_keys < 1 || p->num_tp < 1 || p->num_bs < 1 || p->norm_len < 1) return arg_fail("empty profile tables");
    if (p->devices_per_node < 1 || p->total_devices < 1) return arg_fail("empty cluster");
    return METIS_OK;
}

}  // namespace metis

using namespace metis;

extern "C" {

const char *metis_last_error(void) { return g_err; }
int metis_abi_version(void) { return METIS_ABI_VERSION; }
#ifdef METIS_PROFILE_PHASES
int metis_debug_marks(long long *out32, int reset) {
    long long zero[64] = {0};
    if (out32) cudaMemcpyFromSymbol(out32, g_mark_acc, sizeof(zero));      // caller provides 64 entries
    if (reset) cudaMemcpyToSymbol(g_mark_acc, zero, sizeof(zero));
    return 0;
}
#endif
void metis_set_profile_events(void *before_kernel, void *after_kernel) {
    g_ev_before = static_cast<cudaEvent_t>(before_kernel);
    g_ev_after = static_cast<cudaEvent_t>(after_kernel);
}

static int64_t shard_plan_slots(int64_t num_plans, const MetisShard *sh) {
    const int64_t tile = sh->tile, world = sh->world;
    const int64_t rounds = (num_plans + tile * world - 1) / (tile * world);
    return rounds * tile;
}

constexpr int64_t kFixedWs = 16384;                // summary + counters + round counters + round trace
constexpr int64_t kMaxBlocks = 4096;               // per-block best records
// Bytes of task-list storage before the plan space is cut into waves.  Every wave pays one latency-bound
// tail of near-empty rounds, so the default spends HBM (32 of the B200's 180 GB) to keep the spaces of
// BASELINE.json in one wave; METIS_TASK_MIB overrides it (the tests use it to force many waves).
static int64_t round_budget() {
    const char *e = getenv("METIS_TASK_MIB");
    const long long m = e ? atoll(e) : 0;
    return (int64_t)((m >= 1 && m <= 160 * 1024 ? m : 32 * 1024) << 20);
}

static int64_t task_slot_bytes(int max_stage) { return 16 + (int64_t)max_stage + 8 * (int64_t)max_stage; }

static int64_t wave_size(int64_t slots, int max_stage) {
    int64_t cap = round_budget() / (2 * task_slot_bytes(max_stage));
    cap &= ~(int64_t)127;
    if (cap < 65536) cap = 65536;
    if (cap > slots) cap = (slots + 127) & ~(int64_t)127;
    return cap < 128 ? 128 : cap;
}

int64_t metis_het_workspace_bytes(const MetisProblem *problem, int64_t num_plans, int32_t max_stage) {
    if (check_problem(problem)) return METIS_E_ARG;
    if (max_stage < 1) max_stage = 1;
    if (max_stage > METIS_MAX_STAGES) max_stage = METIS_MAX_STAGES;
    const BlobLayout lay = make_layout(*problem);
    const int64_t cap = wave_size(num_plans, max_stage);
    return 256 + kFixedWs + (int64_t)align16(lay.total) + kMaxBlocks * (int64_t)sizeof(MetisRecord) +
           2 * cap * task_slot_bytes(max_stage) + cap * 4 + 1024;
}

struct Workspace {
    MetisSearchSummary *summary;
    unsigned long long *counters;
    unsigned int *round_counts;
    uint8_t *blob;
    MetisRecord *block_best;
    uint8_t *tasks;
};

static Workspace carve(void *ws, const BlobLayout &lay) {
    uint8_t *b = static_cast<uint8_t *>(ws);
    uintptr_t a = (reinterpret_cast<uintptr_t>(b) + 127) & ~(uintptr_t)127;
    b = reinterpret_cast<uint8_t *>(a);
    Workspace w;
    w.summary = reinterpret_cast<MetisSearchSummary *>(b);
    w.counters = reinterpret_cast<unsigned long long *>(b + 1024);
    w.round_counts = reinterpret_cast<unsigned int *>(b + 2048);
    w.blob = b + kFixedWs;
    w.block_best = reinterpret_cast<MetisRecord *>(b + kFixedWs + align16(lay.total));
    w.tasks = reinterpret_cast<uint8_t *>(w.block_best + kMaxBlocks);
    return w;
}

static TaskBuffers carve_tasks(uint8_t *&p, int64_t cap, int max_stage) {
    TaskBuffers t;
    t.cap = cap;
    t.hdr = reinterpret_cast<uint64_t *>(p);   p += cap * 8;
    t.geo = reinterpret_cast<uint64_t *>(p);   p += cap * 8;
    t.perf = reinterpret_cast<double *>(p);    p += cap * 8 * (int64_t)max_stage;
    t.tpc = p;                                 p += cap * (int64_t)max_stage;
    return t;
}

int metis_het_search(const MetisProblem *problem, const MetisPlanSpace *space, const MetisShard *shard,
                     MetisRecord *records, int64_t capacity, uint8_t *detail, int32_t detail_stride,
                     void *workspace, int64_t workspace_bytes, MetisSearchSummary *summary, void *stream_) {
    int rc = check_problem(problem);
    if (rc) return rc;
    if (!space || !shard || !workspace || !summary) return arg_fail("NULL argument");
    if (shard->world < 1 || shard->rank < 0 || shard->rank >= shard->world || shard->tile < 32 || shard->tile % 32)
        return arg_fail("bad shard (tile must be a positive multiple of 32)");
    if (space->num_plans > 0xFFFFFFF0LL) return arg_fail("more than 2^32 plans");
    if (space->max_stage < 1 || space->max_stage > METIS_MAX_STAGES) return arg_fail("max_stage out of range (METIS_MAX_STAGES)");
    if (space->num_div < 1 || space->num_div > 256) return arg_fail("more than 256 divisors of gbs");
    if (detail && detail_stride < 3 * METIS_MAX_STAGES + 1) return arg_fail("detail_stride too small");
    if (capacity < 0 || (capacity > 0 && !records)) return arg_fail("records/capacity mismatch");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const BlobLayout lay = make_layout(*problem);
    const int64_t slots = shard_plan_slots(space->num_plans, shard);
    const int64_t need = metis_het_workspace_bytes(problem, slots, space->max_stage);
    if (workspace_bytes < need) { snprintf(g_err, sizeof(g_err), "workspace too small: need %lld", (long long)need); return METIS_E_CAPACITY; }
    const Workspace ws = carve(workspace, lay);

    cudaError_t e;
    e = cudaMemsetAsync(ws.counters, 0, kFixedWs - 1024, stream);   // counters, round counters, trace
    if (e != cudaSuccess) return cuda_fail(e, "memset counters");
    e = cudaMemsetAsync(ws.counters + 4, 0xFF, sizeof(unsigned long long), stream);
    if (e != cudaSuccess) return cuda_fail(e, "memset fatal key");
    pack_tables_kernel<<<8, 256, 0, stream>>>(*problem, lay, ws.blob);
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "pack_tables_kernel");

    static const uint32_t blob_max = []() {
        const char *env = getenv("METIS_SMEM_BLOB_MAX");        // tuning knob: tables larger than this stay in global memory
        return env ? (uint32_t)atoi(env) : (uint32_t)kSmemBlobMax;
    }();
    int use_smem = lay.total <= blob_max;
    unsigned int scratch_off = use_smem ? ((lay.total + 127u) & ~127u) : 0u;
    // two instantiations: the small one (S <= 64, L <= 128) halves the per-warp scratch -> more resident warps
    const bool small = space->max_stage <= 64 && problem->num_layers <= 128;
    size_t dyn = scratch_off + (kThreads / 32) * (small ? sizeof(Scratch<64, 128>) : sizeof(Scratch<kMaxS, kMaxL>));
    auto kern = small ? het_search_kernel<64, 128> : het_search_kernel<kMaxS, kMaxL>;
    if (dyn > 48 * 1024) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
    }
    DeviceOut out;
    out.records = records; out.capacity = capacity; out.detail = detail; out.detail_stride = detail_stride;
    out.counters = ws.counters; out.block_best = ws.block_best;
    RoundBuffers rb;
    uint8_t *tp = ws.tasks;
    rb.wave = wave_size(slots, space->max_stage);
    rb.buf[0] = carve_tasks(tp, rb.wave, space->max_stage);
    rb.buf[1] = carve_tasks(tp, rb.wave, space->max_stage);
    rb.seq = reinterpret_cast<unsigned int *>(tp);
    rb.counts = ws.round_counts;
    rb.trace = reinterpret_cast<unsigned long long *>(reinterpret_cast<uint8_t *>(ws.summary) + 4096);
    // cooperative persistent grid: every block is resident, rounds are separated by grid.sync()
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, dyn);
    if (e != cudaSuccess || per_sm < 1 || sms < 1) return cuda_fail(e, "occupancy query");
    int64_t grid = (int64_t)sms * per_sm;
    const int64_t useful = (slots + kThreads - 1) / kThreads;
    if (grid > useful) grid = useful;
    if (grid > kMaxBlocks) grid = kMaxBlocks;
    if (slots > 0) {
        MetisProblem p_arg = *problem;
        MetisPlanSpace s_arg = *space;
        MetisShard sh_arg = *shard;
        BlobLayout lay_arg = lay;
        const uint8_t *blob_arg = ws.blob;
        long long slots_arg = slots;
        rb.coop_below = (long long)(shard->reserved > 0 ? shard->reserved : 12) * grid * (kThreads / 32);
        void *args[] = {&p_arg, &s_arg, &sh_arg, &lay_arg, &blob_arg, &use_smem, &scratch_off, &slots_arg, &out, &rb};
        if (g_ev_before) cudaEventRecord(g_ev_before, stream);
        e = cudaLaunchCooperativeKernel((const void *)kern, dim3((unsigned)grid), dim3(kThreads), args, dyn, stream);
        if (g_ev_after) cudaEventRecord(g_ev_after, stream);
        g_ev_before = g_ev_after = nullptr;
        if (e != cudaSuccess) return cuda_fail(e, "het_search_kernel (cooperative launch)");
    }
    het_finalize_kernel<<<1, 256, 0, stream>>>(ws.block_best, (int)(slots > 0 ? grid : 0), ws.counters, ws.summary);
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "het_finalize_kernel");
    e = cudaMemcpyAsync(summary, ws.summary, sizeof(MetisSearchSummary), cudaMemcpyDeviceToHost, stream);
    if (e != cudaSuccess) return cuda_fail(e, "copy summary");
    return METIS_OK;
}

int metis_het_detail(const MetisProblem *problem, const MetisPlanSpace *space, const MetisRecord *picks, int64_t n,
                     uint8_t *detail, int32_t detail_stride, void *workspace, int64_t workspace_bytes, void *stream_) {
    int rc = check_problem(problem);
    if (rc) return rc;
    if (!space || !picks || !detail || !workspace) return arg_fail("NULL argument");
    if (detail_stride < 3 * METIS_MAX_STAGES + 1) return arg_fail("detail_stride too small");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const BlobLayout lay = make_layout(*problem);
    if (workspace_bytes < 256 + kFixedWs + (int64_t)align16(lay.total)) return METIS_E_CAPACITY;
    const Workspace ws = carve(workspace, lay);
    pack_tables_kernel<<<8, 256, 0, stream>>>(*problem, lay, ws.blob);
    if (n > 0) {
        const unsigned nb = (unsigned)((n + kThreads - 1) / kThreads);
        het_detail_kernel<kMaxS, kMaxL><<<nb, kThreads, 0, stream>>>(*problem, *space, lay, ws.blob, picks, n, detail,
                                                                     detail_stride);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "het_detail_kernel");
    return METIS_OK;
}

int metis_homo_cost(const MetisProblem *problem, int32_t type_id, const int32_t *plans, int64_t n, double *cost,
                    int32_t *status, void *workspace, int64_t workspace_bytes, void *stream_) {
    int rc = check_problem(problem);
    if (rc) return rc;
    if (!plans || !cost || !status || !workspace) return arg_fail("NULL argument");
    if (type_id < 0 || type_id >= problem->num_types) return arg_fail("type_id out of range");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const BlobLayout lay = make_layout(*problem);
    if (workspace_bytes < 256 + kFixedWs + (int64_t)align16(lay.total)) return METIS_E_CAPACITY;
    const Workspace ws = carve(workspace, lay);
    pack_tables_kernel<<<8, 256, 0, stream>>>(*problem, lay, ws.blob);
    if (n > 0) {
        const unsigned nb = (unsigned)((n + kThreads - 1) / kThreads);
        homo_cost_kernel<<<nb, kThreads, 0, stream>>>(*problem, lay, ws.blob, type_id, plans, n, cost, status);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "homo_cost_kernel");
    return METIS_OK;
}

int metis_layer_balance(const double *capa, const int32_t *num_stage, int64_t n, int32_t stride, const double *lc,
                        int32_t norm_len, int32_t num_layers, uint16_t *partition, void *workspace,
                        int64_t workspace_bytes, void *stream_) {
    if (!capa || !num_stage || !lc || !partition || !workspace) return arg_fail("NULL argument");
    if (stride < 1 || stride > METIS_MAX_STAGES) return arg_fail("stride out of range");
    if (num_layers < 1 || num_layers > METIS_MAX_LAYERS || norm_len < 1) return arg_fail("layers out of range");
    if (workspace_bytes < (int64_t)norm_len * 8 + 256) return METIS_E_CAPACITY;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    double *dlay = reinterpret_cast<double *>((reinterpret_cast<uintptr_t>(workspace) + 127) & ~(uintptr_t)127);
    divide_by_seven_kernel<<<(norm_len + 127) / 128, 128, 0, stream>>>(lc, norm_len, dlay);
    if (n > 0) {
        const unsigned nb = (unsigned)((n + kThreads - 1) / kThreads);
        layer_balance_kernel<kMaxS, kMaxL><<<nb, kThreads, 0, stream>>>(capa, num_stage, n, stride, lc, dlay, norm_len,
                                                                        num_layers, partition);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "layer_balance_kernel");
    return METIS_OK;
}

}  // extern "C"
