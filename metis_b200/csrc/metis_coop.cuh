// metis_coop.cuh - one warp evaluates one inter-stage plan's whole strategy chain.
//
// The chain of a plan (search_space/plan.py:192-268) is sequential: strategy k+1 is chosen from the memory state
// that the partition of strategy k produced, and a partition attempt replays LayerComputeBalancer.run
// (model/load_balancer.py:197-207) whose passes are compare-and-subtract chains in fp64.  What a warp can do in
// parallel is everything that is independent per stage or per layer: profile look-ups, range sums, the 7-way
// vote, arg-max / arg-min reductions, prefix sums.  The execution model is explicit and does not depend on how
// the hardware schedules the lanes of a warp:
//
//   PAR  sections  - a loop over stages / layers strided over the lanes (x.lane(), x.width()); iteration i only
//                    writes scratch entries with index i, and reads nothing that another iteration of the same
//                    section writes;
//   SEQ  sections  - executed by the leader lane alone (x.leader()); results leave the section through the
//                    scratch (w.*), never through registers;
//   x.sync()       - __syncwarp() between any two sections that communicate through the scratch;
//   reductions     - register values are combined with shuffles (x.argmax_first, x.argmin_first, x.max_all,
//                    x.any, x.incl_scan), whose result is identical in every lane.
//
// Control flow outside SEQ sections is uniform (every lane takes the same branches because the conditions are
// reduction results or values read from the scratch after a sync).
//
// The per-stage helpers of PlanEvaluator (metis_eval.cuh: mixed-type stages, bandwidth selection, memory
// capacity) are reused unchanged: inside a PAR section the lane that owns the stage calls them on its own copy
// of the evaluator.  Plain C++ (the policy supplies the warp primitives) so that tests/hostsim can compile the
// same source with g++; the host policy has one lane and can visit the PAR iterations in reverse order, which
// catches a dependence between iterations without a GPU.
#pragma once

#include "metis_eval.cuh"

namespace metis {

// Scalars a SEQ section hands to the following sections (per-warp scratch next to Scratch<>).
struct CoopMail {
    int m;        // first sub-layer of the backward tail (LayerComputeBalancer state)
    int err;      // METIS_FATAL_* raised inside a SEQ section
    int flag;     // generic boolean result
    int pad;
    double val;   // generic fp64 result (totals)
    double val2;
};

// Host / test policy: one lane.  `reverse` visits PAR iterations last-to-first.
struct OneLane {
    bool reverse = false;
    MB_HD int lane() const { return 0; }
    MB_HD int width() const { return 1; }
    MB_HD bool leader() const { return true; }
    MB_HD void sync() const {}
    MB_HD bool any(bool p) const { return p; }
    MB_HD void argmax_first(double &, int &) const {}
    MB_HD void argmin_first(double &, int &) const {}
    MB_HD void imax_first(int &, int &) const {}
    MB_HD double max_all(double v) const { return v; }
    MB_HD int incl_scan(int v) const { return v; }
    MB_HD int last_lane(int v) const { return v; }
    MB_HD void mark(int) const {}
};

// index of PAR iteration number `i` (0-based count of this lane's iterations) - lets the host policy reverse
template <class X>
MB_HD int par_index(const X &, int i, int n) { (void)n; return i; }
MB_HD int par_index(const OneLane &x, int i, int n) { return x.reverse ? n - 1 - i : i; }

#define METIS_PAR(x, var, n) \
    for (int var##_i = (x).lane(), var = par_index((x), var##_i, (n)); var##_i < (n); \
         var##_i += (x).width(), var = par_index((x), var##_i, (n)))

// ---------------------------------------------------------------------------------------------------------
// Forward fill, backward fill and leftovers of LayerComputeBalancer.run (model/load_balancer.py:216-287) for
// the leader lane: the compare-and-subtract chain in fp64.  Same state encoding as balance_run in metis_eval.cuh
// (fe[] interval ends with kBroke / kTaken, lstk[], blk[], the per-layer packed stage map subw[]).
// in : w.capa[0..S) = stage capacities, w.got[] = 0;  out: w.capa, w.fe, w.lstk, w.got, w.blk, w.subw, *m_out
// ---------------------------------------------------------------------------------------------------------
template <int MAXS, int MAXL>
MB_HD int seq_fill(const Tables &T, int S, Scratch<MAXS, MAXL> &w, int &m_out) {
    const int L = T.p.num_layers;
    const double *dlay = T.dlay;
    const int N = kH * L;
    const int lim = (N - 1 - kH) > 0 ? (N - 1 - kH) : 0;   // :218
    const int last = S - 1;

    // ---- forward pass (:216-231) ----
    int k = 0, sTop = -1;
    bool topSkip = false;
    if (S > 1) {
        int s = 0, j = 0;
        double c = w.capa[0];
        uint8_t *subb = reinterpret_cast<uint8_t *>(w.subw);
#pragma unroll 1
        for (int r = 0; r + 1 < L; ++r) {
            const double d = dlay[r];
            const int nsub = (r == L - 2) ? kH - 1 : kH;     // the last 8 sub-layers are reserved
            if (s >= last) break;
            if (nsub == kH && c > 9.0 * d) {
                // whole layer fits with room to spare: all seven compare-and-subtract steps take the "fits"
                // branch (c - 7d > d even after rounding), so only the subtractions remain
                c -= d; c -= d; c -= d; c -= d; c -= d; c -= d; c -= d;
                w.subw[r] = (uint64_t)s * kOnes;
                j += kH;
                continue;
            }
            int q = 0;
#pragma unroll 1
            while (q < nsub && s < last) {
#pragma unroll 1
                while (q < nsub && c > d) { c -= d; subb[r * 8 + q] = (uint8_t)s; ++q; }
                if (q < nsub) {                                      // sub-layer q does not fit: skipped
                    w.capa[s] = c;
                    w.fe[s] = (uint16_t)((j + q) | kBroke);
                    ++s;
                    c = w.capa[s];
                    ++q;
                }
            }
            j += nsub;
        }
        if (s < last) {                                          // ran into the reserved tail
            w.capa[s] = c;
            w.fe[s] = (uint16_t)lim;
#pragma unroll 1
            for (int t = s + 1; t < last; ++t) w.fe[t] = (uint16_t)lim;
            k = lim;
            sTop = s;
        } else {
            k = (w.fe[last - 1] & kPos) + 1;
            sTop = last - 1;
            topSkip = true;
        }
    }

    // ---- backward pass (:233-249): last stage takes a contiguous tail [m, N) ----
    int m;
    {
        double c = w.capa[last];
        const double dl = dlay[L - 1];
#pragma unroll 1
        for (int i = 0; i < kH; ++i) c -= dl;               // unconditional while len < hallucination (:237-241)
        m = N - kH;
        int sp = S - 2;
#pragma unroll 1
        while (m > 0) {
            const int j = m - 1;
            bool un = (j >= k);
            if (!un) {                                       // below k only skipped sub-layers are unassigned
#pragma unroll 1
                while (sp >= 0 && (!(w.fe[sp] & kBroke) || (int)(w.fe[sp] & kPos) > j)) --sp;
                un = (sp >= 0 && (int)(w.fe[sp] & kPos) == j);
            }
            if (!un) break;                                  // (layer_id + 1) != min(...) from here on (:243)
            const double d = dlay[j / kH];
            if (!(c > d)) break;                             // :246 fails; every later id fails :243
            c -= d;
            m = j;
            if (j < k) w.fe[sp] |= kTaken;
        }
        w.capa[last] = c;
    }

    // ---- leftovers (:251-287), ascending: first the skipped sub-layers, then the middle block ----
    {
        int start = 0;                                        // first sub-layer of stage s's forward interval
#pragma unroll 1
        for (int s = 0; s < last; ++s) {
            const uint16_t e = w.fe[s];
            const int pos = e & kPos;
            const int next_start = pos + ((e & kBroke) ? 1 : 0);
            if ((e & (kBroke | kTaken)) != kBroke) { start = next_start; continue; }
            const int j = pos;
            int lo = 0;
            if (pos > start) {
                lo = s;                                       // common case: stage s itself ends right below j
            } else {
#pragma unroll 1
                for (int u = s;; --u) {
                    if (u < s) {                              // skipped sub-layer of stage u (already placed)
                        const int t = w.lstk[u];
                        const bool above = (t == last) || (fwd_nonempty(w, t) && fwd_start(w, t) > j);
                        if (!above) { lo = t; break; }
                    }
                    if (fwd_nonempty(w, u)) { lo = u; break; }
                    if (u == 0) break;
                }
            }
            int hi = s + 1;
            if (hi < last && !((int)(w.fe[hi] & kPos) > next_start && !w.got[hi])) {
                ++hi;                                         // stage s+1 is empty or already holds a leftover
#pragma unroll 1
                while (hi < last && (!fwd_nonempty(w, hi) || w.got[hi])) ++hi;
            }
            if (lo > hi) return METIS_FATAL_SCRATCH;
            int pick = lo;
            double best = w.capa[lo];
#pragma unroll 1
            for (int t = lo + 1; t <= hi; ++t)
                if (w.capa[t] > best) { best = w.capa[t]; pick = t; }
            w.capa[pick] -= dlay[j / kH];
            w.lstk[s] = (uint8_t)pick;
            w.got[pick] = 1;
            sub_store(w.subw, j, pick);
            start = next_start;
        }
    }
    if (m - k > Scratch<MAXS, MAXL>::kBlock) return METIS_FATAL_SCRATCH;
    {
        int below = -1;                                       // stage of the nearest block item not on `last`
#pragma unroll 1
        for (int t = 0; t < m - k; ++t) {
            const int j = k + t;
            int lo = 0;
            if (below >= 0) lo = below;
            else if (sTop >= 0) {
#pragma unroll 1
                for (int u = sTop;; --u) {
                    if (u < sTop || topSkip) {
                        const uint16_t eu = w.fe[u];
                        if ((eu & (kBroke | kTaken)) == kBroke) {
                            const int t2 = w.lstk[u];
                            if (t2 != last) { lo = t2; break; }   // forward intervals all lie below the block
                        }
                    }
                    if (fwd_nonempty(w, u)) { lo = u; break; }
                    if (u == 0) break;
                }
            }
            int pick = lo;
            double best = w.capa[lo];
#pragma unroll 1
            for (int t2 = lo + 1; t2 <= last; ++t2)
                if (w.capa[t2] > best) { best = w.capa[t2]; pick = t2; }
            w.capa[pick] -= dlay[j / kH];
            w.blk[t] = (uint8_t)pick;
            if (pick != last) below = pick;
            sub_store(w.subw, j, pick);
        }
    }
    m_out = m;
    return METIS_FATAL_NONE;
}

// CPython sum() of w-resident values v[0..n) in index order, for the leader lane (rolled: code size)
MB_HD double seq_py_sum(const double *v, int n) {
    if (n <= 0) return 0.0;
    double f = 0.0 + v[0], c = 0.0;
#pragma unroll 2
    for (int i = 1; i < n; ++i) {
        const double x = v[i];
        const double t = f + x;
        if (fabs(f) >= fabs(x)) c += (f - t) + x;
        else c += (x - t) + f;
        f = t;
    }
    if (c != 0.0 && isfinite(c)) f += c;
    return f;
}

// ---------------------------------------------------------------------------------------------------------
// The chain evaluator.  One instance per lane (registers); `w` and `mail` are the warp's shared scratch.
// ---------------------------------------------------------------------------------------------------------
template <int MAXS, int MAXL, class X>
struct CoopEvaluator : PlanEvaluator<MAXS, MAXL, SerialUniform> {
    using Base = PlanEvaluator<MAXS, MAXL, SerialUniform>;
    using Base::T; using Base::w; using Base::pd; using Base::bs_total; using Base::nbad; using Base::aux;
    X x;
    CoopMail &mail;

    MB_HD CoopEvaluator(const Tables &t, Scratch<MAXS, MAXL> &s, CoopMail &mb, const X &lanes)
        : Base(t, s), x(lanes), mail(mb) {}

    // Start of a plan: groups, rank starts and the first strategy that can be valid (PlanEvaluator::begin,
    // search_space/plan.py:231-249).  The admission pass already dropped plans whose first strategy is invalid.
    MB_HD void begin_coop(const PlanDesc &plan) {
        pd = plan;
        bs_total = T.p.gbs / pd.batches;
        int lb = 0;
#pragma unroll 1
        while ((2 << lb) <= bs_total) ++lb;
        nbad = 0;
        x.sync();                                            // the previous chain of this warp is finished in every lane
        int carry = 0;
#pragma unroll 1
        for (int base = 0; base < pd.S; base += x.width()) { // rank starts: prefix sum of the group sizes
            const int s = base + x.lane();
            const int g = s < pd.S ? pd.row[s] : 0;
            const int sz = s < pd.S ? (1 << g) : 0;
            const int inc = x.incl_scan(sz);
            if (s < pd.S) {
                w.gcode[s] = (uint8_t)g;
                w.tpc[s] = (uint8_t)(g > lb ? g - lb : 0);
                w.rs[s] = (uint16_t)(carry + inc - sz);
            }
            carry += x.last_lane(inc);
        }
        if (x.leader()) w.rs[pd.S] = (uint16_t)carry;
        x.sync();
    }

    // IntraStagePlanGenerator._next_strategy (search_space/plan.py:251-268): the stage with the smallest memory
    // state (or, without a state, the largest dp) that still has dp != 1 halves its dp; first one among equals.
    MB_HD bool next_strategy_coop(bool have_state) {
        int pick = 0x7FFFFFFF;
        if (have_state) {
            double best = INFINITY;
            bool none = true;
            METIS_PAR(x, s, pd.S) {
                if (w.gcode[s] != w.tpc[s] && (none || w.mstate[s] < best || (w.mstate[s] == best && s < pick))) {
                    pick = s; best = w.mstate[s]; none = false;
                }
            }
            // NaN-free: memory states are differences of finite numbers
            x.argmin_first(best, pick);
        } else {
            int best = -1;
            METIS_PAR(x, s, pd.S) {
                const int ldp = (int)w.gcode[s] - (int)w.tpc[s];
                if (ldp != 0 && (ldp > best || (ldp == best && s < pick))) { pick = s; best = ldp; }
            }
            x.imax_first(best, pick);
        }
        if (pick == 0x7FFFFFFF) return false;
        const int g = w.gcode[pick], t = w.tpc[pick];
        nbad += (this->stage_bad(g, t + 1) ? 1 : 0) - (this->stage_bad(g, t) ? 1 : 0);
        x.sync();                                            // every lane has read tpc[pick]
        if (x.leader()) w.tpc[pick] = (uint8_t)(t + 1);
        x.sync();
        return true;
    }

    // first stage (in stage order) whose error mailbox is set: leader scan, rare path
    MB_HD int first_error(const double *box, int n) {
        x.sync();
        if (x.leader()) {
            mail.err = 0;
#pragma unroll 1
            for (int s = 0; s < n; ++s)
                if (box[s] != 0.0) {
                    const uint64_t code = (uint64_t)box[s];
                    mail.err = (int)(code & 0xFF);
                    mail.pad = (int)(code >> 8);
                    break;
                }
        }
        x.sync();
        aux = (uint32_t)mail.pad;
        return mail.err;
    }

    // StagePerformance.get_intra_stage_compute_performance (model/device_group.py:54-85) -> w.perf
    MB_HD int compute_performance_coop() {
        const bool one_type = T.p.num_types == 1;
        bool failed = false;
        x.sync();
        METIS_PAR(x, s, pd.S) {
            const int g = w.gcode[s], tpc = w.tpc[s];
            double p = 0.0;
            int fail = 0;
            int ta = 0, tb = 0;
            if (!one_type) {
                const int a = this->rank_start(s), b = a + (1 << g);
                ta = type_of_rank(T, pd.ns, a); tb = type_of_rank(T, pd.ns, b - 1);
            }
            if (ta == tb) {
                const int bs = bs_total >> (g - tpc);
                const int key = key_of(T, ta, tpc, bs);
                if (key < 0) { aux = ((uint32_t)tpc << 16) | (uint32_t)bs; fail = METIS_FATAL_KEY_EXEC; }
                else if (T.exec_full[key] == 0.0) fail = METIS_FATAL_ZERODIV;
                else p = T.inv_exec[key];                     // 1. / profile_cost
            } else {
                const int a = this->rank_start(s);
                const int rc = this->hetero_performance(a, a + (1 << g), 1 << (g - tpc), tpc, p);
                if (rc) fail = rc;
            }
            w.perf[s] = p;
            w.extra[s] = fail ? (double)fail + (double)aux * 256.0 : 0.0;   // per-stage error mailbox
            failed = failed || fail != 0;
        }
        if (x.any(failed)) return first_error(w.extra, pd.S);
        x.sync();
        if (x.leader()) mail.val = seq_py_sum(w.perf, pd.S);     // sum(compute_performance) (:82)
        x.sync();
        const double tot = mail.val;
        if (tot == 0.0) return METIS_FATAL_ZERODIV;
        METIS_PAR(x, s, pd.S) w.perf[s] = w.perf[s] / tot;
        x.sync();
        return 0;
    }

    // LayerComputeBalancer.run (model/load_balancer.py:197-207): w.perf -> w.part, w.cnt
    MB_HD int balance_coop() {
        const int S = pd.S;
        const int L = T.p.num_layers;
        if (T.p.norm_len < L) return METIS_FATAL_INDEX;       // expand_lc_demand[layer_id] IndexError (:219/:238)
        const double *lc = T.norm_lc;
        const int last = S - 1;
        x.sync();                                            // earlier readers of capa / got are done
        METIS_PAR(x, s, S) { w.capa[s] = w.perf[s]; w.got[s] = 0; }
        x.sync();
        x.mark(10);
        if (x.leader()) {                                    // SEQ: the compare-and-subtract chain
            int m = 0;
            mail.err = seq_fill<MAXS, MAXL>(T, S, w, m);
            mail.m = m;
        }
        x.sync();
        if (mail.err) return mail.err;
        const int m = mail.m;
        x.mark(13);
        // ---- majority vote back to real layers (:290-308), one layer per lane ----
        const int nw = (L + 7) / 8;
        METIS_PAR(x, r, nw * 8) {
            int own = kDropped;                              // padding of the last owner word
            if (r < L) {
                const int nlow = m - kH * r;                 // sub-layers of r below the backward tail
                if (nlow <= 0) {
                    own = last;
                } else {
                    uint64_t v = w.subw[r];
                    if (nlow < kH) {
                        const uint64_t mask = (1ULL << (8 * nlow)) - 1ULL;
                        v = (v & mask) | (((uint64_t)last * kOnes) & ~mask);
                    }
                    v |= 0xFF00000000000000ULL;
                    const int c3 = (int)((v >> 24) & 0xFF);
                    if (swar_count(v, c3) * 2 > kH) own = c3;     // count > hallucination / 2 (:295)
                    else {
                        const int c0 = (int)(v & 0xFF), c1 = (int)((v >> 8) & 0xFF), c2 = (int)((v >> 16) & 0xFF);
                        if (swar_count(v, c0) * 2 > kH) own = c0;
                        else if (c1 != c0 && swar_count(v, c1) * 2 > kH) own = c1;
                        else if (c2 != c1 && c2 != c0 && swar_count(v, c2) * 2 > kH) own = c2;
                    }
                }
            }
            reinterpret_cast<uint8_t *>(w.ownerw)[r] = (uint8_t)own;
        }
        x.sync();
        x.mark(14);
        // ---- first / last / count of the layers of each stage and the spare capacity (:300-306) ----
        METIS_PAR(x, s, S) {
            int n = 0, fi = 0, la = 0;
#pragma unroll 1
            for (int k = 0; k < nw; ++k) {
                const uint64_t z = swar_eq(w.ownerw[k], s) & 0x8080808080808080ULL;
                if (z) {
                    if (n == 0) fi = 8 * k + (ctz64(z) >> 3);
                    la = 8 * k + ((63 - clz64(z)) >> 3);
                    n += popc64(z);
                }
            }
            w.cnt[s] = (uint16_t)n; w.first[s] = (uint16_t)fi; w.lastl[s] = (uint16_t)la;
            w.capa[s] = n ? w.perf[s] - py_sum_range_compact(lc, fi, la + 1) : w.perf[s];
        }
        x.sync();
        x.mark(15);
        // ---- boundary adjustment (:310-356): at most three committed single-layer moves ----
        uint8_t *owner = reinterpret_cast<uint8_t *>(w.ownerw);
#pragma unroll 1
        for (int n = 1; n <= 3; ++n) {
            int top = 0x7FFFFFFF;
            double maxc = -INFINITY;
            METIS_PAR(x, t, S)                               // stable: lowest index among equal maxima (:329-331)
                if (w.capa[t] > maxc || (w.capa[t] == maxc && t < top)) { maxc = w.capa[t]; top = t; }
            x.argmax_first(maxc, top);
            if (top == 0x7FFFFFFF) top = 0;
            int nb = -1;
            double val = INFINITY;
            if (top - 1 >= 0 && w.capa[top - 1] < val) { nb = top - 1; val = w.capa[top - 1]; }
            if (top + 1 < S && w.capa[top + 1] < val) { nb = top + 1; }
            if (nb < 0 || w.cnt[nb] <= 1) break;             // no-op rounds leave the state unchanged
            const int layer = (top > nb) ? w.lastl[nb] : w.first[nb];
            const double dl = lc[layer];
            const double ntop = w.capa[top] - dl;
            const double nnb = w.capa[nb] + dl;
            double newmax = -INFINITY;
            METIS_PAR(x, t, S) {
                const double v = (t == top) ? ntop : (t == nb) ? nnb : w.capa[t];
                if (v > newmax) newmax = v;
            }
            newmax = x.max_all(newmax);
            if (newmax > maxc) break;                        // :352 (not committed)
            x.sync();                                        // every lane has read the state of this round
            if (x.leader()) {
                owner[layer] = (uint8_t)top;
                w.capa[top] = ntop;
                w.capa[nb] = nnb;
                if (top > nb) { int r = layer - 1; while (owner[r] != nb) --r; w.lastl[nb] = (uint16_t)r; }
                else          { int r = layer + 1; while (owner[r] != nb) ++r; w.first[nb] = (uint16_t)r; }
                if (w.cnt[top] == 0) { w.first[top] = (uint16_t)layer; w.lastl[top] = (uint16_t)layer; }
                else {
                    if (layer < (int)w.first[top]) w.first[top] = (uint16_t)layer;
                    if (layer > (int)w.lastl[top]) w.lastl[top] = (uint16_t)layer;
                }
                ++w.cnt[top];
                --w.cnt[nb];
            }
            x.sync();
        }
        x.mark(16);
        // ---- partition = cumulative layer counts (:358-364) ----
        int carry = 0;
#pragma unroll 1
        for (int base = 0; base < S; base += x.width()) {
            const int s = base + x.lane();
            const int v = s < S ? (int)w.cnt[s] : 0;
            const int inc = x.incl_scan(v);
            if (s < S) w.part[s + 1] = (uint16_t)(carry + inc);
            carry += x.last_lane(inc);
        }
        if (x.leader()) w.part[0] = 0;
        x.sync();
        return METIS_FATAL_NONE;
    }

    // LayerLoadBalancer._adj_compute_performance (model/load_balancer.py:71-107)
    // in: w.perf (c_capa), w.extra (m_demand); out: w.perf; returns 1 = None, 0 ok, <0 fatal (negated code)
    MB_HD int adjust_performance_coop() {
        const int S = pd.S;
        const bool one_type = T.p.num_types == 1;
        double *ratio = reinterpret_cast<double *>(w.subw);      // free after the vote (MAXL >= MAXS)
        double *mcap = ratio + MAXS / 2;                          // stage memory capacity; subw has MAXL >= 2*(MAXS/2).. see static_assert
        static_assert(MAXL >= MAXS, "subw doubles as two S-sized fp64 arrays only when MAXL >= 2 * MAXS / 2");
        (void)mcap;
        x.sync();
        METIS_PAR(x, s, S) {                                     // independent per stage (:80-89)
            const int a = one_type ? 0 : this->rank_start(s), b = a + this->group(s);
            const double c = w.perf[s], md = w.extra[s];
            const double mc = one_type ? T.type_memory[0] * (double)this->group(s) : this->memory_capacity(a, b);
            double av, adj;
            if (mc > md) {
                adj = c;
                av = (c * mc / md) - c;
            } else {
                av = 0.0;
                adj = c * (mc / md) * 0.9;
            }
            w.capa[s] = av;           // available_compute_capacity
            w.mstate[s] = adj;        // adj_sc_capa
            ratio[s] = (mc > md) ? 0.0 : (c - adj);          // this stage's term of extra_required_capacity (:89)
        }
        x.sync();
        if (x.leader()) {                                        // order-dependent accumulations (:89-91)
            double need = 0.;
#pragma unroll 1
            for (int s = 0; s < S; ++s)
                if (w.capa[s] == 0.0 && ratio[s] != 0.0) need += ratio[s];
                else if (ratio[s] != 0.0) need += ratio[s];
            mail.val = need;
            mail.flag = seq_py_sum(w.capa, S) < need ? 1 : 0;
        }
        x.sync();
        if (mail.flag) return 1;
        METIS_PAR(x, s, S) w.extra[s] = 0.;
        x.sync();
        if (x.leader()) {                                        // :96-104, sequential: `need` changes as it goes
            double need = mail.val;
            int guard = 0;
            mail.err = 0;
#pragma unroll 1
            while (need > 0.01) {
                PySum tot;
#pragma unroll 1
                for (int s = 0; s < S; ++s) tot.add(w.capa[s] > 0.001 ? w.perf[s] : 0.0);
                const double tmp_total = tot.result();
#pragma unroll 1
                for (int s = 0; s < S; ++s)                      // c_capa_ratio list (:98), before the updates
                    ratio[s] = w.capa[s] > 0.001 ? w.perf[s] / tmp_total : 0.0;
#pragma unroll 1
                for (int s = 0; s < S; ++s) {
                    const double av = w.capa[s];
                    const double want = need * ratio[s];
                    const double give = want > av ? av : want;
                    w.extra[s] += give;
                    w.capa[s] -= give;
                    need -= give;
                }
                if (++guard > 4096) { mail.err = METIS_FATAL_HANG; break; }
            }
        }
        x.sync();
        if (mail.err) return -mail.err;
        METIS_PAR(x, s, S) w.perf[s] = w.extra[s] + w.mstate[s];
        x.sync();
        return 0;
    }

    // One attempt of LayerLoadBalancer.partition_layer after the balancer (model/load_balancer.py:127-143):
    // memory demand (:29-55), OOM test (:57-63), capacity re-weighting.  Returns like PlanEvaluator::memory_phase.
    MB_HD int memory_phase_coop(int attempt) {
        const int S = pd.S;
        const bool one_type = T.p.num_types == 1;
        const int type0 = T.run_type[pd.ns * T.p.num_types];
        bool failed = false, oom = false;
        x.sync();                                            // the balancer's last readers of capa / extra / mstate are done
        METIS_PAR(x, s, S) {
            const int g = w.gcode[s], tpc = w.tpc[s];
            const int a = one_type ? 0 : this->rank_start(s), b = a + (1 << g);
            double md = 0.001, err = 0.0;
            if (one_type || type_of_rank(T, pd.ns, a) == type_of_rank(T, pd.ns, b - 1)) {
                const int bs = bs_total >> (g - tpc);
                const int key = key_of(T, type0, tpc, bs);
                if (key < 0) err = (double)METIS_FATAL_KEY_MEMORY + (double)(((uint32_t)tpc << 16) | (uint32_t)bs) * 256.0;
                else md += py_sum_range_compact(T.mem + (size_t)key * T.p.lpad, w.part[s], w.part[s + 1]) * kMemCoef;
            } else {
                const int rc = this->hetero_memory_demand(s, type0, md);
                if (rc) err = (double)rc + (double)aux * 256.0;
            }
            const double state = one_type ? T.type_memory[0] * (double)(1 << g) - md : this->memory_capacity(a, b) - md;
            w.extra[s] = md;
            w.capa[s] = state;
            w.mstate[s] = err;
            failed = failed || err != 0.0;
            oom = oom || state < 0;
        }
        failed = x.any(failed);
        oom = x.any(oom);
        if (failed) return -first_error(w.mstate, S);
        x.sync();
        if (!oom) {
            METIS_PAR(x, s, S) w.mstate[s] = w.capa[s];
            x.sync();
            return 1;
        }
        if (attempt >= 3) return 0;
        x.mark(21);
        const int rc = adjust_performance_coop();
        if (rc < 0) return rc;
        return rc == 1 ? 0 : 2;
    }

    // HeteroCostEstimator.get_cost (model/cost_estimator.py:199-244); returns 0 ok, 1 KeyError.  The cost lands in
    // mail.val (every lane reads it after the final sync).
    MB_HD int get_cost_coop(double &cost_out) {
        const int per = T.p.devices_per_node;
        const int Lm = T.p.num_layers;
        const bool one_type = T.p.num_types == 1;
        const bool ubw = T.p.uniform_bw != 0;
        const int nstage = pd.label < pd.S ? pd.label : pd.S;  // zip(range(plan.num_stage), strategies)
        double *ppterm = reinterpret_cast<double *>(w.subw);  // free after the vote (MAXL >= MAXS)
        bool bad = false;
        double max_len = -INFINITY, max_upd = -INFINITY, max_dp = -INFINITY;
        x.sync();
        METIS_PAR(x, s, nstage) {
            const int g = w.gcode[s], tpc = w.tpc[s];
            const int a = one_type ? 0 : this->rank_start(s), b = a + (1 << g);
            const int la = w.part[s], lb = w.part[s + 1];
            const int ldp = g - tpc;
            const int mbs = bs_total >> ldp;
            const int ta = one_type ? 0 : type_of_rank(T, pd.ns, a);
            const int tb = one_type ? 0 : type_of_rank(T, pd.ns, b - 1);
            double len = 0.0;
            if (ta == tb) {                                   // _get_execution_cost :175-188
                const int key = key_of(T, ta, tpc, mbs);
                if (key < 0) bad = true;
                else len = py_sum_range_compact(T.lc + (size_t)key * T.p.lpad, la, lb);
            } else if (this->hetero_exec_cost(a, b, 1 << ldp, tpc, la, lb, len)) {
                bad = true;
            }
            w.capa[s] = len;
            if (len > max_len) max_len = len;
            const double inv_tp = pow2_neg(tpc);              // 1 / tp, exact power of two
            double pp = 0.0;
            if (s < nstage - 1) {
                if (ubw) {                                    // :224-227 via the derived tables
                    pp = (lb == Lm - 1) ? T.pp_vocab[mbs * T.p.num_tp + tpc] : T.pp_hidden[mbs];
                } else {
                    double act;
                    if (lb == Lm - 1)
                        act = (double)((int64_t)mbs * T.p.sequence_length * T.p.vocab_size) * inv_tp;
                    else
                        act = (double)((int64_t)mbs * T.p.sequence_length * T.p.hidden_size);
                    const int a2 = this->rank_start(s), b2 = this->rank_start(s + 2);
                    pp = act / (this->bw_of_node_range(a2 / per, (b2 - 1) / per) * 1048576.0);
                }
            }
            ppterm[s] = pp;
            // get_parameter_size_by_stage (model/activation_parameter.py:40-51)
            int ntr = lb - la;
            double params = 0.0;
            if (la == 0) { params += T.p.input_params * inv_tp; --ntr; }
            if (lb == Lm) { params += T.p.output_params * inv_tp; --ntr; }
            params += T.p.transformer_params * inv_tp * (double)ntr;
            double dpc;                                       // :37-43
            if (ubw) dpc = T.dpk[ldp] * params;
            else {
                const int dp = 1 << ldp;
                dpc = (double)(2 * (dp - 1)) / ((double)dp * (this->dp_bandwidth(this->rank_start(s), dp, 1 << tpc) * 1048576.0)) * params;
            }
            if (dpc > max_dp) max_dp = dpc;
            const double upd = T.p.optimizer_time * inv_tp * T.ratio[lb - la];   // :145-147
            if (upd > max_upd) max_upd = upd;
        }
        if (x.any(bad)) return 1;                             // KeyError raised while costing a stage
        max_len = x.max_all(max_len);
        max_upd = x.max_all(max_upd);
        max_dp = x.max_all(max_dp);
        x.sync();
        if (x.leader()) {                                     // order-dependent sums, stage order
            double pp_cost = 0.;
#pragma unroll 1
            for (int s = 0; s + 1 < nstage; ++s) pp_cost += ppterm[s];
            const double lens = seq_py_sum(w.capa, nstage);
            const int s = nstage - 1;                         // _get_fb_sync_cost of the last costed stage
            const int a = one_type ? 0 : this->rank_start(s), b = a + this->group(s);
            double v = 0.0;
            mail.flag = this->fb_sync_cost(a, b, w.tpc[s], bs_total >> (w.gcode[s] - w.tpc[s]), v);
            const double fb_sync = v * (double)pd.batches;
            const double exec = ((double)(pd.batches - 1) * max_len) + lens;            // :235-236
            const double bg = T.p.batch_generator * (double)pd.batches;
            mail.val = exec + fb_sync + max_upd + max_dp + pp_cost + bg;                // :241-242
        }
        x.sync();
        if (mail.flag) return 1;
        cost_out = mail.val;
        return 0;
    }

    // cost_het_cluster.py:31-48 for one inter-stage plan with IntraStagePlanGenerator.has_next
    // (search_space/plan.py:192-226) inlined: the whole chain, depth first.  `skip_first` = the first partition
    // attempt was already counted by the caller's first-task round (it is recomputed, not recounted).
    template <class Sink>
    MB_HD void run_chain(const PlanDesc &plan, Sink &sink, bool skip_first) {
        begin_coop(plan);
        bool started = false, have_state = false;
        int nrep = 0, step = 0;
#pragma unroll 1
        for (;;) {
            if (nrep == 1) return;                            // plan.py:194-195
            int attempt = 0;
#pragma unroll 1
            for (;;) {
                if (!started) started = true;                 // first strategy that can be valid (see begin)
                else if (!next_strategy_coop(have_state)) return;  // :203-204
                if (!this->valid()) continue;
                if (!skip_first) sink.partition_call();
                x.mark(2);
                int rc = compute_performance_coop();
                if (rc) { sink.fatal(pd.ordinal, rc, aux); return; }
                attempt = 0;
#pragma unroll 1
                for (int a = 1; a <= 3; ++a) {                // LayerLoadBalancer.partition_layer (:121-144)
                    if (!skip_first) sink.balancer_run();
                    skip_first = false;
                    rc = balance_coop();
                    if (rc) { sink.fatal(pd.ordinal, rc, aux); return; }
                    x.mark(20);
                    const int r = memory_phase_coop(a);
                    if (r < 0) { sink.fatal(pd.ordinal, -r, aux); return; }
                    if (r == 1) { attempt = a; break; }
                    if (r == 0) break;
                }
                skip_first = false;
                have_state = attempt > 0;                     // memory_state is None after a failure (:225)
                if (attempt > 0) break;
            }
            nrep = attempt;
            x.mark(22);
            double cost = 0.0;
            if (get_cost_coop(cost) == 0) sink.emit(pd, step, nrep, cost, w.tpc, w.part);
            else sink.keyerror();
            x.sync();                                         // the leader's record is written before the state changes
            ++step;
        }
    }
};

}  // namespace metis
