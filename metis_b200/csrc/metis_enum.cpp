// metis_enum.cpp - host-side enumeration of device-group rows in the reference's order.
//
// Restates gen_dgroups_for_stages_with_variance (search_space/device_group.py:93-107):
// power-of-two shapes filtered by the variance floor, non-decreasing compositions of
// num_gpus in lexicographic shape order (:58-81), pair-merging of the smallest groups down
// to max_permute_len (:7-55) and the multiset permutations of the merged groups in the
// prefix-shift order of Williams' algorithm (search_space/utils.py:56-88).  `dg_idx` in the
// reference is the position in this list, so the order is part of the contract.
//
// Structure: (1) list the compositions, (2) merge each one and count its permutations in closed
// form (n! / prod mult!), (3) prefix-sum the counts into row offsets, (4) generate the
// permutations of disjoint composition ranges on several host threads straight into the output.
// This is enumeration (integer tuples) - the candidate evaluation itself only runs on the GPU.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "../../include/metis_b200.h"

namespace {

using Group = std::vector<int>;   // one merged group = tuple of device-group sizes

int group_sum(const Group &g) {
    int s = 0;
    for (int v : g) s += v;
    return s;
}

uint8_t ilog2(int v) {
    uint8_t c = 0;
    while ((1 << c) < v) ++c;
    return c;
}

// permute() of search_space/device_group.py:7-55 without the final permutations: the merged groups
std::vector<Group> merge_groups(const std::vector<int> &comp, int max_permute_len) {
    std::vector<Group> groups;
    groups.reserve(comp.size());
    for (int v : comp) groups.push_back(Group{v});
    int num_reduce = (int)groups.size() - max_permute_len;
    while (num_reduce > 0) {
        const int min_size = group_sum(groups[0]);
        int num_min = (int)groups.size();                  // find_num_min (:8-12)
        for (int idx = 0; idx < (int)groups.size(); ++idx)
            if (groups[idx] != groups[0]) { num_min = idx + 1; break; }
        if (num_min / 2 > num_reduce) num_reduce = num_min / 2;      // :26-27
        std::vector<Group> merged;
        for (int i = 0; i < (int)groups.size(); i += 2) {             // :31-45
            if (num_reduce <= i / 2) {
                merged.insert(merged.end(), groups.begin() + i, groups.end());
                break;
            }
            if (i + 1 >= (int)groups.size()) {
                merged.push_back(groups[i]);
            } else if (group_sum(groups[i]) == min_size && group_sum(groups[i]) == group_sum(groups[i + 1])) {
                Group g = groups[i];
                g.insert(g.end(), groups[i + 1].begin(), groups[i + 1].end());
                merged.push_back(g);
            } else {
                merged.push_back(groups[i]);
                merged.push_back(groups[i + 1]);
            }
        }
        groups.swap(merged);
        if (num_reduce == (int)groups.size() - max_permute_len) break;   // :48-50
        num_reduce = (int)groups.size() - max_permute_len;
    }
    std::sort(groups.begin(), groups.end());               // utils.py:57 (tuple comparison == lexicographic)
    return groups;
}

// number of distinct permutations of a sorted multiset: n! / prod(multiplicity!)
int64_t multiset_permutation_count(const std::vector<Group> &sorted_items) {
    int64_t total = 1, placed = 0;
    size_t i = 0;
    while (i < sorted_items.size()) {
        size_t j = i;
        while (j < sorted_items.size() && sorted_items[j] == sorted_items[i]) ++j;
        for (int64_t k = 1; k <= (int64_t)(j - i); ++k) total = total * (placed + k) / k;   // *= C(placed+k, k)
        placed += (int64_t)(j - i);
        i = j;
    }
    return total;
}

// multiset permutations of the sorted groups (search_space/utils.py:72-88), visiting order preserved;
// rows (log2 codes, `stages` bytes each) are written consecutively starting at dst
void williams_rows(const std::vector<Group> &items, int stages, uint8_t *dst) {
    const int n = (int)items.size();
    std::vector<int> rank(n), nxt(n, -1), off(n), len(n);
    std::vector<uint8_t> codes;
    for (int k = 0; k < n; ++k) {
        rank[k] = (k > 0 && items[k] == items[k - 1]) ? rank[k - 1] : k;   // equal tuples compare equal
        off[k] = (int)codes.size();
        len[k] = (int)items[k].size();
        for (int v : items[k]) codes.push_back(ilog2(v));
    }
    int head = 0;
    for (int k = 1; k < n; ++k) { nxt[k] = head; head = k; }              // prepend => non-increasing chain
    auto visit = [&]() {
        uint8_t *p = dst;
        for (int h = head; h != -1; h = nxt[h]) { memcpy(p, codes.data() + off[h], (size_t)len[h]); p += len[h]; }
        dst += stages;
    };
    auto nth = [&](int h, int k) {
        while (k > 0 && nxt[h] != -1) { h = nxt[h]; --k; }
        return h;
    };
    int i = nth(head, n - 2), j = nth(head, n - 1);
    visit();
    while (nxt[j] != -1 || rank[j] < rank[head]) {
        const int s = (nxt[j] != -1 && rank[i] >= rank[nxt[j]]) ? j : i;
        const int t = nxt[s];
        nxt[s] = nxt[t];
        nxt[t] = head;
        if (rank[t] < rank[head]) i = t;
        j = nxt[i];
        head = t;
        visit();
    }
}

// gen_dgroups_recursive (:58-81): non-decreasing compositions, lexicographic in shape index
void list_compositions(int stages, int gpus, const std::vector<int> &shapes, std::vector<std::vector<int>> &out) {
    if (shapes.empty()) return;
    std::vector<int> sol;
    sol.reserve(stages);
    const int lo = shapes.front(), hi = shapes.back();
    struct Frame { int next_idx; int sum; };
    std::vector<Frame> stack;
    stack.push_back({0, 0});
    while (!stack.empty()) {
        const int depth = (int)stack.size() - 1;              // elements already chosen
        if (depth == stages) {
            if (stack.back().sum == gpus) out.push_back(sol);
            stack.pop_back();
            if (!sol.empty()) sol.pop_back();
            continue;
        }
        bool descended = false;
        while (stack.back().next_idx < (int)shapes.size()) {
            Frame &f = stack.back();
            const int i = f.next_idx++;
            const int g = shapes[i];
            if (g + f.sum > gpus) { f.next_idx = (int)shapes.size(); break; }   // :73-74
            const int remaining = stages - depth - 1, rest = gpus - f.sum - g;
            if (hi * remaining < rest || lo * remaining > rest) continue;       // :61-66 pruning
            const int sum = f.sum + g;
            sol.push_back(g);
            stack.push_back({i, sum});
            descended = true;
            break;
        }
        if (!descended) {
            stack.pop_back();
            if (!sol.empty()) sol.pop_back();
        }
    }
}

// One stage count: prepare (compositions, merged groups, row offsets), then fill rows.
struct StageTable {
    int stages = 0;
    std::vector<std::vector<int>> comps;
    std::vector<std::vector<Group>> merged;
    std::vector<int64_t> offset;          // row offset of each composition, size merged.size()+1
    int64_t rows() const { return offset.empty() ? 0 : offset.back(); }
};

void list_stage(StageTable &t, int num_stages, int num_gpus, double variance) {
    t.stages = num_stages;
    const int share = std::max(num_gpus / num_stages, num_stages / num_gpus);   // :96-98
    const double floor_share = (double)share * variance;
    std::vector<int> shapes;
    for (int s = 1; s <= num_gpus; s <<= 1)
        if ((double)s >= floor_share) shapes.push_back(s);
    list_compositions(num_stages, num_gpus, shapes, t.comps);
    t.merged.resize(t.comps.size());
    t.offset.assign(t.comps.size() + 1, 0);
}

template <class F>
void parallel_for(int64_t n, int64_t grain, F body) {        // body(begin, end) on chunks of `grain` items
    static const unsigned configured = []() {
        const char *env = getenv("METIS_ENUM_THREADS");
        unsigned v = env ? (unsigned)atoi(env) : 8u;        // 8 host threads measured best on the B200 hosts
        const unsigned hw = std::thread::hardware_concurrency();
        if (hw && v > hw) v = hw;
        return v ? v : 1u;
    }();
    unsigned nthreads = configured;
    const int64_t chunks = (n + grain - 1) / grain;
    if (nthreads < 2 || chunks < 2) { if (n > 0) body((int64_t)0, n); return; }
    if ((int64_t)nthreads > chunks) nthreads = (unsigned)chunks;
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nthreads; ++t)
        pool.emplace_back([=]() {
            for (int64_t c = t; c < chunks; c += nthreads) body(c * grain, std::min(n, (c + 1) * grain));
        });
    for (auto &th : pool) th.join();
}

// All stage counts of a range: compositions per stage count, then merge + count and row generation
// over the flattened (stage count, composition) list so the host threads are evenly loaded.
struct TableSet {
    int first = 0;
    std::vector<StageTable> tables;
    std::vector<std::pair<int, int>> items;                  // (table index, composition index)

    void prepare(int first_stage, int last_stage, int num_gpus, double variance, int max_permute_len) {
        first = first_stage;
        const int n = last_stage - first_stage + 1;
        tables.assign(n, StageTable());
        parallel_for(n, 1, [&](int64_t b, int64_t e) {
            for (int64_t i = b; i < e; ++i) list_stage(tables[i], first_stage + (int)i, num_gpus, variance);
        });
        items.clear();
        for (int i = 0; i < n; ++i)
            for (int c = 0; c < (int)tables[i].comps.size(); ++c) items.emplace_back(i, c);
        parallel_for((int64_t)items.size(), 256, [&](int64_t b, int64_t e) {
            for (int64_t k = b; k < e; ++k) {
                StageTable &t = tables[items[k].first];
                const int c = items[k].second;
                t.merged[c] = merge_groups(t.comps[c], max_permute_len);
                t.offset[c + 1] = multiset_permutation_count(t.merged[c]);   // count; prefix-summed below
            }
        });
        for (StageTable &t : tables)
            for (size_t c = 0; c < t.merged.size(); ++c) t.offset[c + 1] += t.offset[c];
    }

    void fill(const std::vector<int64_t> &byte_off, uint8_t *out) const {
        parallel_for((int64_t)items.size(), 64, [&](int64_t b, int64_t e) {
            for (int64_t k = b; k < e; ++k) {
                const StageTable &t = tables[items[k].first];
                const int c = items[k].second;
                williams_rows(t.merged[c], t.stages, out + byte_off[items[k].first] + t.offset[c] * t.stages);
            }
        });
    }
};

}  // namespace

extern "C" int64_t metis_enum_device_groups(int32_t num_stages, int32_t num_gpus, double variance,
                                            int32_t max_permute_len, uint8_t *out, int64_t capacity_rows) {
    if (num_stages < 1 || num_gpus < 1 || max_permute_len < 1) return METIS_E_ARG;
    TableSet set;
    set.prepare(num_stages, num_stages, num_gpus, variance, max_permute_len);
    const int64_t rows = set.tables[0].rows();
    if (!out) return rows;
    if (rows > capacity_rows) return METIS_E_CAPACITY;
    set.fill(std::vector<int64_t>{0, rows * num_stages}, out);
    return rows;
}

extern "C" int64_t metis_enum_device_group_tables(int32_t first_stage, int32_t last_stage, int32_t num_gpus,
                                                  double variance, int32_t max_permute_len, int64_t *rows_per_stage,
                                                  uint8_t *out, int64_t capacity_bytes) {
    if (first_stage < 1 || last_stage < first_stage || num_gpus < 1 || max_permute_len < 1 || !rows_per_stage)
        return METIS_E_ARG;
    const int n = last_stage - first_stage + 1;
    // the sizing call (out == NULL) and the filling call that follows it share the prepared tables
    struct Prepared { int first, last, gpus, mpl; double variance; TableSet set; };
    static thread_local Prepared cache{0, 0, 0, 0, 0.0, {}};
    const bool hit = cache.first == first_stage && cache.last == last_stage && cache.gpus == num_gpus &&
                     cache.mpl == max_permute_len && cache.variance == variance && (int)cache.set.tables.size() == n;
    if (!hit) {
        cache.first = first_stage; cache.last = last_stage; cache.gpus = num_gpus; cache.mpl = max_permute_len;
        cache.variance = variance;
        cache.set.prepare(first_stage, last_stage, num_gpus, variance, max_permute_len);
    }
    const TableSet &set = cache.set;
    std::vector<int64_t> byte_off(n + 1, 0);
    for (int i = 0; i < n; ++i) {
        rows_per_stage[i] = set.tables[i].rows();
        byte_off[i + 1] = byte_off[i] + set.tables[i].rows() * (first_stage + i);
    }
    if (!out) return byte_off[n];
    if (byte_off[n] > capacity_bytes) return METIS_E_CAPACITY;
    set.fill(byte_off, out);
    const int64_t total_bytes = byte_off[n];
    cache.first = cache.last = 0;                            // release the prepared tables
    cache.set = TableSet();
    return total_bytes;
}
