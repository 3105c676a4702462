// metis_enum.cpp - host-side enumeration of device-group rows in the reference's order.
//
// Restates gen_dgroups_for_stages_with_variance (search_space/device_group.py:93-107):
// power-of-two shapes filtered by the variance floor, non-decreasing compositions of
// num_gpus in lexicographic shape order (:58-81), pair-merging of the smallest groups down
// to max_permute_len (:7-55) and the multiset permutations of the merged groups in the
// prefix-shift order of Williams' algorithm (search_space/utils.py:56-88).  `dg_idx` in the
// reference is the position in this list, so the order is part of the contract.
//
// This is enumeration (integer tuples) - the candidate evaluation itself only runs on the GPU.
#include <stdint.h>

#include <algorithm>
#include <vector>

#include "../../include/metis_b200.h"

namespace {

using Group = std::vector<int>;   // one merged group = tuple of device-group sizes

int group_sum(const Group &g) {
    int s = 0;
    for (int v : g) s += v;
    return s;
}

int ilog2(int v) {
    int c = 0;
    while ((1 << c) < v) ++c;
    return c;
}

struct RowWriter {
    uint8_t *out;
    int64_t capacity;
    int64_t count;
    int stages;
    bool overflow;
    void push(const std::vector<const Group *> &perm) {
        if (out) {
            if (count >= capacity) { overflow = true; ++count; return; }
            uint8_t *dst = out + count * stages;
            int k = 0;
            for (const Group *g : perm)
                for (int v : *g) dst[k++] = (uint8_t)ilog2(v);
        }
        ++count;
    }
};

// multiset permutations of `items` (search_space/utils.py:72-88), visiting order preserved
void williams(std::vector<Group> items, RowWriter &w) {
    std::sort(items.begin(), items.end());                 // utils.py:57 (tuple comparison == lexicographic)
    const int n = (int)items.size();
    std::vector<int> nxt(n, -1);
    int head = 0;
    for (int k = 1; k < n; ++k) { nxt[k] = head; head = k; }   // prepend => non-increasing chain
    std::vector<const Group *> perm(n);
    auto visit = [&]() {
        int h = head, k = 0;
        while (h != -1) { perm[k++] = &items[h]; h = nxt[h]; }
        w.push(perm);
    };
    auto nth = [&](int h, int k) {
        while (k > 0 && nxt[h] != -1) { h = nxt[h]; --k; }
        return h;
    };
    int i = nth(head, n - 2), j = nth(head, n - 1);
    visit();
    while (nxt[j] != -1 || items[j] < items[head]) {
        int s = (nxt[j] != -1 && !(items[i] < items[nxt[j]])) ? j : i;
        const int t = nxt[s];
        nxt[s] = nxt[t];
        nxt[t] = head;
        if (items[t] < items[head]) i = t;
        j = nxt[i];
        head = t;
        visit();
    }
}

// permute() of search_space/device_group.py:7-55
void merge_and_permute(const std::vector<int> &comp, int max_permute_len, RowWriter &w) {
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
    williams(groups, w);
}

// gen_dgroups_recursive (:58-81): non-decreasing compositions, lexicographic in shape index
void compositions(int stages, int gpus, const std::vector<int> &shapes, int max_permute_len, RowWriter &w) {
    if (shapes.empty()) return;
    std::vector<int> sol;
    sol.reserve(stages);
    const int lo = shapes.front(), hi = shapes.back();
    // iterative DFS keeps the reference's visiting order: for i in range(prev, len(shapes))
    struct Frame { int next_idx; int sum; };
    std::vector<Frame> stack;
    stack.push_back({0, 0});
    while (!stack.empty()) {
        Frame &f = stack.back();
        const int depth = (int)stack.size() - 1;              // elements already chosen
        if (depth == stages) {
            if (f.sum == gpus) merge_and_permute(sol, max_permute_len, w);
            stack.pop_back();
            if (!sol.empty()) sol.pop_back();
            continue;
        }
        bool descended = false;
        while (f.next_idx < (int)shapes.size()) {
            const int i = f.next_idx++;
            const int g = shapes[i];
            if (g + f.sum > gpus) { f.next_idx = (int)shapes.size(); break; }   // :73-74
            const int remaining = stages - depth - 1, rest = gpus - f.sum - g;
            if (hi * remaining < rest || lo * remaining > rest) continue;       // :61-66 pruning
            sol.push_back(g);
            stack.push_back({i, f.sum + g});
            descended = true;
            break;
        }
        if (!descended) {
            stack.pop_back();
            if (!sol.empty()) sol.pop_back();
        }
    }
}

}  // namespace

extern "C" int64_t metis_enum_device_groups(int32_t num_stages, int32_t num_gpus, double variance,
                                            int32_t max_permute_len, uint8_t *out, int64_t capacity_rows) {
    if (num_stages < 1 || num_gpus < 1 || max_permute_len < 1) return METIS_E_ARG;
    const int share = std::max(num_gpus / num_stages, num_stages / num_gpus);   // :96-98
    const double floor_share = (double)share * variance;
    std::vector<int> shapes;
    for (int s = 1; s <= num_gpus; s <<= 1)
        if ((double)s >= floor_share) shapes.push_back(s);
    RowWriter w{out, capacity_rows, 0, num_stages, false};
    compositions(num_stages, num_gpus, shapes, max_permute_len, w);
    if (w.overflow) return METIS_E_CAPACITY;
    return w.count;
}
