// metis_rows.cuh - device-group rows of one composition.
//
// The reference turns every composition (a non-decreasing list of power-of-two stage sizes, merged into at most
// max_permute_len groups) into rows by visiting the multiset permutations of the groups with the prefix-shift walk of
// search_space/utils.py:56-88: the linked list starts non-increasing and every step moves one node to the front.
// The host lists the compositions (metis_enum_compositions); this routine replays the walk for one of them and
// writes the rows - log2 of every stage's device count, `stages` bytes per row - at the composition's offset in the
// row blob.  Integer work only; shared by the CUDA kernel (one thread per composition) and the host test build.
#pragma once

#include "metis_eval.cuh"

namespace metis {

// The walk is sequential (each step moves one node of a linked list to the front); writing a row out is not, so
// the walk is a little state machine: init, then compose / advance until advance() returns false.
struct CompWalk {
    int8_t nxt[METIS_MAX_PERMUTE_GROUPS];
    uint8_t rank[METIS_MAX_PERMUTE_GROUPS], len[METIS_MAX_PERMUTE_GROUPS];
    uint16_t off[METIS_MAX_PERMUTE_GROUPS];
    const uint8_t *codes;
    int n, head, i, j;

    MB_HD void init(const MetisCompRec &rec, const uint8_t *pool) {
        n = rec.num_groups;
        const uint8_t *lens = pool + rec.pool_offset;
        codes = lens + n;
        int o = 0;
        for (int k = 0; k < n; ++k) {
            len[k] = lens[k];
            off[k] = (uint16_t)o;
            bool same = k > 0 && len[k] == len[k - 1];        // equal tuples compare equal (utils.py:80-85)
            for (int b = 0; same && b < len[k]; ++b) same = codes[o + b] == codes[off[k - 1] + b];
            rank[k] = same ? rank[k - 1] : (uint8_t)k;
            nxt[k] = (int8_t)(k - 1);                         // prepended list: k -> k-1 -> ... -> 0 (non-increasing)
            o += len[k];
        }
        head = n - 1;
        i = nth(head, n - 2);
        j = nth(head, n - 1);
    }
    MB_HD int nth(int h, int k) const {
        while (k > 0 && nxt[h] >= 0) { h = nxt[h]; --k; }
        return h;
    }
    // one row: the codes of the groups in list order
    MB_HD void compose(uint8_t *dst) const {
        for (int h = head; h >= 0; h = nxt[h])
            for (int b = 0; b < len[h]; ++b) *dst++ = codes[off[h] + b];
    }
    // the next permutation (utils.py:66-88); false after the last one
    MB_HD bool advance() {
        if (!(nxt[j] >= 0 || rank[j] < rank[head])) return false;
        const int s = (nxt[j] >= 0 && rank[i] >= rank[nxt[j]]) ? j : i;
        const int t = nxt[s];
        nxt[s] = nxt[t];
        nxt[t] = (int8_t)head;
        if (rank[t] < rank[head]) i = t;
        j = nxt[i];
        head = t;
        return true;
    }
};

// sequential form (host test build; the CUDA kernel in metis_search.cu spreads the copy of a row over a warp)
MB_HD void write_composition_rows(const MetisCompRec &rec, const uint8_t *pool, uint8_t *rows) {
    uint8_t *dst = rows + rec.row_offset;
    CompWalk cw;
    cw.init(rec, pool);
    for (uint32_t skip = 0; skip < rec.first_row; ++skip) cw.advance();       // the slice starts here
    for (uint32_t r = 0; r < rec.num_rows; ++r) {
        cw.compose(dst);
        dst += rec.stages;
        if (!cw.advance()) break;
    }
}

}  // namespace metis
