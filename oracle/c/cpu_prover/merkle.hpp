// CPU prover (oracle; TEST INFRASTRUCTURE ONLY).
//
// Mixed-height Blake2s Merkle commitment / decommitment.  Restates stwo prover/vcs/prover.rs (MerkleProver::commit,
// decommit), core/vcs/blake2_merkle.rs (hash_node = Blake2s-256(left || right || column values LE)) and the packed
// commit_on_layer of prover/backend/simd/blake2s.rs @0790eba (un-vendored).  Reference call sites: every
// tree_builder.commit(channel), /root/reference/crates/prover/src/prover.rs:59,179,298.
#pragma once
#include <algorithm>
#include <map>

#include "blake2s.hpp"
#include "buf.hpp"

namespace cpu {

struct ColRef {
    const uint32_t* p;
    int log;
};

struct MerkleTree {
    int max_log = 0;
    // layers[k]: 2^k digests as 8 word planes (plane w at [w << k, (w+1) << k)): the packed hash reads children with
    // two vector loads and one de-interleave per word
    std::vector<Col> layers;
    std::vector<ColRef> cols;  // sorted by log descending (stable)

    Hash node(int log, size_t i) const {
        Hash h;
        for (int w = 0; w < 8; w++) {
            uint32_t x = layers[log][((size_t)w << log) + i];
            memcpy(h.data() + 4 * w, &x, 4);
        }
        return h;
    }
    Hash root() const { return node(0, 0); }
};

static inline void merkle_layer(Col& out, int log, const Col* prev,
                                const std::vector<const uint32_t*>& cols) {
    const size_t n = (size_t)1 << log;
    out.alloc_uninit(8 * n);
    const int nc = (int)cols.size();
    const int words = (prev ? 16 : 0) + nc;
    const int nblk = std::max(1, (words + 15) / 16);
    const uint32_t total_bytes = 4u * words;
    if (n >= (size_t)W) {
#pragma omp parallel for schedule(static) if (n * nblk >= 4096)
        for (size_t i = 0; i < n; i += W) {
            V h[8];
            for (int k = 0; k < 8; k++) h[k] = vset1(B2S_IV[k] ^ (k == 0 ? 0x01010020u : 0));
            for (int b = 0; b < nblk; b++) {
                V m[16];
                for (int j = 0; j < 16; j++) {
                    int w = 16 * b + j;
                    if (prev && w < 16) {
                        if (w < 8) {
                            const uint32_t* pl = prev->data() + ((size_t)w << (log + 1)) + 2 * i;
                            vdeinterleave(vload(pl), vload(pl + W), m[w], m[w + 8]);
                        }
                        continue;  // words 8..15 were filled with words 0..7
                    }
                    int c = w - (prev ? 16 : 0);
                    m[j] = c < nc ? vload(cols[c] + i) : vzero();
                }
                bool last = b == nblk - 1;
                b2s_compress_v(h, m, last ? total_bytes : 64u * (b + 1), last ? 0xFFFFFFFFu : 0);
            }
            for (int k = 0; k < 8; k++) vstore(out.data() + ((size_t)k << log) + i, h[k]);
        }
    } else {
        for (size_t i = 0; i < n; i++) {
            std::vector<uint8_t> msg(4 * (size_t)words);
            size_t o = 0;
            if (prev)
                for (int side = 0; side < 2; side++)
                    for (int w = 0; w < 8; w++) {
                        uint32_t x = (*prev)[((size_t)w << (log + 1)) + 2 * i + side];
                        memcpy(msg.data() + o, &x, 4);
                        o += 4;
                    }
            for (int c = 0; c < nc; c++) {
                memcpy(msg.data() + o, &cols[c][i], 4);
                o += 4;
            }
            Hash hh = b2s_hash(msg);
            for (int w = 0; w < 8; w++) memcpy(&out[((size_t)w << log) + i], hh.data() + 4 * w, 4);
        }
    }
}

static inline MerkleTree merkle_commit(const std::vector<ColRef>& columns) {
    MerkleTree t;
    t.cols = columns;
    std::stable_sort(t.cols.begin(), t.cols.end(), [](const ColRef& a, const ColRef& b) { return a.log > b.log; });
    if (t.cols.empty()) {
        Hash e = b2s_hash(nullptr, 0);
        t.layers.resize(1);
        t.layers[0].alloc_uninit(8);
        memcpy(t.layers[0].data(), e.data(), 32);
        return t;
    }
    t.max_log = t.cols[0].log;
    t.layers.resize(t.max_log + 1);
    for (int log = t.max_log; log >= 0; log--) {
        std::vector<const uint32_t*> lc;
        for (auto& c : t.cols)
            if (c.log == log) lc.push_back(c.p);
        merkle_layer(t.layers[log], log, log == t.max_log ? nullptr : &t.layers[log + 1], lc);
    }
    return t;
}

struct Decommitment {
    std::vector<Hash> hash_witness;
    std::vector<uint32_t> column_witness;
};

// MerkleProver::decommit: queries per log size (sorted, deduplicated) -> queried values, witness
static inline void merkle_decommit(const MerkleTree& t, const std::map<int, std::vector<size_t>>& queries,
                                   std::vector<uint32_t>& queried_values, Decommitment& d) {
    std::vector<size_t> last;
    for (int log = (int)t.layers.size() - 1; log >= 0; log--) {
        std::vector<const uint32_t*> lc;
        for (auto& c : t.cols)
            if (c.log == log) lc.push_back(c.p);
        bool has_prev = log + 1 < (int)t.layers.size();
        static const std::vector<size_t> none;
        auto qi = queries.find(log);
        const std::vector<size_t>& colq = qi == queries.end() ? none : qi->second;
        size_t pi = 0, ci = 0;
        std::vector<size_t> total;
        while (pi < last.size() || ci < colq.size()) {
            size_t node = SIZE_MAX;
            if (pi < last.size()) node = std::min(node, last[pi] / 2);
            if (ci < colq.size()) node = std::min(node, colq[ci]);
            if (has_prev) {
                if (pi < last.size() && last[pi] == 2 * node)
                    pi++;
                else
                    d.hash_witness.push_back(t.node(log + 1, 2 * node));
                if (pi < last.size() && last[pi] == 2 * node + 1)
                    pi++;
                else
                    d.hash_witness.push_back(t.node(log + 1, 2 * node + 1));
            }
            bool queried = ci < colq.size() && colq[ci] == node;
            if (queried) ci++;
            for (auto p : lc) (queried ? queried_values : d.column_witness).push_back(p[node]);
            total.push_back(node);
        }
        last.swap(total);
    }
}

}  // namespace cpu
