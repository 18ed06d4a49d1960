// Device-resident prover: the host orchestration of `luminair_prover::prover::prove`
// (/root/reference/crates/prover/src/prover.rs:28-319) and of what it calls in stwo @0790eba
// (CommitmentSchemeProver / TreeBuilder, prover::prove, FriProver, MerkleProver::decommit,
// Blake2sChannel; un-vendored, Cargo.toml:21-28).  Every O(N) step is a CUDA kernel on the
// context's stream; the host keeps only the Fiat-Shamir channel, the sampled values, the query
// bookkeeping and the proof bytes.  Columns are uploaded once (row-major trace tables) and
// never leave HBM until the decommitment gathers.
//
// The reference's host side is compiled Rust; no Rust toolchain exists in this image, so this
// file is the host side in C++ above the same kernels a Rust `CudaBackend` shim would bind
// one by one.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "comm.h"
#include "ctx.h"
#include "host_util.h"
#include "kernels.cuh"
#include "merkle.cuh"

namespace lb {
namespace {

// ------------------------------------------------------------------------------------
// host field / circle helpers
// ------------------------------------------------------------------------------------
inline QM31 q_one() { return q_from_m(1); }
inline QM31 q_conj(QM31 x) { return {x.a, c_neg(x.b)}; }  // a - b u
inline bool q_is_zero(QM31 x) { return q_eq(x, q_zero()); }
inline QM31 q_double_x(QM31 x) {
    QM31 s = q_mul(x, x);
    return q_sub(q_add(s, s), q_one());
}
// SecureField::from_partial_evals: e0 + e1 i + e2 u + e3 iu
inline QM31 q_from_partial_evals(const QM31 e[4]) {
    QM31 r = e[0];
    r = q_add(r, q_mul(e[1], q_make(0, 1, 0, 0)));
    r = q_add(r, q_mul(e[2], q_make(0, 0, 1, 0)));
    r = q_add(r, q_mul(e[3], q_make(0, 0, 0, 1)));
    return r;
}

constexpr uint32_t CIRCLE_ORDER_MASK = 0x7FFFFFFFu;  // indices live in Z / 2^31

Pt host_index_to_point(uint32_t idx) {
    static Pt pow2[31];
    static bool init = false;
    if (!init) {
        Pt p = {2, 1268011823u};
        for (int j = 0; j < 31; ++j) {
            pow2[j] = p;
            p = pt_add(p, p);
        }
        init = true;
    }
    idx &= CIRCLE_ORDER_MASK;
    Pt r = {1, 0};
    for (int j = 0; j < 31; ++j)
        if ((idx >> j) & 1) r = pt_add(r, pow2[j]);
    return r;
}

struct QPt {
    QM31 x, y;
};
inline QPt qpt_add(QPt p, QPt q) {
    return {q_sub(q_mul(p.x, q.x), q_mul(p.y, q.y)), q_add(q_mul(p.x, q.y), q_mul(p.y, q.x))};
}
inline QPt qpt_lift(Pt p) { return {q_from_m(p.x), q_from_m(p.y)}; }
inline bool qpt_eq(const QPt& a, const QPt& b) { return q_eq(a.x, b.x) && q_eq(a.y, b.y); }
inline bool qpt_less(const QPt& a, const QPt& b) {
    // BTreeMap order of CirclePoint<SecureField>: x then y, each by its 4 u32 coordinates
    uint32_t ka[8] = {a.x.a.a, a.x.a.b, a.x.b.a, a.x.b.b, a.y.a.a, a.y.a.b, a.y.b.a, a.y.b.b};
    uint32_t kb[8] = {b.x.a.a, b.x.a.b, b.x.b.a, b.x.b.b, b.y.a.a, b.y.a.b, b.y.b.a, b.y.b.b};
    for (int i = 0; i < 8; ++i)
        if (ka[i] != kb[i]) return ka[i] < kb[i];
    return false;
}

inline uint32_t subgroup_gen(int log) { return (uint32_t)(((uint64_t)1 << (31 - log)) & CIRCLE_ORDER_MASK); }

// core/constraints.rs coset_vanishing for CanonicCoset(log).coset = odds(log)
QM31 coset_vanishing_q(int log, QPt p) {
    uint32_t initial = subgroup_gen(log + 1), step = subgroup_gen(log);
    QPt q = qpt_add(qpt_add(p, qpt_lift(host_index_to_point((0u - initial) & CIRCLE_ORDER_MASK))),
                    qpt_lift(host_index_to_point(step >> 1)));
    QM31 x = q.x;
    for (int i = 1; i < log; ++i) x = q_double_x(x);
    return x;
}
uint32_t coset_vanishing_m(int log, Pt p) {
    uint32_t initial = subgroup_gen(log + 1), step = subgroup_gen(log);
    Pt q = pt_add(pt_add(p, host_index_to_point((0u - initial) & CIRCLE_ORDER_MASK)), host_index_to_point(step >> 1));
    uint32_t x = q.x;
    for (int i = 1; i < log; ++i) x = m_sub(m_mul(2, m_mul(x, x)), 1);
    return x;
}
inline uint32_t bit_reverse(uint32_t i, int log) {
    uint32_t r = 0;
    for (int b = 0; b < log; ++b) r |= ((i >> b) & 1u) << (log - 1 - b);
    return r;
}

// ------------------------------------------------------------------------------------
// Blake2sChannel (core/channel/blake2s.rs; call sites prover.rs:44,177,296)
//   variant 0 "legacy": mix_u64 = raw compress(h = digest, m = [lo, hi, 0..]);
//                       draw = Blake2s(digest || counter as 32 LE bytes)      [artifact-verified]
//   variant 1 "v2":     mix_u64 = Blake2s(digest || lo || hi); draw appends one 0x00 byte
// ------------------------------------------------------------------------------------
class Channel {
   public:
    explicit Channel(int variant, std::vector<Hash32>* log) : variant_(variant), log_(log) { std::memset(digest_.b, 0, 32); }
    const Hash32& digest() const { return digest_; }
    void digest_words(uint32_t w[8]) const { std::memcpy(w, digest_.b, 32); }

    void mix_root(const Hash32& root) {
        uint8_t buf[64];
        std::memcpy(buf, digest_.b, 32);
        std::memcpy(buf + 32, root.b, 32);
        update(blake2s_hash(buf, 64));
    }
    void mix_felts(const std::vector<QM31>& felts) {
        std::vector<uint8_t> buf(32 + 16 * felts.size());
        std::memcpy(buf.data(), digest_.b, 32);
        for (size_t i = 0; i < felts.size(); ++i) {
            uint32_t w[4] = {felts[i].a.a, felts[i].a.b, felts[i].b.a, felts[i].b.b};
            std::memcpy(buf.data() + 32 + 16 * i, w, 16);
        }
        update(blake2s_hash(buf.data(), buf.size()));
    }
    void mix_u64(uint64_t v) {
        uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
        if (variant_ == 0) {
            uint32_t h[8], m[16] = {0};
            std::memcpy(h, digest_.b, 32);
            m[0] = lo;
            m[1] = hi;
            blake2s_compress(h, m, 0, 0, 0);
            Hash32 d;
            std::memcpy(d.b, h, 32);
            update(d);
        } else {
            uint8_t buf[40];
            std::memcpy(buf, digest_.b, 32);
            std::memcpy(buf + 32, &lo, 4);
            std::memcpy(buf + 36, &hi, 4);
            update(blake2s_hash(buf, 40));
        }
    }
    Hash32 draw_random_bytes() {
        uint8_t buf[65];
        std::memset(buf, 0, sizeof(buf));
        std::memcpy(buf, digest_.b, 32);
        uint32_t c = n_sent_;
        std::memcpy(buf + 32, &c, 4);
        ++n_sent_;
        return blake2s_hash(buf, variant_ == 0 ? 64 : 65);
    }
    void draw_base_felts(uint32_t out[8]) {
        for (;;) {
            Hash32 h = draw_random_bytes();
            uint32_t u[8];
            std::memcpy(u, h.b, 32);
            bool ok = true;
            for (int i = 0; i < 8; ++i) ok = ok && (u[i] < 2 * P);
            if (!ok) continue;
            for (int i = 0; i < 8; ++i) out[i] = u[i] >= P ? u[i] - P : u[i];
            return;
        }
    }
    QM31 draw_secure_felt() {
        uint32_t f[8];
        draw_base_felts(f);
        return q_make(f[0], f[1], f[2], f[3]);
    }
    std::vector<QM31> draw_secure_felts(int n) {
        std::vector<QM31> out;
        std::vector<uint32_t> pool;
        while ((int)out.size() < n) {
            if (pool.size() < 4) {
                uint32_t f[8];
                draw_base_felts(f);
                pool.insert(pool.end(), f, f + 8);
            }
            out.push_back(q_make(pool[0], pool[1], pool[2], pool[3]));
            pool.erase(pool.begin(), pool.begin() + 4);
        }
        return out;
    }

    // hand-over to / from the device-side channel of the FRI commit loop
    uint32_t n_sent() const { return n_sent_; }
    void adopt(const Hash32& d, uint32_t n_sent) {
        digest_ = d;
        n_sent_ = n_sent;
    }
    void log_digest(const Hash32& d) {
        if (log_) log_->push_back(d);
    }

   private:
    void update(const Hash32& d) {
        digest_ = d;
        n_sent_ = 0;
        if (log_) log_->push_back(d);
    }
    int variant_;
    Hash32 digest_;
    uint32_t n_sent_ = 0;
    std::vector<Hash32>* log_;
};

// ------------------------------------------------------------------------------------
// Merkle trees over mixed-height column sets (prover/vcs/prover.rs MerkleProver)
// ------------------------------------------------------------------------------------
#ifndef LB_FRI_TAIL
#define LB_FRI_TAIL 1  // the last FRI layers (<= 2^FRI_TAIL_MAX_LOG values) in one single-CTA launch (pcs_kernels.cu fri_tail_kernel)
#endif
#ifndef LB_MERKLE_SUBTREE
#define LB_MERKLE_SUBTREE 1
#endif
struct ColRef {
    const uint32_t* ptr;
    int log;
};

struct MerkleTree {
    int max_log = 0;
    bool empty = true;
    std::vector<uint32_t*> layers;  // layers[k]: 2^k digests (8 u32 each), device
    std::vector<ColRef> sorted;     // stable sort by size, descending
    Hash32 root;
    // Row-sharded tree (sharded prover): this rank holds the sub-tree over its row range - `layers` / `sorted` / `max_log`
    // describe that local sub-tree, whose root is node `rank` of level `logw` of the whole tree; the levels above it
    // (top[k], k <= logw, 2^k digests) are hashed on the host from the all-gathered sub-tree roots.  logw = 0: whole tree.
    int logw = 0, rank = 0;
    std::vector<std::vector<Hash32>> top;
    int global_max_log() const { return max_log + logw; }
};

// collects device word addresses; one gather kernel + one D2H serves the whole decommitment
struct Gatherer {
    std::vector<const uint32_t*> addrs;  // null: a word held by another rank (its value arrives through the all-reduce)
    std::vector<uint32_t> values;
    std::vector<std::pair<size_t, uint32_t>> consts;  // words known on the host (upper levels of a sharded tree)
    size_t add(const uint32_t* a) {
        addrs.push_back(a);
        return addrs.size() - 1;
    }
    size_t add_const(uint32_t v) {
        consts.push_back({addrs.size(), v});
        addrs.push_back(nullptr);
        return addrs.size() - 1;
    }
};

inline size_t add_hash(Gatherer& g, const uint32_t* h) {
    size_t first = g.add(h);
    for (int i = 1; i < 8; ++i) g.add(h ? h + i : nullptr);
    return first;
}
// word `row` (global) of a column held as row shards of 2^local_log rows per rank (logw = 0: replicated, rank 0 reports it)
inline size_t add_col_word(Gatherer& g, const uint32_t* local, int local_log, int logw, int rank, uint32_t row) {
    if (logw == 0) return g.add(rank == 0 ? local + row : nullptr);
    uint32_t owner = row >> local_log;
    return g.add((int)owner == rank ? local + (row & ((1u << local_log) - 1)) : nullptr);
}
// digest `idx` of global level `level` of a (possibly row-sharded) tree
inline size_t add_tree_hash(Gatherer& g, const MerkleTree& t, int level, uint32_t idx, int rank) {
    if (level < t.logw) {
        const Hash32& h = t.top[level][idx];
        uint32_t w[8];
        std::memcpy(w, h.b, 32);
        size_t first = g.add_const(w[0]);
        for (int i = 1; i < 8; ++i) g.add_const(w[i]);
        return first;
    }
    int local_level = level - t.logw;
    if (t.logw == 0) return add_hash(g, rank == 0 ? t.layers[local_level] + (size_t)idx * 8 : nullptr);
    uint32_t owner = idx >> local_level;
    return add_hash(g, (int)owner == t.rank ? t.layers[local_level] + (size_t)(idx & ((1u << local_level) - 1)) * 8 : nullptr);
}

struct DecommitIdx {
    std::vector<size_t> hash_witness;    // index of the first of 8 words
    std::vector<size_t> column_witness;  // word index
    std::vector<size_t> queried_values;  // word index
};

void merkle_commit(lb_ctx* ctx, Arena& arena, const std::vector<ColRef>& cols, MerkleTree& t, bool fetch_root = true) {
    if (cols.empty()) {
        t.empty = true;
        t.max_log = 0;
        t.root = blake2s_hash(nullptr, 0);
        return;
    }
    t.empty = false;
    t.sorted = cols;
    std::stable_sort(t.sorted.begin(), t.sorted.end(), [](const ColRef& a, const ColRef& b) { return a.log > b.log; });
    t.max_log = t.sorted[0].log;
    t.layers.assign(t.max_log + 1, nullptr);
    // one pointer table for all layers
    std::vector<const uint32_t*> table;
    std::vector<int> first(t.max_log + 2, 0), count(t.max_log + 1, 0);
    for (int log = t.max_log; log >= 0; --log) {
        first[log] = (int)table.size();
        for (const ColRef& c : t.sorted)
            if (c.log == log) table.push_back(c.ptr);
        count[log] = (int)table.size() - first[log];
    }
    bool small = true;
    for (int log = 0; log <= t.max_log; ++log) small = small && count[log] <= MERKLE_SMALL_COLS;
    const uint32_t** d_table = small ? nullptr : arena.upload(table);
    // layers below `fused_from` hold no columns and fit one CTA: they are hashed by a single launch
    int fused_from = 0;
    while (fused_from < t.max_log && fused_from < MERKLE_TOP_MAX_LOG && count[fused_from] == 0) ++fused_from;
    const uint32_t* prev = nullptr;
    {
        // one allocation for all layers: layer k (2^k digests of 8 words) at word offset 8 * (2^k - 1)
        uint32_t* all = arena.alloc<uint32_t>((size_t)16 << t.max_log);
        for (int log = t.max_log; log >= 0; --log) t.layers[log] = all + 8 * (((size_t)1 << log) - 1);
    }
    for (int log = t.max_log; log >= fused_from; --log) {
        if (LB_MERKLE_SUBTREE && small && log <= MERKLE_SUBTREE_MAX_LOG && log >= fused_from + 3) {
            // narrow middle of the tree: this layer and the column-less layers above it, down to the layer the fused top
            // starts from, in one launch (only when that saves at least two launches: small trees are faster layer by layer)
            int depth = 1;
            while (depth < MERKLE_SUBTREE_MAX_DEPTH && log - depth >= fused_from && count[log - depth] == 0) ++depth;
            if (depth >= 2) {
                MerkleSubtreeArgs a{};
                for (int d = 0; d < depth; ++d) a.layers[d] = t.layers[log - d];
                a.prev = prev;
                for (int k = 0; k < count[log]; ++k) a.cols.p[k] = table[first[log] + k];
                a.n_cols = count[log];
                a.log_top = log;
                a.depth = depth;
                ck(merkle_commit_subtree(a, ctx->stream), "merkle subtree");
                log -= depth - 1;
                prev = t.layers[log];
                continue;
            }
        }
        if (small) {
            MerkleColsArg ca{};
            for (int k = 0; k < count[log]; ++k) ca.p[k] = table[first[log] + k];
            ck(merkle_commit_layer_small(t.layers[log], prev, ca, count[log], log, ctx->stream), "merkle layer");
        } else {
            ck(merkle_commit_layer(t.layers[log], prev, d_table + first[log], count[log], log, ctx->stream), "merkle layer");
        }
        prev = t.layers[log];
    }
    if (fused_from >= 1) {
        MerkleTopArgs a{};
        for (int k = 0; k <= fused_from; ++k) a.layers[k] = t.layers[k];
        a.from_log = fused_from;
        ck(merkle_commit_top(a, ctx->stream), "merkle top");
    }
    if (!fetch_root) return;  // the caller reads t.layers[0] on the device and fetches the root later
    ck(cudaMemcpyAsync(t.root.b, t.layers[0], 32, cudaMemcpyDeviceToHost, ctx->stream), "root d2h");
    ck(cudaStreamSynchronize(ctx->stream), "root sync");
}

// MerkleProver::decommit: emits, in proof order, what must be gathered
void merkle_decommit_plan(const MerkleTree& t, const std::map<int, std::vector<uint32_t>>& queries_per_log, Gatherer& g,
                          DecommitIdx& out, int rank = 0) {
    std::vector<uint32_t> last;
    const int gmax = t.empty ? 0 : t.global_max_log();
    for (int log = gmax; log >= 0; --log) {
        const bool prev_hashes = !t.empty && log + 1 <= gmax;
        std::vector<uint32_t> col_q;
        auto it = queries_per_log.find(log);
        if (it != queries_per_log.end()) col_q = it->second;
        const std::vector<uint32_t>& prev_q = last;
        size_t pi = 0, ci = 0;
        std::vector<uint32_t> total;
        while (pi < prev_q.size() || ci < col_q.size()) {
            uint32_t node = 0xFFFFFFFFu;
            if (pi < prev_q.size()) node = std::min(node, prev_q[pi] / 2);
            if (ci < col_q.size()) node = std::min(node, col_q[ci]);
            if (prev_hashes) {
                if (pi < prev_q.size() && prev_q[pi] == 2 * node)
                    ++pi;
                else
                    out.hash_witness.push_back(add_tree_hash(g, t, log + 1, 2 * node, rank));
                if (pi < prev_q.size() && prev_q[pi] == 2 * node + 1)
                    ++pi;
                else
                    out.hash_witness.push_back(add_tree_hash(g, t, log + 1, 2 * node + 1, rank));
            }
            bool queried = ci < col_q.size() && col_q[ci] == node;
            if (queried) ++ci;
            for (const ColRef& c : t.sorted) {
                if (c.log + t.logw != log) continue;
                size_t idx = add_col_word(g, c.ptr, c.log, t.logw, rank, node);
                (queried ? out.queried_values : out.column_witness).push_back(idx);
            }
            total.push_back(node);
        }
        last.swap(total);
    }
}


// ------------------------------------------------------------------------------------
// proof structures (indices into the Gatherer until the final D2H) + bincode 1.3 writer
// layout: crates/prover/src/lib.rs:15-32, crates/air/src/lib.rs:29-48,189-207 + stwo serde derives
// ------------------------------------------------------------------------------------
struct FriLayerOut {
    std::vector<size_t> witness;  // per QM31: index of coordinate 0..3 are consecutive gather slots
    DecommitIdx decommit;
    Hash32 commitment;
};

struct Writer {
    std::vector<uint8_t>& o;
    void u8(uint8_t v) { o.push_back(v); }
    void u32(uint32_t v) { raw(&v, 4); }
    void u64(uint64_t v) { raw(&v, 8); }
    void raw(const void* p, size_t n) { o.insert(o.end(), (const uint8_t*)p, (const uint8_t*)p + n); }
    void qm31(QM31 q) {
        uint32_t w[4] = {q.a.a, q.a.b, q.b.a, q.b.b};
        raw(w, 16);
    }
    void hash(const Hash32& h) { raw(h.b, 32); }
};

void write_decommit(Writer& w, const DecommitIdx& d, const Gatherer& g) {
    w.u64(d.hash_witness.size());
    for (size_t i : d.hash_witness) w.raw(&g.values[i], 32);
    w.u64(d.column_witness.size());
    for (size_t i : d.column_witness) w.u32(g.values[i]);
}
void write_fri_layer(Writer& w, const FriLayerOut& l, const Gatherer& g) {
    w.u64(l.witness.size());
    for (size_t i : l.witness) w.raw(&g.values[i], 16);
    write_decommit(w, l.decommit, g);
    w.hash(l.commitment);
}

// ------------------------------------------------------------------------------------
// queries (core/queries.rs, core/fri.rs)
// ------------------------------------------------------------------------------------
std::vector<uint32_t> generate_queries(Channel& ch, int log_domain_size, size_t n_queries) {
    std::vector<uint32_t> qs;
    size_t cnt = 0;
    uint32_t mask = (log_domain_size >= 32) ? 0xFFFFFFFFu : ((1u << log_domain_size) - 1);
    while (cnt < n_queries) {
        Hash32 rb = ch.draw_random_bytes();
        uint32_t w[8];
        std::memcpy(w, rb.b, 32);
        for (int i = 0; i < 8 && cnt < n_queries; ++i) {
            qs.push_back(w[i] & mask);
            ++cnt;
        }
    }
    std::sort(qs.begin(), qs.end());
    qs.erase(std::unique(qs.begin(), qs.end()), qs.end());
    return qs;
}

std::vector<uint32_t> fold_queries(const std::vector<uint32_t>& pos, int n_folds) {
    std::vector<uint32_t> out;
    for (uint32_t q : pos) {
        uint32_t f = q >> n_folds;
        if (out.empty() || out.back() != f) out.push_back(f);
    }
    return out;
}

// compute_decommitment_positions_and_witness_evals with fold_step = 1
void fri_positions_and_witness(uint32_t* const coords[4], const std::vector<uint32_t>& queries, Gatherer& g,
                               std::vector<uint32_t>& positions, std::vector<size_t>& witness, int local_log = 0, int logw = 0,
                               int rank = 0) {
    size_t i = 0;
    while (i < queries.size()) {
        size_t j = i;
        while (j < queries.size() && (queries[j] >> 1) == (queries[i] >> 1)) ++j;
        uint32_t start = (queries[i] >> 1) << 1;
        size_t k = i;
        for (uint32_t pos = start; pos < start + 2; ++pos) {
            positions.push_back(pos);
            if (k < j && queries[k] == pos) {
                ++k;
                continue;
            }
            size_t first = add_col_word(g, coords[0], local_log, logw, rank, pos);
            for (int c = 1; c < 4; ++c) add_col_word(g, coords[c], local_log, logw, rank, pos);
            witness.push_back(first);
        }
        i = j;
    }
}

// ------------------------------------------------------------------------------------
// host evaluators over the component templates (air.cuh)
// ------------------------------------------------------------------------------------
struct InfoEval : LogupMixin<InfoEval, FQ, FQ> {
    typedef FQ F;
    typedef FQ EF;
    int n_main = 0, n_inter = 0, n_constraints = 0, n_pre = 0;
    QM31 cumsum_shift = q_zero();
    FQ constant(uint32_t c) const { return {q_from_m(c)}; }
    FQ get_preprocessed_column(int k) {
        n_pre = std::max(n_pre, k + 1);
        return {q_zero()};
    }
    FQ next_trace_mask() {
        ++n_main;
        return {q_zero()};
    }
    FQ next_ext_mask_cur() {
        n_inter += 4;
        return {q_zero()};
    }
    void next_ext_mask_prev_cur(FQ& prev, FQ& cur) {
        n_inter += 4;
        prev = cur = FQ{q_zero()};
    }
    void add_constraint(FQ) { ++n_constraints; }
    void add_constraint_ef(FQ) { ++n_constraints; }
};

// constraint-framework PointEvaluator: F = EF = SecureField, mask values from the sampled values
struct PointEval : LogupMixin<PointEval, FQ, FQ> {
    typedef FQ F;
    typedef FQ EF;
    const std::vector<std::vector<QM31>>* main;   // sampled_values[1] (global column index)
    const std::vector<std::vector<QM31>>* inter;  // sampled_values[2]
    const std::vector<std::vector<QM31>>* pre;    // sampled_values[0]
    int pre_idx[2];                               // this component's preprocessed columns (global index)
    size_t mc, ic;
    QM31 denom_inverse, random_coeff;
    QM31* acc;
    QM31 cumsum_shift;

    FQ constant(uint32_t c) const { return {q_from_m(c)}; }
    FQ get_preprocessed_column(int k) const { return {(*pre)[pre_idx[k]][0]}; }
    FQ next_trace_mask() { return {(*main)[mc++][0]}; }
    FQ ext_at(size_t k) const {
        QM31 e[4] = {(*inter)[ic][k], (*inter)[ic + 1][k], (*inter)[ic + 2][k], (*inter)[ic + 3][k]};
        return {q_from_partial_evals(e)};
    }
    FQ next_ext_mask_cur() {
        FQ v = ext_at(0);
        ic += 4;
        return v;
    }
    void next_ext_mask_prev_cur(FQ& prev, FQ& cur) {
        prev = ext_at(0);
        cur = ext_at(1);
        ic += 4;
    }
    void add_constraint(FQ c) { *acc = q_add(q_mul(*acc, random_coeff), q_mul(denom_inverse, c.v)); }
    void add_constraint_ef(FQ c) { add_constraint(c); }
};

// ------------------------------------------------------------------------------------
// shared launch helpers (used by prove_impl and by the trait-level C ABI)
// ------------------------------------------------------------------------------------
struct HostBatch {  // ColumnSampleBatch: one sample point and the (column, value) pairs sampled there
    QPt pt;
    std::vector<std::pair<int, QM31>> cols;
    // column-sharded accumulation: this rank holds columns [col_offset, col_offset + cols.size()) of a
    // batch of n_cols_global columns; the partial sums of all ranks add up to the full quotient
    int col_offset = 0;
    int n_cols_global = -1;  // -1: cols.size()
};

// PolyOps::eval_at_point for several sets of equally sized coefficient columns (one set per (size, point) pair): every
// argument table of every set goes to the device in ONE staged upload, then the kernels are queued back to back
struct EvalJob {
    int log;
    QPt pt;
    std::vector<const uint32_t*> cols;
    QM31* d_out;  // cols.size() results
};
void launch_eval_jobs(lb_ctx* ctx, Arena& arena, const std::vector<EvalJob>& jobs) {
    if (jobs.empty()) return;
    // one byte blob: per job the fold factors [y, x, pi(x), ...] then the column pointer table
    std::vector<uint8_t> blob;
    std::vector<size_t> off_map(jobs.size()), off_cols(jobs.size());
    auto append = [&](const void* p, size_t n) {
        size_t at = (blob.size() + 15) & ~(size_t)15;
        blob.resize(at + n);
        std::memcpy(blob.data() + at, p, n);
        return at;
    };
    for (size_t j = 0; j < jobs.size(); ++j) {
        std::vector<QM31> mappings;
        mappings.push_back(jobs[j].pt.y);
        QM31 x = jobs[j].pt.x;
        for (int i = 1; i < jobs[j].log; ++i) {
            mappings.push_back(x);
            x = q_double_x(x);
        }
        off_map[j] = append(mappings.data(), mappings.size() * sizeof(QM31));
        off_cols[j] = append(jobs[j].cols.data(), jobs[j].cols.size() * sizeof(const uint32_t*));
    }
    uint8_t* d_blob = arena.upload(blob);
    for (size_t j = 0; j < jobs.size(); ++j) {
        const EvalJob& job = jobs[j];
        int m = std::min(job.log, EVAL_CHUNK_LOG);
        QM31* d_basis = arena.alloc<QM31>((size_t)1 << m);
        QM31* d_part = arena.alloc<QM31>(job.cols.size() << (job.log - m));
        ck(eval_at_point((const uint32_t* const*)(d_blob + off_cols[j]), (int)job.cols.size(), job.log, (const QM31*)(d_blob + off_map[j]),
                         d_basis, d_part, job.d_out, ctx->stream),
           "eval_at_point");
    }
}
void launch_eval_at_point(lb_ctx* ctx, Arena& arena, const std::vector<const uint32_t*>& cols, int log, const QPt& pt,
                          QM31* d_out) {
    launch_eval_jobs(ctx, arena, {EvalJob{log, pt, cols, d_out}});
}

// QuotientOps::accumulate_quotients: quotient_constants (core/pcs/quotients.rs) on the host,
// the row loop on the device
void launch_quotients(lb_ctx* ctx, Arena& arena, int lg, const std::vector<const uint32_t*>& colptrs,
                      std::vector<HostBatch>& batches, QM31 rc_q, uint32_t* const out[4], uint32_t row0 = 0,
                      uint32_t n_rows = 0) {
    std::sort(batches.begin(), batches.end(), [](const HostBatch& a, const HostBatch& b) { return qpt_less(a.pt, b.pt); });
    if (batches.empty() || batches.size() > (size_t)MAX_QUOTIENT_BATCHES) fail(LB_ERR_BAD_ARG, "quotients: bad number of sample batches");
    QuotientParams qp{};
    qp.n_batches = (int)batches.size();
    std::vector<QuotientEntry> entries;
    for (size_t bi = 0; bi < batches.size(); ++bi) {
        const HostBatch& b = batches[bi];
        QuotientBatch& o = qp.b[bi];
        o.prx = b.pt.x.a;
        o.pix = b.pt.x.b;
        o.pry = b.pt.y.a;
        o.piy = b.pt.y.b;
        o.first = (int)entries.size();
        o.count = (int)b.cols.size();
        o.sum_a = q_zero();
        o.sum_b = q_zero();
        QM31 alpha = q_pow(rc_q, (uint64_t)b.col_offset);
        QM31 cc = q_sub(q_conj(b.pt.y), b.pt.y);
        for (auto& cv : b.cols) {
            if (cv.first < 0 || cv.first >= (int)colptrs.size()) fail(LB_ERR_BAD_ARG, "quotients: column index out of range");
            alpha = q_mul(alpha, rc_q);
            QM31 a = q_sub(q_conj(cv.second), cv.second);
            QM31 bcoef = q_sub(q_mul(cv.second, cc), q_mul(a, b.pt.y));
            o.sum_a = q_add(o.sum_a, q_mul(alpha, a));
            o.sum_b = q_add(o.sum_b, q_mul(alpha, bcoef));
            QuotientEntry e{};
            e.c = q_mul(alpha, cc);
            e.col = cv.first;
            entries.push_back(e);
        }
        o.rc_pow = q_pow(rc_q, b.n_cols_global >= 0 ? (uint64_t)b.n_cols_global : (uint64_t)b.cols.size());
    }
    const uint32_t** d_cols = arena.upload(colptrs);
    QuotientEntry* d_entries = arena.upload(entries);
    {
        int r = lb_twiddles_ensure(ctx, lg);  // domain points are read from the twiddle tables
        if (r) fail(r, ctx->err);
    }
    ck(accumulate_quotients(out, d_cols, d_entries, qp, &ctx->tw, lg, ctx->stream, row0, n_rows), "accumulate quotients");
}

// ------------------------------------------------------------------------------------
// commitment scheme state
// ------------------------------------------------------------------------------------
struct PolyCol {
    uint32_t* coeffs;  // 2^log (sharded prover: valid on the owner rank only)
    uint32_t* lde;     // 2^(log + blowup), filled by commit (sharded prover: this rank's row shard, 2^(log + blowup - logw))
    int log;
    int owner = 0;     // rank that interpolates / extends / samples this column
};
// columns pushed together (same size, contiguous coefficients): the unit of the batched transforms and of the ownership split
struct ColRun {
    size_t first;
    int n, log;
};
struct CommitTree {
    std::vector<PolyCol> cols;
    std::vector<ColRun> runs;
    MerkleTree merkle;
};

// ---- sharded prover plumbing (one rank per GPU; world = 1: everything below degenerates to the single-GPU path) -------
struct Shard {
    lb_comm* comm = nullptr;
    int rank = 0, world = 1, logw = 0;
    int next_start = 0;  // rotating first rank of the ownership split, so that remainders spread over the ranks
    bool on() const { return world > 1; }
};
inline void nck(ncclResult_t r, const char* what) {
    if (r != ncclSuccess) fail(LB_ERR_NCCL, std::string(what) + ": " + nccl_api().GetErrorString(r));
}
// contiguous, balanced split of a run of n columns over the ranks: out[k] = owner of column k
std::vector<int> split_run(Shard& sh, int n) {
    std::vector<int> owner(n, 0);
    if (!sh.on()) return owner;
    int base = n / sh.world, extra = n % sh.world, k = 0;
    for (int i = 0; i < sh.world; ++i) {
        int r = (sh.next_start + i) % sh.world;
        int cnt = base + (i < extra ? 1 : 0);
        for (int c = 0; c < cnt; ++c) owner[k++] = r;
    }
    sh.next_start = (sh.next_start + extra) % sh.world;
    return owner;
}
// [a, b) = this rank's columns of a run (they are contiguous by construction)
inline void own_range(const CommitTree& t, const ColRun& run, int rank, int& a, int& b) {
    a = b = 0;
    bool found = false;
    for (int k = 0; k < run.n; ++k)
        if (t.cols[run.first + k].owner == rank) {
            if (!found) a = k;
            b = k + 1;
            found = true;
        }
}
struct Xfer {
    const uint32_t* src;
    uint32_t* dst;
    size_t n;
    int peer;
    bool send;
};
// one grouped exchange; a transfer whose peer is this rank is a device copy.  Both sides of a pair list their transfers in
// the same (column) order, which is the order NCCL matches sends with receives.
void run_exchange(lb_ctx* ctx, Shard& sh, const std::vector<Xfer>& xs) {
    NcclApi& api = nccl_api();
    static const bool trace = getenv("LB_SHARD_TRACE") != nullptr;  // diagnostics: time every exchange (adds two syncs)
    std::chrono::steady_clock::time_point t0;
    unsigned long long sent0 = sh.comm->bytes_sent;
    if (trace) {
        cudaStreamSynchronize(ctx->stream);
        t0 = std::chrono::steady_clock::now();
    }
    struct TraceEnd {
        bool on;
        lb_ctx* ctx;
        Shard& sh;
        std::chrono::steady_clock::time_point& t0;
        unsigned long long& sent0;
        size_t n;
        ~TraceEnd() {
            if (!on) return;
            cudaStreamSynchronize(ctx->stream);
            double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            unsigned long long b = sh.comm->bytes_sent - sent0;
            fprintf(stderr, "[lb shard %d] exchange: %zu transfers, %.1f MB sent, %.3f ms (%.0f GB/s out)\n", sh.rank, n, b / 1e6, ms,
                    ms > 0 ? b / ms / 1e6 : 0.0);
        }
    } trace_end{trace, ctx, sh, t0, sent0, xs.size()};
    nck(api.GroupStart(), "ncclGroupStart");
    for (const Xfer& x : xs) {
        if (x.peer == sh.rank) {
            if (x.send) ck(cudaMemcpyAsync(x.dst, x.src, x.n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream), "self copy");
            continue;
        }
        if (x.send) {
            nck(api.Send(x.src, x.n, ncclUint32, x.peer, sh.comm->comm, ctx->stream), "ncclSend");
            sh.comm->bytes_sent += x.n * 4;
        } else {
            nck(api.Recv(x.dst, x.n, ncclUint32, x.peer, sh.comm->comm, ctx->stream), "ncclRecv");
            sh.comm->bytes_received += x.n * 4;
        }
    }
    nck(api.GroupEnd(), "ncclGroupEnd");
    sh.comm->n_collectives++;
}
// stream-ordered barrier over the ranks: when it completes on this stream, every rank's stream has passed the same point
void shard_barrier(lb_ctx* ctx, Shard& sh, uint32_t* d_word) {
    nck(nccl_api().AllReduce(d_word, d_word, 1, ncclUint32, ncclSum, sh.comm->comm, ctx->stream), "ncclAllReduce(barrier)");
    sh.comm->n_collectives++;
}
// (collective) make the symmetric heap at least `words` large and mapped on every rank; false: IPC is not usable here
bool ensure_symmetric(lb_ctx* ctx, Arena& arena, Shard& sh, size_t words) {
    lb_comm* c = sh.comm;
    if (c->ipc == 0) return false;
    if (c->ipc < 0 && getenv("LB_SHARD_IPC") && atoi(getenv("LB_SHARD_IPC")) == 0) {
        c->ipc = 0;
        return false;
    }
    if (c->sym_words >= words) return true;
    NcclApi& api = nccl_api();
    uint32_t* d_word = arena.alloc<uint32_t>(1);
    ck(cudaMemsetAsync(d_word, 0, 4, ctx->stream), "memset");
    shard_barrier(ctx, sh, d_word);  // nobody still reads or writes the old heap
    ck(cudaStreamSynchronize(ctx->stream), "sym sync");
    release_symmetric(c);
    size_t want = words + words / 4 + 4096;
    int ok = cudaMalloc(&c->sym_base, want * sizeof(uint32_t)) == cudaSuccess;
    cudaIpcMemHandle_t mine;
    std::memset(&mine, 0, sizeof(mine));
    if (ok) ok = cudaIpcGetMemHandle(&mine, c->sym_base) == cudaSuccess;
    // all-gather the handles (+ an ok flag per rank)
    struct Slot {
        cudaIpcMemHandle_t h;
        int ok;
        int pad[15];
    };
    static_assert(sizeof(Slot) == 128, "slot size");
    std::vector<Slot> all(sh.world);
    Slot me{};
    me.h = mine;
    me.ok = ok;
    Slot* d_all = arena.alloc<Slot>(sh.world);
    ck(cudaMemcpyAsync(d_all + sh.rank, &me, sizeof(Slot), cudaMemcpyHostToDevice, ctx->stream), "handle h2d");
    nck(api.AllGather(d_all + sh.rank, d_all, sizeof(Slot), ncclChar, c->comm, ctx->stream), "ncclAllGather(ipc handles)");
    ck(cudaMemcpyAsync(all.data(), d_all, sizeof(Slot) * sh.world, cudaMemcpyDeviceToHost, ctx->stream), "handle d2h");
    ck(cudaStreamSynchronize(ctx->stream), "handle sync");
    for (auto& sl : all) ok = ok && sl.ok;
    c->sym_peer.assign(sh.world, nullptr);
    if (ok) {
        for (int r = 0; r < sh.world && ok; ++r) {
            if (r == sh.rank) {
                c->sym_peer[r] = c->sym_base;
                continue;
            }
            void* ptr = nullptr;
            ok = cudaIpcOpenMemHandle(&ptr, all[r].h, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
            c->sym_peer[r] = (uint32_t*)ptr;
        }
        cudaGetLastError();  // clear a refused mapping (same-process ranks, no peer access)
    }
    // agree: one refusal anywhere switches every rank to the NCCL exchange
    uint32_t flag = ok ? 0u : 1u;
    ck(cudaMemcpyAsync(d_word, &flag, 4, cudaMemcpyHostToDevice, ctx->stream), "flag h2d");
    shard_barrier(ctx, sh, d_word);
    ck(cudaMemcpyAsync(&flag, d_word, 4, cudaMemcpyDeviceToHost, ctx->stream), "flag d2h");
    ck(cudaStreamSynchronize(ctx->stream), "flag sync");
    if (flag != 0) {
        release_symmetric(c);
        c->ipc = 0;
        return false;
    }
    c->sym_words = want;
    c->ipc = 1;
    return true;
}
inline uint32_t* sym_alloc(lb_comm* c, size_t words) {
    size_t off = (c->sym_used + 63) & ~(size_t)63;
    if (off + words > c->sym_words) fail(LB_ERR_OOM, "sharded prove: symmetric heap too small (internal size estimate)");
    c->sym_used = off + words;
    return c->sym_base + off;
}
inline uint32_t* peer_ptr(const lb_comm* c, int r, const uint32_t* local) { return c->sym_peer[r] + (local - c->sym_base); }

// root of a row-sharded tree: all-gather the sub-tree roots, hash the log2(world) levels above them on the host
void finish_sharded_root(lb_ctx* ctx, Arena& arena, Shard& sh, MerkleTree& t) {
    t.logw = sh.logw;
    t.rank = sh.rank;
    NcclApi& api = nccl_api();
    uint32_t* d_roots = arena.alloc<uint32_t>(8 * (size_t)sh.world);
    nck(api.AllGather(t.layers[0], d_roots, 8, ncclUint32, sh.comm->comm, ctx->stream), "ncclAllGather(roots)");
    sh.comm->n_collectives++;
    sh.comm->bytes_sent += 32;
    sh.comm->bytes_received += 32 * (size_t)(sh.world - 1);
    std::vector<Hash32> lvl(sh.world);
    ck(cudaMemcpyAsync(lvl.data(), d_roots, 32 * (size_t)sh.world, cudaMemcpyDeviceToHost, ctx->stream), "roots d2h");
    ck(cudaStreamSynchronize(ctx->stream), "roots sync");
    t.top.assign(sh.logw + 1, {});
    t.top[sh.logw] = lvl;
    for (int k = sh.logw - 1; k >= 0; --k) {
        t.top[k].resize((size_t)1 << k);
        for (size_t i = 0; i < t.top[k].size(); ++i) {
            uint8_t buf[64];
            std::memcpy(buf, t.top[k + 1][2 * i].b, 32);
            std::memcpy(buf + 32, t.top[k + 1][2 * i + 1].b, 32);
            t.top[k][i] = blake2s_hash(buf, 64);
        }
    }
    t.root = t.top[0][0];
}

struct Component {
    int kind, slot, log;
    size_t main_loc, inter_loc;  // first column in tree 1 / tree 2
    uint32_t* main_evals;        // trace values (n_main x 2^log), kept for the LogUp pass
    QM31 claimed_sum;
    int pre_idx[2] = {-1, -1};   // preprocessed columns read by a lookup-table component (index into tree 0)
    int eval_log = 0;            // max_constraint_log_degree_bound
    uint32_t* inter_prev = nullptr;  // sharded prover: row shard of the [-1]-shifted last LogUp column (4 coordinate columns)
};

// one preprocessed column as committed in tree 0
struct PreCol {
    int lut, col_index, log;
    uint32_t* evals;  // values on CanonicCoset(log) as the caller generated them (kept for the LogUp pass)
};

int kind_of_slot(int slot, int n_slots, int air_era) {
    // LuminairClaim field order, crates/air/src/lib.rs:30-48 (17 slots)
    if (slot == 0) return COMP_ADD;
    if (slot == 1) return air_era == 1 ? COMP_MUL_ARTIFACT : COMP_MUL;
    if (n_slots != 17) return -1;
    static const int kinds[17] = {COMP_ADD,        COMP_MUL,         COMP_RECIP,     COMP_SIN,         COMP_SIN_LOOKUP,
                                  COMP_SUM_REDUCE, COMP_MAX_REDUCE,  COMP_SQRT,      COMP_REM,         COMP_EXP2,
                                  COMP_EXP2_LOOKUP, COMP_LOG2,       COMP_LOG2_LOOKUP, COMP_LESS_THAN, COMP_RANGE_CHECK_LOOKUP,
                                  COMP_INPUTS,     COMP_CONTIGUOUS};
    return (slot >= 0 && slot < 17) ? kinds[slot] : -1;
}

const uint2* inv_y_twiddles(const Twiddles& tw, int domain_log) { return tw.inv + tw.y_off + ((size_t)1 << (domain_log - 1)); }
const uint2* inv_x_twiddles(const Twiddles& tw, int line_log) { return tw.inv + ((size_t)1 << (line_log - 1)); }

struct StageTimer {
    lb_ctx* ctx;
    std::chrono::steady_clock::time_point t0;
    explicit StageTimer(lb_ctx* c) : ctx(c), t0(std::chrono::steady_clock::now()) {}
    void lap() {
        cudaStreamSynchronize(ctx->stream);
        auto t1 = std::chrono::steady_clock::now();
        ctx->stage_ms.push_back(std::chrono::duration<float, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

// Grind: the smallest-index nonce found by the device search whose mix leaves pow_bits trailing zero bits (stwo
// prover/backend/simd/grind.rs semantics: any valid nonce verifies; the search order makes it the lowest one).
uint64_t grind_nonce(const Channel& channel, const lb_prove_config& cfg, Arena& arena, cudaStream_t st) {
    unsigned long long* d_found = arena.alloc<unsigned long long>(1);
    uint32_t dg[8];
    channel.digest_words(dg);
    uint64_t base = 0;
    // a nonce works with probability 2^-pow_bits: the first launch tries 2^(pow_bits + 7) of them (it misses with
    // probability e^-128), later ones grow 16-fold up to 2^24
    uint64_t chunk = (uint64_t)1 << std::min<uint32_t>(24, cfg.pow_bits + 7);
    for (;;) {
        ck(cudaMemsetAsync(d_found, 0xFF, 8, st), "grind memset");
        ck(grind_range(dg, cfg.channel_variant, cfg.pow_bits, base, chunk, d_found, st), "grind");
        unsigned long long found;
        ck(cudaMemcpyAsync(&found, d_found, 8, cudaMemcpyDeviceToHost, st), "grind d2h");
        ck(cudaStreamSynchronize(st), "grind sync");
        if (found != ~0ull) return found;
        base += chunk;
        chunk = std::min<uint64_t>(chunk << 4, (uint64_t)1 << 24);
    }
}

// FriProver::commit_last_layer: the last line evaluation (4 coordinate columns of 2^line_log values, bit-reversed order as
// stored on the device) interpolated on the host over LineDomain(half_odds(line_log)); the coefficients above the degree
// bound must vanish, the ones below are the proof's last_layer_poly.
std::vector<QM31> interpolate_last_layer(const std::vector<uint32_t>& hv, int line_log, uint32_t log_degree_bound) {
    const size_t n = (size_t)1 << line_log;
    std::vector<QM31> vals(n);
    for (size_t i = 0; i < n; ++i) {
        size_t s = bit_reverse((uint32_t)i, line_log);  // natural order
        vals[i] = q_make(hv[s], hv[n + s], hv[2 * n + s], hv[3 * n + s]);
    }
    uint32_t d_init = subgroup_gen(line_log + 2), d_step = subgroup_gen(line_log);
    size_t size = n;
    while (size > 1) {
        for (size_t start = 0; start < n; start += size)
            for (size_t i = 0; i < size / 2; ++i) {
                uint32_t x = host_index_to_point(d_init + d_step * (uint32_t)i).x;
                uint32_t xinv = m_inv(x);
                QM31 l = vals[start + i], r = vals[start + size / 2 + i];
                vals[start + i] = q_add(l, r);
                vals[start + size / 2 + i] = q_mul_m(q_sub(l, r), xinv);
            }
        d_init = (d_init * 2) & CIRCLE_ORDER_MASK;
        d_step = (d_step * 2) & CIRCLE_ORDER_MASK;
        size /= 2;
    }
    uint32_t inv_n = m_inv((uint32_t)(n % P));
    std::vector<QM31> coeffs(n);
    for (size_t i = 0; i < n; ++i) coeffs[i] = q_mul_m(vals[bit_reverse((uint32_t)i, line_log)], inv_n);
    size_t bound = (size_t)1 << log_degree_bound;
    for (size_t i = bound; i < n; ++i)
        if (!q_is_zero(coeffs[i])) fail(LB_ERR_CONSTRAINTS, "fri: invalid last-layer degree (ConstraintsNotSatisfied)");
    return std::vector<QM31>(coeffs.begin(), coeffs.begin() + bound);
}

using SampledValues = std::vector<std::vector<std::vector<QM31>>>;  // [tree][column][sample]

// The check stwo::prover::prove ends with: the composition polynomial at the OODS point, rebuilt from its four
// coordinate samples, equals the constraints evaluated on the sampled mask values.
void check_oods(const std::vector<Component>& comps, const SampledValues& sampled, const Relations& rels, QM31 random_coeff,
                QPt oods) {
    QM31 e[4] = {sampled[3][0][0], sampled[3][1][0], sampled[3][2][0], sampled[3][3][0]};
    QM31 composition_oods = q_from_partial_evals(e);
    QM31 accv = q_zero();
    for (const Component& c : comps) {
        PointEval pe;
        pe.main = &sampled[1];
        pe.inter = &sampled[2];
        pe.pre = &sampled[0];
        pe.pre_idx[0] = c.pre_idx[0];
        pe.pre_idx[1] = c.pre_idx[1];
        pe.mc = c.main_loc;
        pe.ic = c.inter_loc;
        pe.random_coeff = random_coeff;
        pe.denom_inverse = q_inv(coset_vanishing_q(c.log, oods));
        pe.acc = &accv;
        pe.cumsum_shift = q_mul_m(c.claimed_sum, m_inv((uint32_t)(((uint64_t)1 << c.log) % P)));
        eval_component(c.kind, pe, rels);
    }
    if (!q_eq(composition_oods, accv)) fail(LB_ERR_CONSTRAINTS, "ConstraintsNotSatisfied");
}

// Everything LuminairProof holds (crates/air/src/lib.rs:21-27), by reference; `Gatherer` carries the queried words.
struct ProofParts {
    const std::vector<int>& claim;
    const std::vector<Component>& comps;
    const std::vector<CommitTree>& trees;
    const SampledValues& sampled;
    const std::vector<DecommitIdx>& tree_dec;
    uint64_t nonce;
    const FriLayerOut& first_layer;
    const std::vector<FriLayerOut>& inner_layers;
    const std::vector<QM31>& last_layer_poly;
};

// bincode 1.x (fixed-width little-endian integers, u64 lengths) of LuminairProof { claim, interaction_claim, proof }
void write_proof(std::vector<uint8_t>& out, const lb_prove_config& cfg, const ProofParts& p, const Gatherer& g) {
    out.clear();
    Writer w{out};
    for (int s = 0; s < cfg.n_slots; ++s) {
        if (p.claim[s] < 0)
            w.u8(0);
        else {
            w.u8(1);
            w.u32((uint32_t)p.claim[s]);
        }
    }
    {
        std::vector<const Component*> slot_comp(cfg.n_slots, nullptr);
        for (const Component& c : p.comps) slot_comp[c.slot] = &c;
        for (int s = 0; s < cfg.n_slots; ++s) {
            if (!slot_comp[s])
                w.u8(0);
            else {
                w.u8(1);
                w.qm31(slot_comp[s]->claimed_sum);
            }
        }
    }
    w.u32(cfg.pow_bits);
    w.u32(cfg.log_blowup_factor);
    w.u32(cfg.log_last_layer_degree_bound);
    w.u64(cfg.n_queries);
    w.u64(4);
    for (int t = 0; t < 4; ++t) w.hash(p.trees[t].merkle.root);
    w.u64(4);
    for (int t = 0; t < 4; ++t) {
        w.u64(p.sampled[t].size());
        for (auto& col : p.sampled[t]) {
            w.u64(col.size());
            for (QM31 v : col) w.qm31(v);
        }
    }
    w.u64(4);
    for (int t = 0; t < 4; ++t) write_decommit(w, p.tree_dec[t], g);
    w.u64(4);
    for (int t = 0; t < 4; ++t) {
        w.u64(p.tree_dec[t].queried_values.size());
        for (size_t i : p.tree_dec[t].queried_values) w.u32(g.values[i]);
    }
    w.u64(p.nonce);
    write_fri_layer(w, p.first_layer, g);
    w.u64(p.inner_layers.size());
    for (auto& l : p.inner_layers) write_fri_layer(w, l, g);
    w.u64(p.last_layer_poly.size());
    for (QM31 v : p.last_layer_poly) w.qm31(v);
    uint32_t lg = 0;
    while (((size_t)1 << lg) < p.last_layer_poly.size()) ++lg;
    w.u32(lg);
}


// non-committed columns shipped with a tree's exchange (sharded prover): the [-1]-shifted copies of the 4 coordinate columns
// col .. col + 3 (the last LogUp column of a component)
struct AuxReq {
    size_t col;
    int domain_log;
    uint32_t* dst;  // this rank's row shards of them: 4 x rl, contiguous
};

// one DEEP-quotient column (a size class of compute_fri_quotients)
struct QuotCol {
    int log;               // of the whole column
    uint32_t* coords[4];   // sharded: this rank's rows [rank * 2^(log - logw), ...)
    uint32_t* full[4];     // sharded: the whole column once it was all-gathered for the replicated FRI layers (else null)
};

// one inner FRI layer (line evaluation + its Merkle tree)
struct FriLayer {
    int log;               // of the whole layer
    uint32_t* coords[4];   // row shard when logw > 0
    MerkleTree tree;
    int logw = 0;
};

// PcsConfig::default() when the caller passes none (prover.rs:36); rejects what the protocol loops cannot handle
lb_prove_config checked_config(const lb_prove_config* cfg_in) {
    lb_prove_config cfg;
    if (cfg_in)
        cfg = *cfg_in;
    else {
        cfg.pow_bits = 5;
        cfg.log_blowup_factor = 1;
        cfg.log_last_layer_degree_bound = 0;
        cfg.n_queries = 3;
        cfg.channel_variant = 0;
        cfg.n_slots = 17;
        cfg.air_era = 0;
        cfg.draw_lookup_elements = 1;
    }
    if (cfg.n_slots < 2 || cfg.n_slots > 64) fail(LB_ERR_BAD_ARG, "prove: bad n_slots");
    if (cfg.log_blowup_factor < 1 || cfg.log_blowup_factor > 4) fail(LB_ERR_BAD_ARG, "prove: bad blow-up");
    // the grind loop ends only when a nonce exists: a 64-bit nonce cannot be asked for more than 64 zero bits, and anything
    // near that never finishes; stwo's own configurations stay below 32
    if (cfg.pow_bits > 40) fail(LB_ERR_BAD_ARG, "prove: pow_bits > 40");
    if (cfg.channel_variant != 0 && cfg.channel_variant != 1) fail(LB_ERR_BAD_ARG, "prove: unknown channel variant");
    if (cfg.log_last_layer_degree_bound > 16 || cfg.n_queries < 1 || cfg.n_queries > 4096)
        fail(LB_ERR_BAD_ARG, "prove: bad FRI configuration");
    return cfg;
}

Shard make_shard(lb_ctx* ctx, lb_comm* comm) {
    Shard sh;
    if (comm && comm->world > 1) {
        if (comm->ctx != ctx) fail(LB_ERR_BAD_ARG, "prove: the communicator belongs to another context");
        sh.comm = comm;
        sh.rank = comm->rank;
        sh.world = comm->world;
        sh.logw = comm->log_world;
        comm->bytes_sent = comm->bytes_received = 0;
        comm->n_collectives = 0;
    }
    return sh;
}

// One luminair_prover::prover::prove call (crates/prover/src/prover.rs:28-319 followed by stwo::prover::prove): the state the
// phases hand to each other, one method per phase.  `sh` is the row/column shard of this rank when the proof is produced by
// several GPUs together (lb_prove_sharded); with one GPU every sharded branch is skipped.
class ProveJob {
   public:
    ProveJob(lb_ctx* ctx_, const lb_trace_table* tables_, int n_tables_, const lb_preprocessed_column* pre_in_, int n_pre_,
             const lb_prove_config& cfg_, lb_comm* comm)
        : ctx(ctx_), tables(tables_), n_tables(n_tables_), pre_in(pre_in_), n_pre(n_pre_), cfg(cfg_), sh(make_shard(ctx_, comm)),
          blowup((int)cfg_.log_blowup_factor), st(ctx_->stream), tw(ctx_->tw), timer(ctx_), arena(ctx_),
          channel(cfg_.channel_variant, &ctx_->transcript), trees(4), lut_log(REL_COUNT, -1) {}

    void run(std::vector<uint8_t>& out) {
        setup();
        commit_preprocessed();
        commit_main_trace();
        timer.lap();  // stage 0: upload + interpolate + LDE + Merkle of the main trace
        commit_interaction_trace();
        timer.lap();  // stage 1: LogUp + interpolate + LDE + Merkle of the interaction trace
        commit_composition();
        timer.lap();  // stage 2: constraint quotients + composition commit
        sample_at_oods();
        timer.lap();  // stage 3: OODS sampling
        deep_quotients();
        timer.lap();  // stage 4: DEEP quotients
        fri_commit();
        timer.lap();  // stage 5: FRI commit
        nonce = grind_nonce(channel, cfg, arena, st);
        channel.mix_u64(nonce);
        decommit();
        timer.lap();  // stage 6: grind + queries + decommitment
        // prove() returns ConstraintsNotSatisfied otherwise
        check_oods(comps, sampled, rels, random_coeff, oods);
        ProofParts parts{claim, comps, trees, sampled, tree_dec, nonce, first_out, inner_out, last_layer_poly};
        write_proof(out, cfg, parts, g);
        timer.lap();  // stage 7: OODS check + serialisation
    }

   private:
    // ---- the call
    lb_ctx* ctx;
    const lb_trace_table* tables;
    int n_tables;
    const lb_preprocessed_column* pre_in;
    int n_pre;
    const lb_prove_config cfg;
    Shard sh;
    const int blowup;
    cudaStream_t st;
    const Twiddles& tw;
    StageTimer timer;
    Arena arena;
    Channel channel;
    std::vector<CommitTree> trees;  // preprocessed, main, interaction, composition
    bool use_ipc = false;           // fused exchange: LDE tiles stored straight into the owner rank's row-shard buffer (NVLink)
    uint32_t* d_barrier = nullptr;
    // ---- what the phases hand on
    std::vector<PreCol> pre_cols;
    std::vector<int> lut_log;
    std::vector<int> claim;
    std::vector<Component> by_slot;
    Relations rels{};
    std::vector<Component> comps;  // slot order = LuminairComponents order
    std::vector<AuxReq> inter_aux;
    QM31 random_coeff;
    QPt oods;
    std::vector<std::vector<std::vector<QPt>>> sample_points;  // [tree][col] = list of points
    SampledValues sampled;
    std::vector<QuotCol> quotients;
    MerkleTree fri_first_tree;
    std::vector<FriLayer> inner;
    std::vector<QM31> last_layer_poly;
    uint64_t nonce = 0;
    Gatherer g;
    FriLayerOut first_out;
    std::vector<FriLayerOut> inner_out;
    std::vector<DecommitIdx> tree_dec;

    void setup() {
        // ---- sizes, twiddles -------------------------------------------------------------
        int max_log = 0;
        for (int t = 0; t < n_tables; ++t) {
            if (tables[t].n_rows == 0) fail(LB_ERR_BAD_ARG, "TraceError::EmptyTrace");
            int lg = 0;
            while (((uint64_t)1 << lg) < tables[t].n_rows) ++lg;
            if (lg < 4) lg = 4;  // N_LANES = 16, crates/air/src/utils.rs:22-27
            if (lg > 24) fail(LB_ERR_BAD_ARG, "prove: table too large");
            max_log = std::max(max_log, lg);
        }
        if (n_pre < 0 || n_pre > 16 || (n_pre > 0 && !pre_in)) fail(LB_ERR_BAD_ARG, "prove: bad preprocessed column list");
        for (int k = 0; k < n_pre; ++k) {
            if (pre_in[k].log_size < 4 || pre_in[k].log_size > 24) fail(LB_ERR_BAD_ARG, "prove: bad LUT column size");
            if (pre_in[k].lut < REL_SIN || pre_in[k].lut > REL_RANGE_CHECK || pre_in[k].col_index < 0 || pre_in[k].col_index > 1)
                fail(LB_ERR_BAD_ARG, "prove: unknown LUT column");
            max_log = std::max(max_log, pre_in[k].log_size);
        }
        {
            int r = lb_twiddles_ensure(ctx, max_log + 1 + blowup);
            if (r) fail(r, ctx->err);
        }
        if (sh.on()) {
            if (!nccl_api().load()) fail(LB_ERR_NCCL, nccl_api().error);
            // symmetric heap: the row shards of every committed column + the shifted LogUp copies, the same layout on all ranks
            size_t words = 0;
            auto shard_words = [&](int lde_log, size_t n_cols) { return ((((size_t)n_cols << lde_log) >> sh.logw) + 64); };
            int lut_sz[REL_COUNT];
            for (int k = 0; k < REL_COUNT; ++k) lut_sz[k] = -1;
            for (int k = 0; k < n_pre; ++k)
                if (pre_in[k].lut >= 0 && pre_in[k].lut < REL_COUNT) lut_sz[pre_in[k].lut] = pre_in[k].log_size;
            int top_eval = 0;
            for (int t = 0; t < n_tables; ++t) {
                int lg = 0;
                while (((uint64_t)1 << lg) < tables[t].n_rows) ++lg;
                if (lg < 4) lg = 4;
                int kind = kind_of_slot(tables[t].slot, cfg.n_slots, cfg.air_era);
                if (kind < 0) continue;
                ComponentShape shp = component_shape(kind);
                int ev = ((consumes_lut(kind) && shp.lut && lut_sz[shp.lut] >= 0) ? std::max(lg, lut_sz[shp.lut]) : lg) + 1;
                top_eval = std::max(top_eval, ev);
                words += shard_words(lg + blowup, shp.n_main) + shard_words(lg + blowup, 4 * shp.n_fracs) + shard_words(lg + blowup, 4);
                // constraint evaluation on another domain than the committed one: the columns are sharded a second time
                if (ev != lg + blowup) words += shard_words(ev, shp.n_main) + shard_words(ev, 4 * shp.n_fracs) + shard_words(ev, 4 + 2);
            }
            for (int k = 0; k < n_pre; ++k) words += shard_words(pre_in[k].log_size + blowup, 1) + shard_words(top_eval, 1);
            words += shard_words(top_eval + blowup, 4);  // composition: log = the largest evaluation domain
            d_barrier = arena.alloc<uint32_t>(1);
            ck(cudaMemsetAsync(d_barrier, 0, 4, st), "memset");
            use_ipc = ensure_symmetric(ctx, arena, sh, words);
            sh.comm->sym_used = 0;
            sh.comm->bytes_peer_stored = 0;
            // no rank may store into a peer's heap before that peer is done with the previous proof
            if (use_ipc) shard_barrier(ctx, sh, d_barrier);
        }
    }

    ColRun push_run(CommitTree& tree, uint32_t* coeffs, int n, int lg) {
        // n columns of 2^lg coefficients, contiguous; returns this rank's sub-range [a, b) of them
        std::vector<int> owner = split_run(sh, n);
        ColRun run{tree.cols.size(), n, lg};
        for (int k = 0; k < n; ++k) {
            PolyCol pc{coeffs ? coeffs + ((size_t)k << lg) : nullptr, nullptr, lg};
            pc.owner = owner[k];
            tree.cols.push_back(pc);
        }
        tree.runs.push_back(run);
        return run;
    }

    // Sharded: evaluate the polynomials of one run (each on its owner rank) on CanonicCoset(L) and leave every rank with
    // its row shard [rank R/W, (rank+1) R/W) of ALL of them (run.n columns at stride R/W; returned).  Transfers that need
    // NCCL are appended to `xs` (the caller runs ONE grouped exchange for everything it shards), temporaries to `to_free`.
    // `aux`: [-1]-shifted copies of columns of this run to ship along (constraint evaluation on row shards).
    uint32_t* shard_run(CommitTree& tree, const ColRun& run, int L, const std::vector<AuxReq>* aux, std::vector<Xfer>& xs,
                        std::vector<uint32_t*>& to_free) {
        const int lg = run.log;
        const size_t stride = (size_t)1 << lg, out_stride = (size_t)1 << L;
        if (L - sh.logw < 2) fail(LB_ERR_BAD_ARG, "prove (sharded): a committed column has fewer than 4 rows per rank");
        int a, b;
        own_range(tree, run, sh.rank, a, b);
        const size_t rl = out_stride >> sh.logw;
        uint32_t* local = use_ipc ? sym_alloc(sh.comm, rl * run.n) : arena.alloc<uint32_t>(rl * run.n);
        // Large columns: the last pass of the transform (cfft_evaluate_scatter) writes every 4096-row tile straight
        // to where it belongs.  Fused exchange (use_ipc): that is the owner rank's row-shard buffer itself, mapped over
        // NVLink - no staging, no message, the transfer overlaps the butterflies tile by tile.  Otherwise: a
        // per-destination staging block, so the exchange is ONE contiguous NCCL message per peer and run, received in
        // place (an owner's columns are contiguous in the run).  Small columns: plain transform, one message per column
        // and peer.
        const bool packed = L >= 16 && L - sh.logw >= 12 && sh.world <= 8;
        uint32_t* own = nullptr;   // !packed: the owned columns, whole
        uint32_t* pack = nullptr;  // packed: [peer][owned column][rl] (this rank's slot unused)
        if (b > a) {
            own = arena.alloc<uint32_t>(out_stride * (b - a));
            to_free.push_back(own);
            if (packed) {
                uint32_t* peers[8];
                if (use_ipc) {
                    for (int r = 0; r < sh.world; ++r) peers[r] = peer_ptr(sh.comm, r, local) + (size_t)a * rl;
                    sh.comm->bytes_peer_stored += (size_t)(b - a) * rl * 4 * (size_t)(sh.world - 1);
                } else {
                    pack = arena.alloc<uint32_t>(out_stride * (b - a));
                    to_free.push_back(pack);
                    for (int r = 0; r < sh.world; ++r)
                        peers[r] = r == sh.rank ? local + (size_t)a * rl : pack + (size_t)r * (b - a) * rl;
                }
                ck(cfft_evaluate_scatter(&tw, tree.cols[run.first + a].coeffs, stride, lg, own, out_stride, L, b - a, peers,
                                         sh.world, 0, ctx->sm_count, st),
                   "LDE evaluate + scatter (own columns)");
                if (!use_ipc)
                    for (int r = 0; r < sh.world; ++r)
                        if (r != sh.rank) xs.push_back({peers[r], nullptr, (size_t)(b - a) * rl, r, true});
            } else {
                ck(cfft_evaluate(&tw, tree.cols[run.first + a].coeffs, stride, lg, own, out_stride, L, b - a, ctx->sm_count, st),
                   "LDE evaluate (own columns)");
            }
        }
        if (packed && !use_ipc) {
            for (int r = 0; r < sh.world; ++r) {
                if (r == sh.rank) continue;
                int ra, rb;
                own_range(tree, run, r, ra, rb);
                if (rb > ra) xs.push_back({nullptr, local + (size_t)ra * rl, (size_t)(rb - ra) * rl, r, false});
            }
        }
        for (int k = 0; k < run.n; ++k) {
            const PolyCol& pc = tree.cols[run.first + k];
            uint32_t* mine = local + (size_t)k * rl;
            if (!packed) {
                if (pc.owner == sh.rank) {
                    const uint32_t* full = own + (size_t)(k - a) * out_stride;
                    for (int r = 0; r < sh.world; ++r) xs.push_back({full + (size_t)r * rl, r == sh.rank ? mine : nullptr, rl, r, true});
                } else {
                    xs.push_back({nullptr, mine, rl, pc.owner, false});
                }
            }
        }
        if (aux)
            for (const AuxReq& rq : *aux) {
                if (rq.col < run.first || rq.col >= run.first + (size_t)run.n) continue;
                // the four columns may have different owners; each owner's share is contiguous: one message per
                // owner and peer, written by the shift kernel straight into per-destination staging blocks
                const int k0 = (int)(rq.col - run.first);
                for (int r = 0; r < sh.world; ++r) {
                    int q0 = 4, q1 = 0;  // columns k0 + q0 .. k0 + q1 - 1 are owned by rank r
                    for (int q = 0; q < 4; ++q)
                        if (tree.cols[rq.col + q].owner == r) {
                            q0 = std::min(q0, q);
                            q1 = std::max(q1, q + 1);
                        }
                    if (q1 <= q0) continue;
                    const size_t cnt = (size_t)(q1 - q0) * rl;
                    if (r != sh.rank) {
                        if (!use_ipc) xs.push_back({nullptr, rq.dst + (size_t)q0 * rl, cnt, r, false});
                        continue;
                    }
                    uint32_t* stage = nullptr;  // [peer][owned shifted column][rl]; fused exchange: not needed
                    if (!use_ipc) {
                        stage = arena.alloc<uint32_t>(cnt * sh.world);
                        to_free.push_back(stage);
                    } else {
                        sh.comm->bytes_peer_stored += cnt * 4 * (size_t)(sh.world - 1);
                    }
                    for (int q = q0; q < q1; ++q) {
                        const int k = k0 + q;
                        const uint32_t* src[8];
                        uint32_t* dst[8];
                        for (int t = 0; t < sh.world; ++t) {
                            if (packed && use_ipc)  // the column's shards already sit in their owners' buffers
                                src[t] = peer_ptr(sh.comm, t, local) + (size_t)k * rl;
                            else if (packed)
                                src[t] = (t == sh.rank ? local + (size_t)a * rl : pack + (size_t)t * (b - a) * rl) + (size_t)(k - a) * rl;
                            else
                                src[t] = own + (size_t)(k - a) * out_stride + (size_t)t * rl;
                            if (use_ipc)
                                dst[t] = peer_ptr(sh.comm, t, rq.dst) + (size_t)q * rl;
                            else
                                dst[t] = (t == sh.rank ? rq.dst + (size_t)q * rl : stage + (size_t)t * cnt + (size_t)(q - q0) * rl);
                        }
                        ck(shifted_prev_column(dst, sh.world, src, sh.world, rq.domain_log, L, st), "shifted column");
                    }
                    if (!use_ipc)
                        for (int t = 0; t < sh.world; ++t)
                            if (t != sh.rank) xs.push_back({stage + (size_t)t * cnt, nullptr, cnt, t, true});
                }
            }
        return local;
    }

    void commit_tree(CommitTree& tree, const std::vector<AuxReq>* aux = nullptr) {
        // evaluate every polynomial on CanonicCoset(log + blowup), Merkle-commit, mix the root.  Sharded: every rank
        // extends the columns it owns, one grouped exchange turns the column shards into row shards (rank r gets rows
        // [r R/W, (r+1) R/W) of EVERY column), each rank hashes the sub-tree over its rows, the roots are all-gathered.
        std::vector<ColRef> refs;
        std::vector<Xfer> xs;
        std::vector<uint32_t*> to_free;
        for (const ColRun& run : tree.runs) {
            const int lg = run.log, L = lg + blowup;
            const size_t stride = (size_t)1 << lg, out_stride = (size_t)1 << L;
            if (!sh.on()) {
                uint32_t* lde = arena.alloc<uint32_t>(out_stride * run.n);
                ck(cfft_evaluate(&tw, tree.cols[run.first].coeffs, stride, lg, lde, out_stride, L, run.n, ctx->sm_count, st),
                   "LDE evaluate");
                for (int k = 0; k < run.n; ++k) {
                    tree.cols[run.first + k].lde = lde + (size_t)k * out_stride;
                    refs.push_back({tree.cols[run.first + k].lde, L});
                }
                continue;
            }
            uint32_t* local = shard_run(tree, run, L, aux, xs, to_free);
            const size_t rl = out_stride >> sh.logw;
            for (int k = 0; k < run.n; ++k) {
                tree.cols[run.first + k].lde = local + (size_t)k * rl;
                refs.push_back({tree.cols[run.first + k].lde, L - sh.logw});
            }
        }
        if (sh.on()) {
            if (!xs.empty()) run_exchange(ctx, sh, xs);
            // fused exchange: my rows are complete once EVERY rank's kernels have run - a stream-ordered barrier
            if (use_ipc) shard_barrier(ctx, sh, d_barrier);
            for (uint32_t* f : to_free) arena.release(f);  // stream-ordered: freed after the sends have read them
        }
        merkle_commit(ctx, arena, refs, tree.merkle, /*fetch_root=*/!sh.on());
        if (sh.on() && !tree.merkle.empty) finish_sharded_root(ctx, arena, sh, tree.merkle);
        channel.mix_root(tree.merkle.root);
    }

    void commit_preprocessed() {
        // ---- phase 0: preprocessed trace (prover.rs:52-59) -----------------------------------------
        // lookups_to_preprocessed_column order from the caller, PreProcessedTrace::new sorts it (stable) by
        // log_size, descending (preprocessed.rs:152-155).  The LUT values themselves are the caller's
        // (host libm, preprocessed.rs:351-383); the path only interpolates and commits them.
        {
            std::vector<int> order(n_pre);
            for (int k = 0; k < n_pre; ++k) order[k] = k;
            std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return pre_in[a].log_size > pre_in[b].log_size; });
            for (int k : order) {
                const lb_preprocessed_column& pc = pre_in[k];
                size_t n = (size_t)1 << pc.log_size;
                uint32_t* evals = arena.alloc<uint32_t>(n);
                ck(cudaMemcpyAsync(evals, pc.values, n * sizeof(uint32_t), pc.on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st),
                   "LUT upload");
                uint32_t* coeffs = arena.alloc<uint32_t>(n);
                ColRun run = push_run(trees[0], coeffs, 1, pc.log_size);
                if (trees[0].cols[run.first].owner == sh.rank) {
                    ck(cudaMemcpyAsync(coeffs, evals, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st), "copy");
                    ck(cfft_interpolate(&tw, coeffs, n, 1, pc.log_size, ctx->sm_count, st), "interpolate LUT");
                }
                pre_cols.push_back({pc.lut, pc.col_index, pc.log_size, evals});
                if (lut_log[pc.lut] >= 0 && lut_log[pc.lut] != pc.log_size) fail(LB_ERR_BAD_ARG, "prove: LUT columns of one table differ in size");
                lut_log[pc.lut] = pc.log_size;
            }
            bool any_host = false;
            for (int k = 0; k < n_pre; ++k) any_host = any_host || !pre_in[k].on_device;
            if (any_host) ck(cudaStreamSynchronize(st), "LUT upload sync");  // host columns may be pageable
        }
        commit_tree(trees[0]);
    }

    void commit_main_trace() {
        // ---- phase 1: main trace (prover.rs:66-179) ------------------------------------------
        claim.assign(cfg.n_slots, -1);
        by_slot.assign(cfg.n_slots, Component{});
        for (int t = 0; t < n_tables; ++t) {
            const lb_trace_table& tb = tables[t];
            if (tb.slot < 0 || tb.slot >= cfg.n_slots) fail(LB_ERR_BAD_ARG, "prove: slot out of range");
            int kind = kind_of_slot(tb.slot, cfg.n_slots, cfg.air_era);
            if (kind < 0) fail(LB_ERR_BAD_ARG, "prove: component not supported by this backend yet");
            ComponentShape shp = component_shape(kind);
            if (tb.n_cols != shp.n_main) fail(LB_ERR_BAD_ARG, "prove: wrong column count for component");
            if (claim[tb.slot] >= 0) fail(LB_ERR_BAD_ARG, "prove: duplicate table for slot");
            // LuminairGraph::gen_trace emits tables in claim-slot order (graph.rs:502-593); the column spans the
            // components read (TraceLocationAllocator) only line up with the committed order in that case
            if (t > 0 && tables[t - 1].slot > tb.slot) fail(LB_ERR_BAD_ARG, "prove: trace tables must come in claim-slot order");
            int lg = 0;
            while (((uint64_t)1 << lg) < tb.n_rows) ++lg;
            if (lg < 4) lg = 4;
            size_t n = (size_t)1 << lg;
            const uint32_t* d_rows;
            uint32_t* staged = nullptr;
            if (tb.rows_on_device) {
                d_rows = tb.rows;
            } else {
                staged = arena.alloc<uint32_t>(tb.n_rows * tb.n_cols);
                ck(cudaMemcpyAsync(staged, tb.rows, tb.n_rows * tb.n_cols * sizeof(uint32_t), cudaMemcpyHostToDevice, st),
                   "trace upload");
                d_rows = staged;
            }
            uint32_t* evals = arena.alloc<uint32_t>(n * shp.n_main);
            ck(transpose_pad(evals, n, d_rows, tb.n_rows, shp.n_main, lg, kind, st), "transpose");
            uint32_t* coeffs = arena.alloc<uint32_t>(n * shp.n_main);
            size_t main_first = trees[1].cols.size();
            {
                // every rank has the whole table (the trace is replicated input); it interpolates the columns it owns
                ColRun run = push_run(trees[1], coeffs, shp.n_main, lg);
                int a, b;
                own_range(trees[1], run, sh.rank, a, b);
                if (b > a) {
                    ck(cudaMemcpyAsync(coeffs + (size_t)a * n, evals + (size_t)a * n, n * (size_t)(b - a) * sizeof(uint32_t),
                                       cudaMemcpyDeviceToDevice, st), "copy");
                    ck(cfft_interpolate(&tw, coeffs + (size_t)a * n, n, b - a, lg, ctx->sm_count, st), "interpolate");
                }
            }
            if (staged) {
                ck(cudaStreamSynchronize(st), "upload sync");  // host rows may be pageable
                arena.release(staged);
            }
            Component c{};
            c.kind = kind;
            c.slot = tb.slot;
            c.log = lg;
            c.main_evals = evals;
            if (shp.lut) {
                if (lut_log[shp.lut] < 0) fail(LB_ERR_BAD_ARG, "prove: component needs a lookup table that was not supplied");
                for (int q = 0; q < shp.n_pre; ++q) {
                    for (size_t i = 0; i < pre_cols.size(); ++i)
                        if (pre_cols[i].lut == shp.lut && pre_cols[i].col_index == q) c.pre_idx[q] = (int)i;
                    if (c.pre_idx[q] < 0) fail(LB_ERR_BAD_ARG, "prove: missing LUT column");
                    if (pre_cols[c.pre_idx[q]].log != lg) fail(LB_ERR_BAD_ARG, "prove: lookup-table component and LUT column differ in size");
                }
            }
            // max_constraint_log_degree_bound (add/component.rs:33-35; LUT consumers exp2/component.rs:41-43)
            c.eval_log = (consumes_lut(kind) ? std::max(lg, lut_log[shp.lut]) : lg) + 1;
            c.main_loc = main_first;  // location by pie order; components use slot order (see below)
            claim[tb.slot] = lg;
            by_slot[tb.slot] = c;
        }
        for (int s = 0; s < cfg.n_slots; ++s)
            if (claim[s] >= 0) channel.mix_u64((uint64_t)claim[s]);  // LuminairClaim::mix_into
        commit_tree(trees[1]);
    }

    void commit_interaction_trace() {
        // ---- phase 2: interaction trace (prover.rs:181-298) -----------------------------------
        // LuminairInteractionElements::draw (components/mod.rs:227-235): node, then sin, exp2, log2, range_check
        {
            std::vector<QM31> el = channel.draw_secure_felts(2);
            rels.r[REL_NODE] = Relation2{el[0], el[1]};
            if (cfg.draw_lookup_elements)
                for (int k = REL_SIN; k <= REL_RANGE_CHECK; ++k) {
                    el = channel.draw_secure_felts(2);  // relation!(X, 1) still draws (z, alpha)
                    rels.r[k] = Relation2{el[0], el[1]};
                }
        }
        if (!cfg.draw_lookup_elements && n_pre) fail(LB_ERR_BAD_ARG, "prove: LUT columns need draw_lookup_elements");

        {
            // TraceLocationAllocator hands out spans in component (slot) order
            size_t main_next = 0;
            int n_present = 0;
            for (int s = 0; s < cfg.n_slots; ++s) n_present += claim[s] >= 0;
            // claimed sums of all components: one device array, one copy back, one synchronisation
            uint32_t* d_claimed = arena.alloc<uint32_t>(4 * (size_t)std::max(n_present, 1));
            ck(cudaMemsetAsync(d_claimed, 0, 4 * (size_t)std::max(n_present, 1) * sizeof(uint32_t), st), "memset");
            std::vector<Xfer> xs;
            std::vector<ColRun> inter_runs;
            std::vector<uint32_t*> inter_bufs;
            int ci = 0;
            for (int s = 0; s < cfg.n_slots; ++s) {
                if (claim[s] < 0) continue;
                Component c = by_slot[s];
                ComponentShape shp = component_shape(c.kind);
                size_t n = (size_t)1 << c.log;
                int n_ic = 4 * shp.n_fracs;
                uint32_t* inter = arena.alloc<uint32_t>(n * n_ic);
                // Sharded: the LogUp columns of component i are generated by rank i mod W (the trace is replicated input) and
                // every column is sent, whole, to the rank that owns it for interpolation / LDE; the claimed sums are summed
                // over the ranks (one contributor each)
                const int compute_rank = sh.on() ? ci % sh.world : 0;
                if (compute_rank == sh.rank) {
                    uint32_t* scan_tmp = arena.alloc<uint32_t>(4 * n);
                    uint32_t* block_sums = arena.alloc<uint32_t>(4 * (n / 1024 + 1));
                    PreCols pc{};
                    for (int q = 0; q < shp.n_pre; ++q) pc.p[q] = pre_cols[c.pre_idx[q]].evals;
                    ck(logup_interaction_trace(c.kind, c.main_evals, n, pc, inter, n, c.log, rels, scan_tmp, block_sums,
                                               d_claimed + 4 * (size_t)ci, st),
                       "logup");
                    arena.release(scan_tmp);  // stream-ordered
                    arena.release(block_sums);
                }
                arena.release(c.main_evals);
                c.main_evals = nullptr;
                c.inter_loc = trees[2].cols.size();
                ColRun run = push_run(trees[2], inter, n_ic, c.log);
                inter_runs.push_back(run);
                inter_bufs.push_back(inter);
                if (sh.on()) {
                    for (int o = 0; o < sh.world; ++o) {
                        if (o == compute_rank) continue;
                        int oa, ob;
                        own_range(trees[2], run, o, oa, ob);
                        if (ob <= oa) continue;
                        if (sh.rank == compute_rank) xs.push_back({inter + (size_t)oa * n, nullptr, (size_t)(ob - oa) * n, o, true});
                        if (sh.rank == o) xs.push_back({nullptr, inter + (size_t)oa * n, (size_t)(ob - oa) * n, compute_rank, false});
                    }
                    // the [-1] mask of the last LogUp column reads a predecessor row that lives in another rank's row shard:
                    // the column's owner ships a shifted copy with the exchange of this tree
                    // (a component evaluated on another domain - lookup table larger than its trace, blow-up other than 2 -
                    // gets its shifted copy when its columns are re-sharded on that domain, see the constraint loop)
                    if (c.eval_log == c.log + blowup) {
                        size_t rl = ((size_t)1 << (c.log + blowup)) >> sh.logw;
                        c.inter_prev = use_ipc ? sym_alloc(sh.comm, 4 * rl) : arena.alloc<uint32_t>(4 * rl);
                        inter_aux.push_back({c.inter_loc + (size_t)(n_ic - 4), c.log, c.inter_prev});
                    }
                }
                c.main_loc = main_next;  // span in slot order (equals the pie-order location when the pie is slot-ordered)
                main_next += shp.n_main;
                comps.push_back(c);
                ++ci;
            }
            if (sh.on()) {
                if (!xs.empty()) run_exchange(ctx, sh, xs);
                nck(nccl_api().AllReduce(d_claimed, d_claimed, 4 * (size_t)n_present, ncclUint32, ncclSum, sh.comm->comm, st),
                    "ncclAllReduce(claimed sums)");
                sh.comm->n_collectives++;
            }
            std::vector<uint32_t> cl(4 * (size_t)std::max(n_present, 1));
            ck(cudaMemcpyAsync(cl.data(), d_claimed, cl.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, st), "claimed d2h");
            for (size_t k = 0; k < inter_runs.size(); ++k) {
                int a, b;
                own_range(trees[2], inter_runs[k], sh.rank, a, b);
                size_t n = (size_t)1 << inter_runs[k].log;
                if (b > a)
                    ck(cfft_interpolate(&tw, inter_bufs[k] + (size_t)a * n, n, b - a, inter_runs[k].log, ctx->sm_count, st),
                       "interpolate interaction");
            }
            ck(cudaStreamSynchronize(st), "claimed sync");
            for (size_t k = 0; k < comps.size(); ++k) comps[k].claimed_sum = q_make(cl[4 * k], cl[4 * k + 1], cl[4 * k + 2], cl[4 * k + 3]);
        }
        for (const Component& c : comps) channel.mix_felts({c.claimed_sum});  // LuminairInteractionClaim::mix_into
        commit_tree(trees[2], sh.on() ? &inter_aux : nullptr);
    }

    void commit_composition() {
        // ---- stwo::prover::prove -------------------------------------------------------------
        random_coeff = channel.draw_secure_felt();
        // constraint counts from the AIR itself
        std::vector<int> n_constraints;
        int total_constraints = 0;
        for (const Component& c : comps) {
            InfoEval info;
            eval_component(c.kind, info, rels);
            ComponentShape shp = component_shape(c.kind);
            if (info.n_main != shp.n_main || info.n_inter != 4 * shp.n_fracs || info.n_constraints != shp.n_constraints ||
                info.n_pre != shp.n_pre || shp.n_constraints > MAX_CONSTRAINTS || shp.n_main > MAX_MAIN_COLS)
                fail(LB_ERR_BAD_ARG, "internal: component shape table out of date");
            n_constraints.push_back(info.n_constraints);
            total_constraints += info.n_constraints;
        }
        std::vector<QM31> powers(total_constraints);
        {
            QM31 cur = q_one();
            for (int i = 0; i < total_constraints; ++i) {
                powers[i] = cur;
                cur = q_mul(cur, random_coeff);
            }
        }
        // per evaluation-domain size accumulators (DomainEvaluationAccumulator)
        std::map<int, uint32_t*> acc;  // eval_log -> 4 coordinate columns (stride 2^eval_log)
        {
            int remaining = total_constraints;
            for (size_t ci = 0; ci < comps.size(); ++ci) {
                const Component& c = comps[ci];
                int nc = n_constraints[ci];
                int eval_log = c.eval_log;
                ComponentShape shp = component_shape(c.kind);
                ConstraintParams p{};
                size_t ne = (size_t)1 << eval_log;
                // constraint-framework `need_to_extend`: columns not committed on the evaluation domain are
                // re-evaluated there from their polynomials
                std::vector<uint32_t*> scratch;
                const size_t ne_local = ne >> sh.logw;  // rows of the evaluation domain this rank holds
                std::vector<Xfer> ext_xs;            // sharded: the re-evaluated columns go column shards -> row shards like a tree
                std::vector<uint32_t*> ext_free;
                uint32_t* ext_prev = nullptr;
                bool extended = false;
                auto on_eval_domain = [&](CommitTree& tree, size_t first, int n_cols, bool with_shifted = false) -> const uint32_t* {
                    if (tree.cols[first].log + blowup == eval_log) return tree.cols[first].lde;
                    if (sh.on()) {
                        const ColRun* run = nullptr;
                        for (const ColRun& r : tree.runs)
                            if (r.first == first && r.n == n_cols) run = &r;
                        if (!run) fail(LB_ERR_BAD_ARG, "internal: component columns are not one run");
                        std::vector<AuxReq> one;
                        if (with_shifted) {
                            ext_prev = use_ipc ? sym_alloc(sh.comm, 4 * ne_local) : arena.alloc<uint32_t>(4 * ne_local);
                            one.push_back({first + (size_t)(n_cols - 4), c.log, ext_prev});
                        }
                        extended = true;
                        return shard_run(tree, *run, eval_log, with_shifted ? &one : nullptr, ext_xs, ext_free);
                    }
                    uint32_t* ext = arena.alloc<uint32_t>(ne * n_cols);
                    scratch.push_back(ext);
                    int lg = tree.cols[first].log;
                    ck(cfft_evaluate(&tw, tree.cols[first].coeffs, (size_t)1 << lg, lg, ext, ne, eval_log, n_cols, ctx->sm_count, st),
                       "extend to evaluation domain");
                    return ext;
                };
                p.main = on_eval_domain(trees[1], c.main_loc, shp.n_main);
                p.main_stride = ne_local;
                p.inter = on_eval_domain(trees[2], c.inter_loc, 4 * shp.n_fracs, /*with_shifted=*/true);
                p.inter_stride = ne_local;
                for (int q = 0; q < shp.n_pre; ++q) p.pre.p[q] = on_eval_domain(trees[0], (size_t)c.pre_idx[q], 1);
                if (sh.on()) {
                    if (extended) {
                        if (!ext_xs.empty()) run_exchange(ctx, sh, ext_xs);
                        if (use_ipc) shard_barrier(ctx, sh, d_barrier);
                        for (uint32_t* f : ext_free) arena.release(f);
                    }
                    p.row0 = (uint32_t)(ne_local * sh.rank);
                    p.n_rows = (uint32_t)ne_local;
                    p.inter_prev = ext_prev ? ext_prev : c.inter_prev;
                    if (!p.inter_prev) fail(LB_ERR_BAD_ARG, "internal: no shifted LogUp column for a row-sharded component");
                }
                bool fresh = acc.find(eval_log) == acc.end();
                if (fresh) acc[eval_log] = arena.alloc<uint32_t>(4 * ne_local);
                for (int k = 0; k < 4; ++k) p.acc[k] = acc[eval_log] + (size_t)k * ne_local;
                p.accumulate = fresh ? 0 : 1;
                p.log_size = c.log;
                p.eval_log = eval_log;
                p.rels = rels;
                p.cumsum_shift = q_mul_m(c.claimed_sum, m_inv((uint32_t)(((uint64_t)1 << c.log) % P)));
                // this component owns the last `nc` of the remaining powers, highest first
                for (int k = 0; k < nc; ++k) p.pows[k] = powers[remaining - 1 - k];
                remaining -= nc;
                // 1 / Z_H on the 2^(eval_log - log) cosets of the evaluation domain
                int log_expand = eval_log - c.log;
                uint32_t init = subgroup_gen(eval_log + 1), step = subgroup_gen(eval_log - 1);
                std::vector<uint32_t> dinv((size_t)1 << log_expand);
                for (uint32_t i = 0; i < (1u << log_expand); ++i) {
                    Pt pt = host_index_to_point(init + step * bit_reverse(i, log_expand));
                    dinv[i] = m_inv(coset_vanishing_m(c.log, pt));
                }
                p.denom_inv = arena.upload(dinv);
                ck(constraint_quotients(c.kind, p, st), "constraint quotients");
                for (uint32_t* e : scratch) arena.release(e);  // stream-ordered: freed once the constraint kernel has read them
            }
        }
        // finalize: lift smaller accumulators into larger ones, interpolate -> composition coefficients
        uint32_t* comp_coeffs = nullptr;
        int comp_log = 0;
        if (!sh.on()) {
            for (auto& kv : acc) {  // ascending eval_log
                int lg = kv.first;
                uint32_t* vals = kv.second;
                size_t n = (size_t)1 << lg;
                if (comp_coeffs) {
                    uint32_t* lifted = arena.alloc<uint32_t>(4 * n);
                    ck(cfft_evaluate(&tw, comp_coeffs, (size_t)1 << comp_log, comp_log, lifted, n, lg, 4, ctx->sm_count, st),
                       "lift composition");
                    ck(add_inplace(vals, lifted, 4 * n, st), "accumulate");
                }
                ck(cfft_interpolate(&tw, vals, n, 4, lg, ctx->sm_count, st), "interpolate composition");
                comp_coeffs = vals;
                comp_log = lg;
            }
            push_run(trees[3], comp_coeffs, 4, comp_log);
        } else {
            // the accumulators are row-sharded; interpolation is column-wise: coordinate column k of every size class goes
            // to the rank that owns composition column k (a rows -> columns exchange of 16 B per row), which lifts, adds and
            // interpolates it
            std::vector<int> owner = split_run(sh, 4);
            int a = 4, b = 0;
            for (int k = 0; k < 4; ++k)
                if (owner[k] == sh.rank) {
                    a = std::min(a, k);
                    b = std::max(b, k + 1);
                }
            std::map<int, uint32_t*> full;  // eval_log -> this rank's coordinate columns, whole domain
            std::vector<Xfer> xs;
            for (auto& kv : acc) {
                const int lg = kv.first;
                const size_t n = (size_t)1 << lg, rl = n >> sh.logw;
                uint32_t* f = b > a ? arena.alloc<uint32_t>(n * (size_t)(b - a)) : nullptr;
                full[lg] = f;
                for (int k = 0; k < 4; ++k) {
                    const int o = owner[k];
                    uint32_t* slot = o == sh.rank ? f + (size_t)(k - a) * n : nullptr;
                    xs.push_back({kv.second + (size_t)k * rl, slot ? slot + rl * sh.rank : nullptr, rl, o, true});
                    if (o == sh.rank)
                        for (int r = 0; r < sh.world; ++r)
                            if (r != sh.rank) xs.push_back({nullptr, slot + rl * r, rl, r, false});
                }
                comp_log = lg;
            }
            run_exchange(ctx, sh, xs);
            if (b > a) {
                int prev_log = 0;
                for (auto& kv : full) {
                    const int lg = kv.first;
                    const size_t n = (size_t)1 << lg;
                    uint32_t* vals = kv.second;
                    if (comp_coeffs) {
                        uint32_t* lifted = arena.alloc<uint32_t>(n * (size_t)(b - a));
                        ck(cfft_evaluate(&tw, comp_coeffs, (size_t)1 << prev_log, prev_log, lifted, n, lg, b - a, ctx->sm_count, st),
                           "lift composition");
                        ck(add_inplace(vals, lifted, n * (size_t)(b - a), st), "accumulate");
                    }
                    ck(cfft_interpolate(&tw, vals, n, b - a, lg, ctx->sm_count, st), "interpolate composition");
                    comp_coeffs = vals;
                    prev_log = lg;
                }
            }
            ColRun run{trees[3].cols.size(), 4, comp_log};
            for (int k = 0; k < 4; ++k) {
                PolyCol pc{owner[k] == sh.rank ? comp_coeffs + ((size_t)(k - a) << comp_log) : nullptr, nullptr, comp_log};
                pc.owner = owner[k];
                trees[3].cols.push_back(pc);
            }
            trees[3].runs.push_back(run);
        }
        commit_tree(trees[3]);
    }

    void sample_at_oods() {
        // ---- OODS point and mask points ---------------------------------------------------------
        {
            QM31 t = channel.draw_secure_felt();
            QM31 t2 = q_mul(t, t);
            QM31 inv = q_inv(q_add(t2, q_one()));
            oods.x = q_mul(q_sub(q_one(), t2), inv);
            oods.y = q_mul(q_add(t, t), inv);
        }
        sample_points.assign(4, {});
        sample_points[0].resize(trees[0].cols.size());  // only the columns a component reads are sampled
        sample_points[1].resize(trees[1].cols.size());
        sample_points[2].resize(trees[2].cols.size());
        for (const Component& c : comps) {
            ComponentShape shp = component_shape(c.kind);
            for (int q = 0; q < shp.n_pre; ++q) sample_points[0][c.pre_idx[q]] = {oods};
            for (int k = 0; k < shp.n_main; ++k) sample_points[1][c.main_loc + k] = {oods};
            int n_ic = 4 * shp.n_fracs;
            QPt prev = qpt_add(oods, qpt_lift(host_index_to_point((0u - subgroup_gen(c.log)) & CIRCLE_ORDER_MASK)));
            for (int k = 0; k < n_ic; ++k) {
                if (k >= n_ic - 4)
                    sample_points[2][c.inter_loc + k] = {prev, oods};  // mask offsets [-1, 0]
                else
                    sample_points[2][c.inter_loc + k] = {oods};
            }
        }
        sample_points[3].assign(4, {oods});

        // ---- prove_values: sample every polynomial -----------------------------------------------
        sampled.assign(4, {});
        for (int t = 0; t < 4; ++t) {
            sampled[t].resize(trees[t].cols.size());
            for (size_t c = 0; c < trees[t].cols.size(); ++c) sampled[t][c].resize(sample_points[t][c].size());
        }
        {
            struct Job {
                int log;
                QPt pt;
                std::vector<const uint32_t*> cols;
                std::vector<QM31*> dst;
            };
            std::vector<Job> jobs;
            for (int t = 0; t < 4; ++t)
                for (size_t c = 0; c < trees[t].cols.size(); ++c)
                    for (size_t s = 0; s < sample_points[t][c].size(); ++s) {
                        if (sh.on() && trees[t].cols[c].owner != sh.rank) continue;  // sampled by its owner, all-reduced below
                        const QPt& pt = sample_points[t][c][s];
                        int lg = trees[t].cols[c].log;
                        Job* job = nullptr;
                        for (Job& j : jobs)
                            if (j.log == lg && qpt_eq(j.pt, pt)) job = &j;
                        if (!job) {
                            jobs.push_back({lg, pt, {}, {}});
                            job = &jobs.back();
                        }
                        job->cols.push_back(trees[t].cols[c].coeffs);
                        job->dst.push_back(&sampled[t][c][s]);
                    }
            std::vector<std::vector<QM31>> results(jobs.size());
            std::vector<QM31*> d_results(jobs.size());
            {
                std::vector<EvalJob> ej;
                for (size_t ji = 0; ji < jobs.size(); ++ji) {
                    d_results[ji] = arena.alloc<QM31>(jobs[ji].cols.size());
                    ej.push_back(EvalJob{jobs[ji].log, jobs[ji].pt, jobs[ji].cols, d_results[ji]});
                }
                launch_eval_jobs(ctx, arena, ej);
            }
            for (size_t ji = 0; ji < jobs.size(); ++ji) {
                results[ji].resize(jobs[ji].cols.size());
                ck(cudaMemcpyAsync(results[ji].data(), d_results[ji], results[ji].size() * sizeof(QM31), cudaMemcpyDeviceToHost, st),
                   "samples d2h");
            }
            ck(cudaStreamSynchronize(st), "samples sync");
            for (size_t ji = 0; ji < jobs.size(); ++ji)
                for (size_t k = 0; k < jobs[ji].dst.size(); ++k) *jobs[ji].dst[k] = results[ji][k];
            if (sh.on()) {
                // every sample has exactly one owner: a sum over the ranks (zeros elsewhere) hands all of them to everybody
                std::vector<uint32_t> flat;
                for (int t = 0; t < 4; ++t)
                    for (size_t c = 0; c < trees[t].cols.size(); ++c)
                        for (QM31 v : sampled[t][c]) {
                            bool mine = trees[t].cols[c].owner == sh.rank;
                            uint32_t w[4] = {v.a.a, v.a.b, v.b.a, v.b.b};
                            for (int q = 0; q < 4; ++q) flat.push_back(mine ? w[q] : 0u);
                        }
                uint32_t* d_flat = arena.upload(flat);
                nck(nccl_api().AllReduce(d_flat, d_flat, flat.size(), ncclUint32, ncclSum, sh.comm->comm, st), "ncclAllReduce(samples)");
                sh.comm->n_collectives++;
                sh.comm->bytes_sent += flat.size() * 4;
                sh.comm->bytes_received += flat.size() * 4;
                ck(cudaMemcpyAsync(flat.data(), d_flat, flat.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, st), "samples d2h");
                ck(cudaStreamSynchronize(st), "samples sync");
                size_t at = 0;
                for (int t = 0; t < 4; ++t)
                    for (size_t c = 0; c < trees[t].cols.size(); ++c)
                        for (QM31& v : sampled[t][c]) {
                            v = q_make(flat[at], flat[at + 1], flat[at + 2], flat[at + 3]);
                            at += 4;
                        }
            }
        }
        {
            std::vector<QM31> flat;
            for (int t = 0; t < 4; ++t)
                for (auto& col : sampled[t])
                    for (QM31 v : col) flat.push_back(v);
            channel.mix_felts(flat);
        }
    }

    void deep_quotients() {
        // ---- DEEP quotients (compute_fri_quotients) -----------------------------------------------
        const QM31 rc_q = channel.draw_secure_felt();
        {
            struct FlatCol {
                const uint32_t* lde;
                int log;
                const std::vector<QPt>* pts;
                const std::vector<QM31>* vals;
            };
            std::vector<FlatCol> flat;
            for (int t = 0; t < 4; ++t)
                for (size_t c = 0; c < trees[t].cols.size(); ++c)
                    flat.push_back({trees[t].cols[c].lde, trees[t].cols[c].log + blowup, &sample_points[t][c], &sampled[t][c]});
            std::vector<int> sizes;
            for (auto& f : flat) sizes.push_back(f.log);
            std::sort(sizes.begin(), sizes.end(), std::greater<int>());
            sizes.erase(std::unique(sizes.begin(), sizes.end()), sizes.end());
            for (int lg : sizes) {
                std::vector<const FlatCol*> grp;
                for (auto& f : flat)
                    if (f.log == lg) grp.push_back(&f);
                // ColumnSampleBatch::new_vec: group by point (BTreeMap order), columns in order
                typedef HostBatch Batch;
                std::vector<Batch> batches;
                for (size_t ci = 0; ci < grp.size(); ++ci)
                    for (size_t s = 0; s < grp[ci]->pts->size(); ++s) {
                        const QPt& pt = (*grp[ci]->pts)[s];
                        Batch* b = nullptr;
                        for (Batch& bb : batches)
                            if (qpt_eq(bb.pt, pt)) b = &bb;
                        if (!b) {
                            batches.push_back({pt, {}});
                            b = &batches.back();
                        }
                        b->cols.push_back({(int)ci, (*grp[ci]->vals)[s]});
                    }
                std::vector<const uint32_t*> colptrs;
                for (auto* f : grp) colptrs.push_back(f->lde);
                QuotCol qc{};
                qc.log = lg;
                const size_t rl = ((size_t)1 << lg) >> sh.logw;
                uint32_t* buf = arena.alloc<uint32_t>(4 * rl);
                for (int k = 0; k < 4; ++k) qc.coords[k] = buf + (size_t)k * rl;
                launch_quotients(ctx, arena, lg, colptrs, batches, rc_q, qc.coords, sh.on() ? (uint32_t)(rl * sh.rank) : 0u,
                                 sh.on() ? (uint32_t)rl : 0u);
                quotients.push_back(qc);
            }
        }
    }

    void fri_commit() {
        // ---- FRI commit (FriProver::commit) ----------------------------------------------------------
        {
            std::vector<ColRef> refs;
            for (auto& q : quotients)
                for (int k = 0; k < 4; ++k) refs.push_back({q.coords[k], q.log - sh.logw});
            merkle_commit(ctx, arena, refs, fri_first_tree, /*fetch_root=*/!sh.on());
            if (sh.on()) finish_sharded_root(ctx, arena, sh, fri_first_tree);
            channel.mix_root(fri_first_tree.root);
        }
        {
            QM31 folding_alpha = channel.draw_secure_felt();
            int line_log = quotients[0].log - 1;
            int last_log = (int)(cfg.log_last_layer_degree_bound + cfg.log_blowup_factor);
            size_t qi = 0;
            // The per-layer Fiat-Shamir step (mix_root, draw the next folding coefficient) runs on the device
            // (channel_mix_root_draw), so fold -> Merkle -> mix -> draw -> fold is enqueued for every layer without a host
            // round trip; roots, digests and the final channel state come back in one copy after the loop.
            const int n_layers = std::max(0, line_log - last_log);
            DevChannel h_ch{};
            channel.digest_words(h_ch.digest);
            h_ch.n_sent = channel.n_sent();
            DevChannel* d_ch = arena.alloc<DevChannel>(1);
            QM31* d_alphas = arena.alloc<QM31>(n_layers + 1);
            uint32_t* d_digests = arena.alloc<uint32_t>(8 * (size_t)std::max(n_layers, 1));
            arena.upload_to(d_ch, &h_ch, sizeof(h_ch));
            arena.upload_to(d_alphas, &folding_alpha, sizeof(QM31));
            int li = 0;
            uint32_t* gathered = nullptr;  // sharded: the first replicated layer, all-gathered from the row shards
            const size_t top_words = 8 * (((size_t)2 << sh.logw) - 1);
            uint32_t* d_tops = nullptr;    // sharded: per row-sharded layer, the top levels of its tree (device)
            int n_sharded_layers = 0;
            if (sh.on()) {
                // Row-sharded layers: a fold pairs adjacent rows, so it stays inside a rank's row range; the layer tree is a
                // sub-tree per rank plus an all-gather of the 32-byte roots, from which every rank hashes the top levels and
                // runs the Fiat-Shamir step on the device (identical inputs, identical state on every rank).  Once a layer is
                // down to 2^FRI_SHARD_MIN_LOG rows per rank it is all-gathered and the remaining (latency-bound) layers run
                // replicated on every rank.
                static const int FRI_SHARD_MIN_LOG = getenv("LB_FRI_SHARD_MIN_LOG") ? atoi(getenv("LB_FRI_SHARD_MIN_LOG")) : 15;
                NcclApi& api = nccl_api();
                uint32_t* cur = nullptr;  // current layer, row shard: 4 coordinate columns of 2^(line_log - logw)
                auto local_rows = [&](int lg) { return ((size_t)1 << lg) >> sh.logw; };
                d_tops = arena.alloc<uint32_t>(top_words * (size_t)std::max(n_layers, 1));
                uint32_t* d_roots = arena.alloc<uint32_t>(8 * (size_t)sh.world * (size_t)std::max(n_layers, 1));
                while (line_log > last_log && line_log - sh.logw > FRI_SHARD_MIN_LOG) {
                    const size_t rl = local_rows(line_log);
                    if (!cur) {
                        cur = arena.alloc<uint32_t>(4 * rl);
                        ck(cudaMemsetAsync(cur, 0, 4 * rl * sizeof(uint32_t), st), "memset");
                    }
                    uint32_t* coords[4];
                    for (int k = 0; k < 4; ++k) coords[k] = cur + (size_t)k * rl;
                    while (qi < quotients.size() && quotients[qi].log - 1 == line_log) {
                        // inverse y twiddles of this rank's rows: entry i of the whole table belongs to rows 2i, 2i + 1
                        ck(fold_circle_into_line_dev(coords, quotients[qi].coords, inv_y_twiddles(tw, quotients[qi].log) + rl * sh.rank,
                                                     quotients[qi].log - sh.logw, d_alphas + li, st),
                           "fold circle");
                        ++qi;
                    }
                    FriLayer L;
                    L.log = line_log;
                    L.logw = sh.logw;
                    for (int k = 0; k < 4; ++k) L.coords[k] = coords[k];
                    std::vector<ColRef> refs;
                    for (int k = 0; k < 4; ++k) refs.push_back({coords[k], line_log - sh.logw});
                    merkle_commit(ctx, arena, refs, L.tree, /*fetch_root=*/false);
                    L.tree.logw = sh.logw;
                    L.tree.rank = sh.rank;
                    uint32_t* roots = d_roots + 8 * (size_t)sh.world * (size_t)li;
                    nck(api.AllGather(L.tree.layers[0], roots, 8, ncclUint32, sh.comm->comm, st), "ncclAllGather(layer roots)");
                    sh.comm->n_collectives++;
                    sh.comm->bytes_sent += 32;
                    sh.comm->bytes_received += 32 * (size_t)(sh.world - 1);
                    ck(channel_mix_sharded_root_draw(d_ch, roots, sh.logw, cfg.channel_variant, d_alphas + li + 1,
                                                     d_digests + 8 * (size_t)li, d_tops + top_words * (size_t)li, st),
                       "channel mix/draw (sharded layer)");
                    inner.push_back(L);
                    const size_t rl_next = local_rows(line_log - 1);
                    uint32_t* next = arena.alloc<uint32_t>(4 * rl_next);
                    uint32_t* ncoords[4];
                    for (int k = 0; k < 4; ++k) ncoords[k] = next + (size_t)k * rl_next;
                    ck(fold_line_dev(ncoords, coords, inv_x_twiddles(tw, line_log) + rl_next * sh.rank, line_log - sh.logw,
                                     d_alphas + li + 1, st),
                       "fold line");
                    cur = next;
                    --line_log;
                    ++li;
                }
                n_sharded_layers = li;
                // hand over to the replicated path: the current layer (if any fold happened) and the quotient columns that are
                // still to be folded in, all-gathered coordinate by coordinate (rank order = row order)
                if (cur) {
                    const size_t rl = local_rows(line_log);
                    gathered = arena.alloc<uint32_t>((size_t)4 << line_log);
                    nck(api.GroupStart(), "ncclGroupStart");
                    for (int k = 0; k < 4; ++k)
                        nck(api.AllGather(cur + (size_t)k * rl, gathered + ((size_t)k << line_log), rl, ncclUint32, sh.comm->comm, st),
                            "ncclAllGather(fri layer)");
                    nck(api.GroupEnd(), "ncclGroupEnd");
                    sh.comm->n_collectives++;
                    sh.comm->bytes_sent += 16 * rl;
                    sh.comm->bytes_received += 16 * rl * (size_t)(sh.world - 1);
                }
                for (size_t q = qi; q < quotients.size(); ++q) {
                    const size_t n = (size_t)1 << quotients[q].log, rl = n >> sh.logw;
                    uint32_t* f = arena.alloc<uint32_t>(4 * n);
                    nck(api.GroupStart(), "ncclGroupStart");
                    for (int k = 0; k < 4; ++k) {
                        quotients[q].full[k] = f + (size_t)k * n;
                        nck(api.AllGather(quotients[q].coords[k], quotients[q].full[k], rl, ncclUint32, sh.comm->comm, st),
                            "ncclAllGather(quotient column)");
                    }
                    nck(api.GroupEnd(), "ncclGroupEnd");
                    sh.comm->n_collectives++;
                    sh.comm->bytes_sent += 16 * rl;
                    sh.comm->bytes_received += 16 * rl * (size_t)(sh.world - 1);
                }
            }
            // all line-layer evaluations in one allocation: layer of log k (4 coordinate columns) at word offset 4 * (2^k - 1)
            uint32_t* fri_buf = arena.alloc<uint32_t>((size_t)8 << line_log);
            auto layer_at = [&](int lg) { return fri_buf + 4 * (((size_t)1 << lg) - 1); };
            uint32_t* cur = layer_at(line_log);
            if (gathered)
                ck(cudaMemcpyAsync(cur, gathered, ((size_t)4 << line_log) * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st), "copy");
            else
                ck(cudaMemsetAsync(cur, 0, ((size_t)4 << line_log) * sizeof(uint32_t), st), "memset");
            while (line_log > last_log) {
                uint32_t* coords[4];
                for (int k = 0; k < 4; ++k) coords[k] = cur + ((size_t)k << line_log);
                while (qi < quotients.size() && quotients[qi].log - 1 == line_log) {
                    uint32_t* const* src = quotients[qi].full[0] ? quotients[qi].full : quotients[qi].coords;
                    ck(fold_circle_into_line_dev(coords, src, inv_y_twiddles(tw, quotients[qi].log),
                                                 quotients[qi].log, d_alphas + li, st),
                       "fold circle");
                    ++qi;
                }
                if (LB_FRI_TAIL && line_log <= FRI_TAIL_MAX_LOG && qi == quotients.size()) {
                    // every remaining layer (tree, mix_root, draw, fold_line) in one single-CTA launch: fri_tail_kernel
                    FriTailArgs a{};
                    a.from_log = line_log;
                    a.last_log = last_log;
                    a.variant = cfg.channel_variant;
                    a.ch = d_ch;
                    a.alphas = d_alphas + li;
                    a.digests = d_digests + 8 * (size_t)li;
                    for (int lg = line_log; lg > last_log; --lg) {
                        FriLayer L;
                        L.log = lg;
                        uint32_t* base = layer_at(lg);
                        for (int k = 0; k < 4; ++k) L.coords[k] = base + ((size_t)k << lg);
                        L.tree.empty = false;
                        L.tree.max_log = lg;
                        for (int k = 0; k < 4; ++k) L.tree.sorted.push_back({L.coords[k], lg});
                        uint32_t* all = arena.alloc<uint32_t>((size_t)16 << lg);
                        L.tree.layers.assign(lg + 1, nullptr);
                        for (int k = 0; k <= lg; ++k) L.tree.layers[k] = all + 8 * (((size_t)1 << k) - 1);
                        a.vals[lg] = base;
                        a.tree[lg] = all;
                        a.itw[lg] = inv_x_twiddles(tw, lg);
                        inner.push_back(L);
                    }
                    a.vals[last_log] = layer_at(last_log);
                    ck(fri_tail(a, st), "fri tail");
                    cur = layer_at(last_log);
                    line_log = last_log;
                    break;
                }
                FriLayer L;
                L.log = line_log;
                for (int k = 0; k < 4; ++k) L.coords[k] = coords[k];
                std::vector<ColRef> refs;
                for (int k = 0; k < 4; ++k) refs.push_back({coords[k], line_log});
                merkle_commit(ctx, arena, refs, L.tree, /*fetch_root=*/false);
                ck(channel_mix_root_draw(d_ch, L.tree.layers[0], cfg.channel_variant, d_alphas + li + 1, d_digests + 8 * (size_t)li, st),
                   "channel mix/draw");
                inner.push_back(L);
                uint32_t* next = layer_at(line_log - 1);
                uint32_t* ncoords[4];
                for (int k = 0; k < 4; ++k) ncoords[k] = next + ((size_t)k << (line_log - 1));
                ck(fold_line_dev(ncoords, coords, inv_x_twiddles(tw, line_log), line_log, d_alphas + li + 1, st), "fold line");
                cur = next;
                --line_log;
                ++li;
            }
            if (n_layers > 0) {
                std::vector<uint32_t> h_digests(8 * (size_t)n_layers);
                std::vector<uint32_t> h_tops(top_words * (size_t)n_sharded_layers);
                if (n_sharded_layers)
                    ck(cudaMemcpyAsync(h_tops.data(), d_tops, h_tops.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, st), "tops d2h");
                for (int k = n_sharded_layers; k < n_layers; ++k)
                    ck(cudaMemcpyAsync(inner[k].tree.root.b, inner[k].tree.layers[0], 32, cudaMemcpyDeviceToHost, st), "root d2h");
                ck(cudaMemcpyAsync(h_digests.data(), d_digests, h_digests.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, st), "digests d2h");
                ck(cudaMemcpyAsync(&h_ch, d_ch, sizeof(h_ch), cudaMemcpyDeviceToHost, st), "channel d2h");
                ck(cudaStreamSynchronize(st), "fri loop sync");
                for (int k = 0; k < n_sharded_layers; ++k) {
                    // top levels of the row-sharded layer trees (hashed on the device): level j at word 8 * (2^j - 1)
                    MerkleTree& t = inner[k].tree;
                    t.top.assign(sh.logw + 1, {});
                    for (int j = 0; j <= sh.logw; ++j) {
                        t.top[j].resize((size_t)1 << j);
                        std::memcpy(t.top[j].data(), h_tops.data() + top_words * (size_t)k + 8 * (((size_t)1 << j) - 1), 32 * ((size_t)1 << j));
                    }
                    t.root = t.top[0][0];
                }
                for (int k = 0; k < n_layers; ++k) {
                    Hash32 d;
                    std::memcpy(d.b, h_digests.data() + 8 * (size_t)k, 32);
                    channel.log_digest(d);
                }
                Hash32 d;
                std::memcpy(d.b, h_ch.digest, 32);
                channel.adopt(d, h_ch.n_sent);
            }
            if (qi != quotients.size()) fail(LB_ERR_BAD_ARG, "fri: columns left unfolded");
            // last layer: interpolate on the host (2^(bound + blowup) values)
            size_t n = (size_t)1 << line_log;
            std::vector<uint32_t> hv(4 * n);
            ck(cudaMemcpyAsync(hv.data(), cur, 4 * n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st), "last layer d2h");
            ck(cudaStreamSynchronize(st), "last layer sync");
            last_layer_poly = interpolate_last_layer(hv, line_log, cfg.log_last_layer_degree_bound);
            channel.mix_felts(last_layer_poly);
        }
    }

    void decommit() {
        // ---- queries + decommitment plan ---------------------------------------------------------------
        int max_lde_log = quotients[0].log;
        std::vector<uint32_t> queries = generate_queries(channel, max_lde_log, cfg.n_queries);
        std::map<int, std::vector<uint32_t>> qpos;
        for (auto& q : quotients) qpos[q.log] = fold_queries(queries, max_lde_log - q.log);

        first_out.commitment = fri_first_tree.root;
        {
            std::map<int, std::vector<uint32_t>> pos_by_size;
            for (auto& q : quotients) {
                std::vector<uint32_t> cq = fold_queries(queries, max_lde_log - q.log);
                std::vector<uint32_t> positions;
                fri_positions_and_witness(q.coords, cq, g, positions, first_out.witness, q.log - sh.logw, sh.logw, sh.rank);
                pos_by_size[q.log] = positions;
            }
            merkle_decommit_plan(fri_first_tree, pos_by_size, g, first_out.decommit, sh.rank);
            first_out.decommit.queried_values.clear();
        }
        inner_out.assign(inner.size(), FriLayerOut{});
        {
            std::vector<uint32_t> lq = fold_queries(queries, 1);
            for (size_t li = 0; li < inner.size(); ++li) {
                std::vector<uint32_t> positions;
                fri_positions_and_witness(inner[li].coords, lq, g, positions, inner_out[li].witness, inner[li].log - inner[li].logw,
                                          inner[li].logw, sh.rank);
                std::map<int, std::vector<uint32_t>> m;
                m[inner[li].log] = positions;
                merkle_decommit_plan(inner[li].tree, m, g, inner_out[li].decommit, sh.rank);
                inner_out[li].decommit.queried_values.clear();
                inner_out[li].commitment = inner[li].tree.root;
                lq = fold_queries(lq, 1);
            }
        }
        tree_dec.assign(4, DecommitIdx{});
        for (int t = 0; t < 4; ++t) merkle_decommit_plan(trees[t].merkle, qpos, g, tree_dec[t], sh.rank);

        // one gather for everything
        g.values.resize(g.addrs.size());
        if (!g.addrs.empty()) {
            const uint32_t** d_addrs = arena.upload(g.addrs);
            uint32_t* d_vals = arena.alloc<uint32_t>(g.addrs.size());
            ck(gather_words(d_vals, d_addrs, (int)g.addrs.size(), st), "gather");
            if (sh.on()) {
                // every word has one owner (null addresses read as 0): the sum over the ranks is the decommitment
                nck(nccl_api().AllReduce(d_vals, d_vals, g.addrs.size(), ncclUint32, ncclSum, sh.comm->comm, st), "ncclAllReduce(decommitment)");
                sh.comm->n_collectives++;
                sh.comm->bytes_sent += g.addrs.size() * 4;
                sh.comm->bytes_received += g.addrs.size() * 4;
            }
            ck(cudaMemcpyAsync(g.values.data(), d_vals, g.values.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, st), "gather d2h");
            ck(cudaStreamSynchronize(st), "gather sync");
            for (auto& kv : g.consts) g.values[kv.first] = kv.second;
        }
    }
};

}  // namespace

// ======================================================================================
int prove_impl(lb_ctx* ctx, const lb_trace_table* tables, int n_tables, const lb_preprocessed_column* pre_in, int n_pre,
               const lb_prove_config* cfg_in, std::vector<uint8_t>& out, lb_comm* comm) {
    try {
        const lb_prove_config cfg = checked_config(cfg_in);
        if (n_tables < 1 || !tables) fail(LB_ERR_BAD_ARG, "prove: no trace tables");
        ck(cudaSetDevice(ctx->device), "set device");
        if (!ctx->kernels_ready) {
            ck(kernels_init(ctx->stream), "kernels init");
            ctx->kernels_ready = true;
        }
        ctx->transcript.clear();
        ctx->stage_ms.clear();
        ProveJob job(ctx, tables, n_tables, pre_in, n_pre, cfg, comm);
        job.run(out);
        return LB_OK;
    } catch (const ProveError& e) {
        cudaStreamSynchronize(ctx->stream);
        ctx->err = e.msg;
        return e.code;
    } catch (const std::bad_alloc&) {  // nothing may unwind through the extern "C" boundary
        cudaStreamSynchronize(ctx->stream);
        ctx->err = "prove: host allocation failed";
        return LB_ERR_OOM;
    } catch (const std::exception& e) {
        cudaStreamSynchronize(ctx->stream);
        ctx->err = std::string("prove: ") + e.what();
        return LB_ERR_BAD_ARG;
    } catch (...) {
        cudaStreamSynchronize(ctx->stream);
        ctx->err = "prove: unknown exception";
        return LB_ERR_CUDA;
    }
}

// ======================================================================================
// trait-level entry points (bound one by one by a Rust `CudaBackend`)
// ======================================================================================
namespace {
inline QM31 q_from_words(const uint32_t* w) { return q_make(w[0], w[1], w[2], w[3]); }
void ensure_kernels(lb_ctx* ctx) {
    ck(cudaSetDevice(ctx->device), "set device");
    if (!ctx->kernels_ready) {
        ck(kernels_init(ctx->stream), "kernels init");
        ctx->kernels_ready = true;
    }
}
template <class Fn>
int guarded(lb_ctx* ctx, Fn&& fn) {
    try {
        fn();
        return LB_OK;
    } catch (const ProveError& e) {
        cudaStreamSynchronize(ctx->stream);
        ctx->err = e.msg;
        return e.code;
    } catch (const std::bad_alloc&) {
        cudaStreamSynchronize(ctx->stream);
        ctx->err = "host allocation failed";
        return LB_ERR_OOM;
    } catch (const std::exception& e) {
        cudaStreamSynchronize(ctx->stream);
        ctx->err = e.what();
        return LB_ERR_BAD_ARG;
    } catch (...) {
        cudaStreamSynchronize(ctx->stream);
        ctx->err = "unknown exception";
        return LB_ERR_CUDA;
    }
}
}  // namespace

int eval_at_point_impl(lb_ctx* ctx, const uint32_t* const* h_cols, int n_cols, int log, const uint32_t point[8],
                       uint32_t* h_out) {
    return guarded(ctx, [&] {
        ensure_kernels(ctx);
        Arena arena(ctx);
        std::vector<const uint32_t*> cols(h_cols, h_cols + n_cols);
        QPt pt{q_from_words(point), q_from_words(point + 4)};
        QM31* d_out = arena.alloc<QM31>(n_cols);
        launch_eval_at_point(ctx, arena, cols, log, pt, d_out);
        ck(cudaMemcpyAsync(h_out, d_out, (size_t)n_cols * sizeof(QM31), cudaMemcpyDeviceToHost, ctx->stream), "eval d2h");
        ck(cudaStreamSynchronize(ctx->stream), "eval sync");
    });
}

int accumulate_quotients_impl(lb_ctx* ctx, int log, const uint32_t* const* h_cols, int n_cols,
                              const lb_sample_batch* batches, const lb_batch_shard* shards, int n_batches,
                              const uint32_t random_coeff[4], uint32_t* const d_out[4]) {
    return guarded(ctx, [&] {
        ensure_kernels(ctx);
        Arena arena(ctx);
        std::vector<const uint32_t*> cols(h_cols, h_cols + n_cols);
        std::vector<HostBatch> hb(n_batches);
        for (int b = 0; b < n_batches; ++b) {
            hb[b].pt = QPt{q_from_words(batches[b].point), q_from_words(batches[b].point + 4)};
            for (int k = 0; k < batches[b].n_cols; ++k)
                hb[b].cols.push_back({batches[b].col_idx[k], q_from_words(batches[b].values + 4 * k)});
            if (shards) {
                if (shards[b].col_offset < 0 || shards[b].n_cols_global < shards[b].col_offset + batches[b].n_cols)
                    fail(LB_ERR_BAD_ARG, "quotients: bad shard description");
                hb[b].col_offset = shards[b].col_offset;
                hb[b].n_cols_global = shards[b].n_cols_global;
            }
        }
        launch_quotients(ctx, arena, log, cols, hb, q_from_words(random_coeff), d_out);
        ck(cudaStreamSynchronize(ctx->stream), "quotients sync");  // arena scratch is released on return
    });
}

int fold_impl(lb_ctx* ctx, int circle, uint32_t* const d_dst[4], const uint32_t* const d_src[4], int log,
              const uint32_t alpha[4]) {
    return guarded(ctx, [&] {
        ensure_kernels(ctx);
        if (log < 1) fail(LB_ERR_BAD_ARG, "fold: log_size < 1");
        int r = lb_twiddles_ensure(ctx, circle ? log : log + 1);
        if (r) fail(r, ctx->err);
        if (circle)
            ck(fold_circle_into_line(d_dst, d_src, inv_y_twiddles(ctx->tw, log), log, q_from_words(alpha), ctx->stream), "fold circle");
        else
            ck(fold_line(d_dst, d_src, inv_x_twiddles(ctx->tw, log), log, q_from_words(alpha), ctx->stream), "fold line");
    });
}

int grind_impl(lb_ctx* ctx, const uint32_t digest[8], int variant, uint32_t pow_bits, uint64_t* nonce_out) {
    return guarded(ctx, [&] {
        ensure_kernels(ctx);
        if (pow_bits > 64) fail(LB_ERR_BAD_ARG, "grind: pow_bits > 64");
        Arena arena(ctx);
        unsigned long long* d_found = arena.alloc<unsigned long long>(1);
        uint64_t base = 0;
        const uint64_t chunk = (uint64_t)1 << 24;
        for (;;) {
            ck(cudaMemsetAsync(d_found, 0xFF, 8, ctx->stream), "grind memset");
            ck(grind_range(digest, variant, pow_bits, base, chunk, d_found, ctx->stream), "grind");
            unsigned long long found;
            ck(cudaMemcpyAsync(&found, d_found, 8, cudaMemcpyDeviceToHost, ctx->stream), "grind d2h");
            ck(cudaStreamSynchronize(ctx->stream), "grind sync");
            if (found != ~0ull) {
                *nonce_out = found;
                return;
            }
            base += chunk;
        }
    });
}

namespace {
Relations relations_from(const lb_relation* rels, int n) {
    Relations r{};
    for (int k = 0; k < n && k < REL_COUNT; ++k) r.r[k] = Relation2{q_from_words(rels[k].z), q_from_words(rels[k].alpha)};
    return r;
}
}  // namespace

int logup_impl(lb_ctx* ctx, int kind, const uint32_t* d_main, size_t main_stride, const uint32_t* const d_lut[2],
               uint32_t* d_inter, size_t inter_stride, int log, const lb_relation* rels, int n_rels, uint32_t claimed_out[4]) {
    return guarded(ctx, [&] {
        ensure_kernels(ctx);
        if (kind < 0 || kind >= COMP_KIND_COUNT || log < 1) fail(LB_ERR_BAD_ARG, "logup: bad args");
        ComponentShape sh = component_shape(kind);
        if (sh.lut && n_rels < REL_COUNT) fail(LB_ERR_BAD_ARG, "logup: component needs the LUT relations");
        PreCols pc{};
        for (int q = 0; q < sh.n_pre; ++q) {
            if (!d_lut || !d_lut[q]) fail(LB_ERR_BAD_ARG, "logup: component needs its LUT columns");
            pc.p[q] = d_lut[q];
        }
        Arena arena(ctx);
        size_t n = (size_t)1 << log;
        uint32_t* scan_tmp = arena.alloc<uint32_t>(4 * n);
        uint32_t* block_sums = arena.alloc<uint32_t>(4 * (n / 1024 + 1));
        uint32_t* d_claimed = arena.alloc<uint32_t>(4);
        Relations r = relations_from(rels, n_rels);
        ck(logup_interaction_trace(kind, d_main, main_stride, pc, d_inter, inter_stride, log, r, scan_tmp, block_sums, d_claimed,
                                   ctx->stream),
           "logup");
        ck(cudaMemcpyAsync(claimed_out, d_claimed, 16, cudaMemcpyDeviceToHost, ctx->stream), "claimed d2h");
        ck(cudaStreamSynchronize(ctx->stream), "logup sync");
    });
}

int constraint_quotients_impl(lb_ctx* ctx, int kind, const uint32_t* d_main, size_t main_stride, const uint32_t* d_inter,
                              size_t inter_stride, const uint32_t* const d_lut[2], int log_size, int eval_log,
                              const lb_relation* rels, int n_rels, const uint32_t claimed_sum[4], const uint32_t* pows,
                              int n_pows, uint32_t* const d_acc[4], int accumulate) {
    return guarded(ctx, [&] {
        ensure_kernels(ctx);
        if (kind < 0 || kind >= COMP_KIND_COUNT) fail(LB_ERR_BAD_ARG, "constraints: unknown component");
        ComponentShape sh = component_shape(kind);
        if (n_pows != sh.n_constraints) fail(LB_ERR_BAD_ARG, "constraints: wrong number of random-coefficient powers");
        if (eval_log <= log_size || eval_log > 30) fail(LB_ERR_BAD_ARG, "constraints: bad evaluation domain");
        if (sh.lut && n_rels < REL_COUNT) fail(LB_ERR_BAD_ARG, "constraints: component needs the LUT relations");
        ConstraintParams p{};
        p.main = d_main;
        p.main_stride = main_stride;
        p.inter = d_inter;
        p.inter_stride = inter_stride;
        for (int q = 0; q < sh.n_pre; ++q) {
            if (!d_lut || !d_lut[q]) fail(LB_ERR_BAD_ARG, "constraints: component needs its LUT columns");
            p.pre.p[q] = d_lut[q];
        }
        for (int k = 0; k < 4; ++k) p.acc[k] = d_acc[k];
        p.accumulate = accumulate;
        p.log_size = log_size;
        p.eval_log = eval_log;
        p.rels = relations_from(rels, n_rels);
        p.cumsum_shift = q_mul_m(q_from_words(claimed_sum), m_inv((uint32_t)(((uint64_t)1 << log_size) % P)));
        for (int k = 0; k < n_pows; ++k) p.pows[k] = q_from_words(pows + 4 * k);
        int log_expand = eval_log - log_size;
        uint32_t init = subgroup_gen(eval_log + 1), step = subgroup_gen(eval_log - 1);
        std::vector<uint32_t> dinv((size_t)1 << log_expand);
        for (uint32_t i = 0; i < (1u << log_expand); ++i) {
            Pt pt = host_index_to_point(init + step * bit_reverse(i, log_expand));
            dinv[i] = m_inv(coset_vanishing_m(log_size, pt));
        }
        Arena arena(ctx);
        p.denom_inv = arena.upload(dinv);
        ck(constraint_quotients(kind, p, ctx->stream), "constraint quotients");
        ck(cudaStreamSynchronize(ctx->stream), "constraints sync");  // arena scratch is released on return
    });
}

}  // namespace lb
