// CPU prover (oracle; TEST INFRASTRUCTURE ONLY - only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
// --impl reference legs may load this library; luminair_b200 never does).
//
// A compiled, packed (AVX-512 / AVX2) and multi-threaded (OpenMP) restatement of the reference's CPU prover:
// luminair_prover::prover::prove (/root/reference/crates/prover/src/prover.rs:28-319) over stwo's SimdBackend with the
// "parallel" feature (/root/reference/Cargo.toml:21-26) - CommitmentSchemeProver / TreeBuilder, the LogUp interaction
// trace (crates/air/src/components/*/witness.rs), stwo::prover::prove (composition polynomial, OODS sampling, DEEP
// quotients, FRI, grind, decommitment) and bincode (crates/prover/src/lib.rs:25-32).  The reference itself cannot be
// built here (no Rust toolchain; stwo / numerair / luminal un-vendored), so this is a "port": it is pinned by
// reproducing the reference's committed proof ui/demo/public/proof byte for byte (tests/test_cpu_prover.py) and by
// byte equality with the numpy oracle (oracle/prover.py) on every fixture.
#include <omp.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "air.hpp"
#include "pcs.hpp"

namespace cpu {

struct Config {
    uint32_t pow_bits = 5, log_blowup = 1, log_last = 0;
    uint64_t n_queries = 3;
    int channel_variant = 0, n_slots = 17, air_era = 0, draw_lookup = 1;
};

struct ProvingError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

struct Tree {
    std::vector<Col> polys, evals;
    MerkleTree merkle;
};

// coefficients zero-padded to n (CirclePoly::extend)
static void zero_extend(Col& dst, const Col& poly, size_t n) {
    dst.alloc_uninit(n);
    par_copy(dst.data(), poly.data(), poly.size());
    par_fill(dst.data() + poly.size(), n - poly.size(), 0);
}

static int log2_exact(size_t n) {
    int l = 0;
    while (((size_t)1 << l) < n) l++;
    if (((size_t)1 << l) != n) throw std::runtime_error("size is not a power of two");
    return l;
}

// run the CFFT over columns grouped by size
static void transform_all(std::vector<Col>& cols, bool fwd) {
    std::map<int, std::vector<uint32_t*>> by;
    for (auto& c : cols) by[log2_exact(c.size())].push_back(c.data());
    for (auto& kv : by) cfft(kv.second.data(), (int)kv.second.size(), kv.first, fwd);
}

// CommitmentTreeProver::new: evaluate every polynomial on CanonicCoset(log + blowup), Merkle-commit, mix the root
static void commit_tree(Tree& t, uint32_t log_blowup, Channel& ch) {
    t.evals.resize(t.polys.size());
    for (size_t i = 0; i < t.polys.size(); i++) {
        zero_extend(t.evals[i], t.polys[i], t.polys[i].size() << log_blowup);
    }
    transform_all(t.evals, true);
    std::vector<ColRef> refs;
    for (auto& e : t.evals) refs.push_back({e.data(), log2_exact(e.size())});
    t.merkle = merkle_commit(refs);
    ch.mix_root(t.merkle.root());
}

struct Component {
    CompCtx cx;
    int slot;
    QM claimed_sum;
    int n_constraints, n_main, n_inter, main_off, inter_off;
    std::vector<int> pre_idx;  // indices into the preprocessed tree
    int bound;                 // max_constraint_log_degree_bound
};

// storage row of the [-1] mask for every row (constraint-framework offset_bit_reversed_circle_domain_index)
static std::vector<uint32_t> prev_row_index(int domain_log, int eval_log) {
    size_t n = (size_t)1 << eval_log, half = n >> 1;
    size_t step = (size_t)1 << (eval_log - domain_log - 1);
    std::vector<uint32_t> out(n);
#pragma omp parallel for schedule(static) if (n >= 4096)
    for (size_t j = 0; j < n; j++) {
        size_t prev = bit_reverse((uint32_t)j, eval_log), res;
        if (prev < half)
            res = (prev + half - step) % half;
        else
            res = ((prev - half + step) % half) + half;
        out[j] = bit_reverse((uint32_t)res, eval_log);
    }
    return out;
}

// LogupTraceGenerator over one component: 4 * n_frac base columns (values on the trace domain) + the claimed sum
static void gen_interaction_trace(int kind, const std::vector<Col>& main, int log, const Relation* rels,
                                  const std::vector<const uint32_t*>& lut_cols, std::vector<Col>& out, QM& claimed) {
    auto terms = lookup_terms(kind);
    const size_t n = (size_t)1 << log;
    const int nt = (int)terms.size();
    size_t base = out.size();
    out.resize(base + 4 * nt);
    for (int k = 0; k < 4 * nt; k++) out[base + k].assign(n, 0);
    // fractions, running sum over the relation uses of a row (one batch inversion per chunk of rows)
    const size_t CHV = 16;
    const size_t nvec = n / W, nchunk = (nvec + CHV - 1) / CHV;
#pragma omp parallel for schedule(static) if (n >= 4096)
    for (size_t ch = 0; ch < nchunk; ch++) {
        size_t v0 = ch * CHV, v1 = std::min(nvec, v0 + CHV), nv = v1 - v0;
        VQ den[CHV * 8], pre[CHV * 8];
        VQ run = vq_set1(qm(1));
        for (size_t i = 0; i < nv; i++)
            for (int t = 0; t < nt; t++) {
                const LookupTerm& T = terms[t];
                const Relation& r = rels[T.rel];
                size_t row = (v0 + i) * W;
                VQ d = vq_zero() - vq_set1(r.z);
                for (int v = 0; v < T.nvals; v++) {
                    V x = vload((T.table ? lut_cols[T.val[v]] : main[T.val[v]].data()) + row);
                    if (v == 0)
                        d.c[0] = d.c[0] + x;
                    else
                        for (int k = 0; k < 4; k++) d.c[k] = d.c[k] + vset1(r.pw[v].c[k]) * x;
                }
                den[i * nt + t] = d;
                pre[i * nt + t] = run;
                run = run * d;
            }
        VQ inv = vq_inv(run);
        for (size_t i = nv; i-- > 0;)
            for (int t = nt; t-- > 0;) {
                VQ di = inv * pre[i * nt + t];
                inv = inv * den[i * nt + t];
                den[i * nt + t] = di;
            }
        for (size_t i = 0; i < nv; i++) {
            size_t row = (v0 + i) * W;
            VQ acc = vq_zero();
            for (int t = 0; t < nt; t++) {
                V m = vload(main[terms[t].mult].data() + row);
                if (terms[t].table) m = vneg(m);
                acc = acc + den[i * nt + t] * m;
                for (int k = 0; k < 4; k++) vstore(out[base + 4 * t + k].data() + row, acc.c[k]);
            }
        }
    }
    // claimed sum = sum of the last column; shift by claimed / n; inclusive prefix sum in canonic-coset order
    Col* last[4] = {&out[base + 4 * (nt - 1)], &out[base + 4 * (nt - 1) + 1], &out[base + 4 * (nt - 1) + 2],
                    &out[base + 4 * (nt - 1) + 3]};
    for (int k = 0; k < 4; k++) {
        uint64_t s = 0;
        const uint32_t* p = last[k]->data();
#pragma omp parallel for reduction(+ : s) schedule(static) if (n >= 4096)
        for (size_t i = 0; i < n; i++) s += p[i];
        claimed.c[k] = (uint32_t)(s % P);
    }
    QM shift = claimed * m_inv((uint32_t)(n % P));
    // canonic-coset point k lives at circle-domain natural index k/2 (k even) or n - (k+1)/2 (k odd), stored bit-reversed
    auto storage = [&](size_t k) -> size_t {
        size_t dom = (k & 1) ? n - (k + 1) / 2 : k / 2;
        return bit_reverse((uint32_t)dom, log);
    };
    const int nth = std::max(1, omp_get_max_threads());
    const size_t per = (n + nth - 1) / nth;
    for (int k = 0; k < 4; k++) {
        uint32_t* p = last[k]->data();
        uint32_t sh = shift.c[k];
        std::vector<uint32_t> part(nth + 1, 0);
#pragma omp parallel for schedule(static, 1)
        for (int t = 0; t < nth; t++) {
            uint32_t a = 0;
            for (size_t i = t * per; i < std::min(n, (t + 1) * per); i++) a = m_add(a, m_sub(p[storage(i)], sh));
            part[t + 1] = a;
        }
        for (int t = 0; t < nth; t++) part[t + 1] = m_add(part[t + 1], part[t]);
#pragma omp parallel for schedule(static, 1)
        for (int t = 0; t < nth; t++) {
            uint32_t a = part[t];
            for (size_t i = t * per; i < std::min(n, (t + 1) * per); i++) {
                size_t s = storage(i);
                a = m_add(a, m_sub(p[s], sh));
                p[s] = a;
            }
        }
    }
}

struct Prover {
    Config cfg;
    Channel ch;
    std::vector<Tree> trees;
    Relation rels[REL_COUNT];
    std::vector<Component> comps;
    std::vector<int> pre_log;  // log size of every preprocessed column (committed order)
    double stage_ms[8] = {0};

    // ComponentProver::evaluate_constraint_quotients_on_domain -> accumulates into acc (2^bound rows)
    void constraint_quotients(const Component& c, const QM* pows, QCol& acc) {
        const int eval_log = c.bound, log = c.cx.log_size;
        const size_t N = (size_t)1 << eval_log;
        std::vector<const uint32_t*> main(c.n_main), inter(c.n_inter), prep(c.pre_idx.size());
        std::vector<Col> ext;  // columns re-evaluated on the evaluation domain (need_to_extend)
        bool extend = false;
        for (int k = 0; k < c.n_main; k++) extend |= trees[1].evals[c.main_off + k].size() != N;
        for (int k = 0; k < c.n_inter; k++) extend |= trees[2].evals[c.inter_off + k].size() != N;
        for (int i : c.pre_idx) extend |= trees[0].evals[i].size() != N;
        if (extend) {
            auto add = [&](const Col& poly) {
                ext.emplace_back();
                zero_extend(ext.back(), poly, N);
            };
            for (int k = 0; k < c.n_main; k++) add(trees[1].polys[c.main_off + k]);
            for (int k = 0; k < c.n_inter; k++) add(trees[2].polys[c.inter_off + k]);
            for (int i : c.pre_idx) add(trees[0].polys[i]);
            transform_all(ext, true);
            size_t o = 0;
            for (int k = 0; k < c.n_main; k++) main[k] = ext[o++].data();
            for (int k = 0; k < c.n_inter; k++) inter[k] = ext[o++].data();
            for (size_t k = 0; k < c.pre_idx.size(); k++) prep[k] = ext[o++].data();
        } else {
            for (int k = 0; k < c.n_main; k++) main[k] = trees[1].evals[c.main_off + k].data();
            for (int k = 0; k < c.n_inter; k++) inter[k] = trees[2].evals[c.inter_off + k].data();
            for (size_t k = 0; k < c.pre_idx.size(); k++) prep[k] = trees[0].evals[c.pre_idx[k]].data();
        }
        // 1 / coset_vanishing(trace coset) takes 2^log_expand values, constant over runs of 2^log rows
        const int log_expand = eval_log - log;
        std::vector<uint32_t> dinv((size_t)1 << log_expand);
        Coset trace = Coset::odds(log), half = Coset::half_odds(eval_log - 1);
        for (size_t i = 0; i < dinv.size(); i++) {
            size_t nat = bit_reverse((uint32_t)i, log_expand);
            Pt p = half.at(nat);  // eval_domain.at(nat), nat < 2^(eval_log - 1)
            dinv[i] = m_inv(coset_vanishing(trace, p));
        }
        auto prev_idx = prev_row_index(log, eval_log);
        QM shift = c.claimed_sum * m_inv((uint32_t)(((uint64_t)1 << log) % P));
        uint32_t* ac[4] = {acc.c[0].data(), acc.c[1].data(), acc.c[2].data(), acc.c[3].data()};
#pragma omp parallel for schedule(static)
        for (size_t row = 0; row < N; row += W) {
            DomainEval ev;
            ev.main = main.data();
            ev.inter = inter.data();
            ev.prep = prep.data();
            ev.prev_idx = prev_idx.data();
            ev.row = row;
            ev.pows = pows;
            ev.shift = shift;
            ev.row_res = vq_zero();
            evaluate_component(ev, c.cx);
            VQ r = ev.row_res * vset1(dinv[row >> log]);
            vq_store(ac, row, vq_load(ac, row) + r);
        }
    }

    // ComponentProvers::compute_composition_polynomial + DomainEvaluationAccumulator::finalize
    std::vector<Col> composition(QM random_coeff) {
        int total = 0;
        for (auto& c : comps) total += c.n_constraints;
        std::vector<QM> powers(total);  // powers[k] = r^(total-1-k): the first constraint takes the highest power
        QM cur = qm(1);
        for (int k = total - 1; k >= 0; k--) {
            powers[k] = cur;
            cur = cur * random_coeff;
        }
        std::map<int, QCol> sub;
        int off = 0, max_log = 0;
        for (auto& c : comps) {
            max_log = std::max(max_log, c.bound);
            auto it = sub.find(c.bound);
            if (it == sub.end()) {
                sub[c.bound].alloc(c.bound);
                it = sub.find(c.bound);
            }
            constraint_quotients(c, powers.data() + off, it->second);
            off += c.n_constraints;
        }
        std::vector<Col> curp;  // coefficients of the running sum
        for (auto& kv : sub) {
            QCol& vals = kv.second;
            if (!curp.empty()) {
                std::vector<Col> lifted(4);
                for (int k = 0; k < 4; k++) {
                    zero_extend(lifted[k], curp[k], (size_t)1 << kv.first);
                }
                transform_all(lifted, true);
                for (int k = 0; k < 4; k++) {
                    uint32_t* a = vals.c[k].data();
                    const uint32_t* b = lifted[k].data();
                    size_t n = vals.c[k].size();
#pragma omp parallel for schedule(static) if (n >= 4096)
                    for (size_t i = 0; i < n; i++) a[i] = m_add(a[i], b[i]);
                }
            }
            curp.assign(4, Col());
            for (int k = 0; k < 4; k++) curp[k] = std::move(vals.c[k]);
            transform_all(curp, false);
        }
        if (curp[0].size() != ((size_t)1 << max_log)) throw std::runtime_error("composition log size");
        return curp;
    }

    // stwo::prover::prove: returns the StarkProof parts through `w`
    struct StarkParts {
        std::vector<std::vector<std::vector<QM>>> sampled;  // tree -> column -> samples
        std::vector<Decommitment> decommitments;
        std::vector<std::vector<uint32_t>> queried;
        uint64_t nonce;
        FriProof fri;
    };

    static double now_ms() {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
    }

    StarkParts stark_prove() {
        double t0 = now_ms();
        QM random_coeff = ch.draw_secure_felt();
        {
            Tree t;
            t.polys = composition(random_coeff);
            trees.push_back(std::move(t));
            commit_tree(trees.back(), cfg.log_blowup, ch);
        }
        double t1 = now_ms();
        stage_ms[2] = t1 - t0;
        // OODS point (CirclePoint::get_random_point)
        QM t = ch.draw_secure_felt();
        QM t2 = t * t;
        QM inv = qm_inv(t2 + qm(1));
        QPt oods{(qm(1) - t2) * inv, (t + t) * inv};
        const size_t nt = trees.size();
        std::vector<std::vector<std::vector<QPt>>> pts(nt);
        pts[0].resize(trees[0].polys.size());
        pts[1].resize(trees[1].polys.size());
        pts[2].resize(trees[2].polys.size());
        for (auto& c : comps) {
            for (int i : c.pre_idx)
                if (pts[0][i].empty()) pts[0][i] = {oods};
            for (int k = 0; k < c.n_main; k++) pts[1][c.main_off + k] = {oods};
            Pt sp = index_to_point((0u - Coset::odds(c.cx.log_size).step()) & IDX_MASK);
            QPt prev = qpt_add(oods, QPt{qm(sp.x), qm(sp.y)});
            for (int k = 0; k < c.n_inter; k++) {
                if (k >= c.n_inter - 4)
                    pts[2][c.inter_off + k] = {prev, oods};
                else
                    pts[2][c.inter_off + k] = {oods};
            }
        }
        pts[nt - 1].assign(4, std::vector<QPt>{oods});
        // prove_values: sample every polynomial at its points (batched by size and point)
        StarkParts sp;
        sp.sampled.resize(nt);
        {
            struct Job {
                int n;
                QPt pt;
                std::vector<const uint32_t*> cols;
                std::vector<QM*> dst;
            };
            std::map<std::pair<int, std::array<uint32_t, 8>>, Job> jobs;
            for (size_t tr = 0; tr < nt; tr++) {
                sp.sampled[tr].resize(trees[tr].polys.size());
                for (size_t c = 0; c < trees[tr].polys.size(); c++) sp.sampled[tr][c].resize(pts[tr][c].size());
            }
            for (size_t tr = 0; tr < nt; tr++)
                for (size_t c = 0; c < trees[tr].polys.size(); c++)
                    for (size_t s = 0; s < pts[tr][c].size(); s++) {
                        int n = log2_exact(trees[tr].polys[c].size());
                        Job& j = jobs[{n, pt_key(pts[tr][c][s])}];
                        j.n = n;
                        j.pt = pts[tr][c][s];
                        j.cols.push_back(trees[tr].polys[c].data());
                        j.dst.push_back(&sp.sampled[tr][c][s]);
                    }
            for (auto& kv : jobs) {
                Job& j = kv.second;
                std::vector<QM> res(j.cols.size());
                eval_at_point(j.cols.data(), (int)j.cols.size(), j.n, j.pt, res.data());
                for (size_t i = 0; i < res.size(); i++) *j.dst[i] = res[i];
            }
        }
        {
            std::vector<QM> flat;
            for (auto& tv : sp.sampled)
                for (auto& col : tv)
                    for (auto& v : col) flat.push_back(v);
            ch.mix_felts(flat.data(), flat.size());
        }
        double t2m = now_ms();
        stage_ms[3] = t2m - t1;
        QM rc = ch.draw_secure_felt();
        // compute_fri_quotients: all committed columns, grouped by size (descending)
        std::vector<QCol> quotients;
        {
            struct Ref {
                const Col* e;
                std::vector<Sample> s;
            };
            std::vector<Ref> flat;
            for (size_t tr = 0; tr < nt; tr++)
                for (size_t c = 0; c < trees[tr].evals.size(); c++) {
                    Ref r;
                    r.e = &trees[tr].evals[c];
                    for (size_t s = 0; s < pts[tr][c].size(); s++) r.s.push_back({pts[tr][c][s], sp.sampled[tr][c][s]});
                    flat.push_back(std::move(r));
                }
            std::stable_sort(flat.begin(), flat.end(), [](const Ref& a, const Ref& b) { return a.e->size() > b.e->size(); });
            size_t i = 0;
            while (i < flat.size()) {
                size_t n = flat[i].e->size(), j = i;
                std::vector<const uint32_t*> cols;
                std::vector<std::vector<Sample>> samples;
                while (j < flat.size() && flat[j].e->size() == n) {
                    cols.push_back(flat[j].e->data());
                    samples.push_back(flat[j].s);
                    j++;
                }
                quotients.emplace_back();
                accumulate_quotients(log2_exact(n), cols, samples, rc, quotients.back());
                i = j;
            }
        }
        double t3 = now_ms();
        stage_ms[4] = t3 - t2m;
        FriProver fri;
        int max_log = quotients[0].log;
        std::vector<int> qlogs;
        for (auto& q : quotients) qlogs.push_back(q.log);
        fri.commit(ch, cfg.log_blowup, cfg.log_last, std::move(quotients));
        double t4 = now_ms();
        stage_ms[5] = t4 - t3;
        sp.nonce = grind(ch, cfg.pow_bits);
        ch.mix_u64(sp.nonce);
        auto queries = generate_queries(ch, max_log, cfg.n_queries);
        sp.fri = fri.decommit(queries);
        // query positions per log size: every size any *committed column* has
        std::map<int, std::vector<size_t>> qpos;
        for (int l : qlogs) qpos[l] = fold_queries(queries, max_log - l);
        for (auto& tr : trees) {
            sp.queried.emplace_back();
            sp.decommitments.emplace_back();
            merkle_decommit(tr.merkle, qpos, sp.queried.back(), sp.decommitments.back());
        }
        double t5 = now_ms();
        stage_ms[6] = t5 - t4;
        // OODS check: the composition polynomial's sampled value equals the constraints evaluated on the samples
        QM comp_e[4] = {sp.sampled[nt - 1][0][0], sp.sampled[nt - 1][1][0], sp.sampled[nt - 1][2][0], sp.sampled[nt - 1][3][0]};
        QM composition_oods = from_partial_evals(comp_e);
        QM acc = qm(0);
        for (auto& c : comps) {
            PointEval ev;
            ev.main = sp.sampled[1].data() + c.main_off;
            ev.inter = sp.sampled[2].data() + c.inter_off;
            std::vector<const std::vector<QM>*> pp;
            for (int i : c.pre_idx) pp.push_back(&sp.sampled[0][i]);
            ev.prep = pp.data();
            ev.denom_inverse = qm_inv(coset_vanishing(Coset::odds(c.cx.log_size), oods));
            ev.shift = c.claimed_sum * m_inv((uint32_t)(((uint64_t)1 << c.cx.log_size) % P));
            ev.r = random_coeff;
            ev.acc = &acc;
            evaluate_component(ev, c.cx);
        }
        if (composition_oods != acc) throw ProvingError("ConstraintsNotSatisfied");
        stage_ms[7] = now_ms() - t5;
        return sp;
    }
};

struct Writer {
    std::vector<uint8_t> b;
    void u8(uint8_t v) { b.push_back(v); }
    void u32(uint32_t v) { raw(&v, 4); }
    void u64(uint64_t v) { raw(&v, 8); }
    void raw(const void* p, size_t n) {
        const uint8_t* q = (const uint8_t*)p;
        b.insert(b.end(), q, q + n);
    }
    void felt(const QM& q) { raw(q.c, 16); }
    void decommitment(const Decommitment& d) {
        u64(d.hash_witness.size());
        for (auto& h : d.hash_witness) raw(h.data(), 32);
        u64(d.column_witness.size());
        for (uint32_t v : d.column_witness) u32(v);
    }
    void fri_layer(const FriLayerProof& l) {
        u64(l.fri_witness.size());
        for (auto& q : l.fri_witness) felt(q);
        decommitment(l.decommitment);
        raw(l.commitment.data(), 32);
    }
};

struct TableIn {
    int slot, n_cols;
    uint64_t n_rows;
    const uint32_t* rows;
};
struct LutIn {
    int lut, col_index, log_size;
    const uint32_t* values;
};

static std::vector<uint8_t> prove(const TableIn* tables, int n_tables, const LutIn* luts, int n_luts, const Config& cfg,
                                  double* stage_ms) {
    Prover pr;
    pr.cfg = cfg;
    pr.ch.variant = cfg.channel_variant;
    double t0 = Prover::now_ms();
    // phase 0: preprocessed trace (prover.rs:52-59); PreProcessedTrace::new sorts by log size, descending (stable)
    std::vector<LutIn> pre(luts, luts + n_luts);
    std::stable_sort(pre.begin(), pre.end(), [](const LutIn& a, const LutIn& b) { return a.log_size > b.log_size; });
    pr.trees.emplace_back();
    for (auto& l : pre) {
        pr.trees[0].polys.emplace_back(l.values, l.values + ((size_t)1 << l.log_size));
        pr.pre_log.push_back(l.log_size);
    }
    transform_all(pr.trees[0].polys, false);
    commit_tree(pr.trees[0], cfg.log_blowup, pr.ch);
    // phase 1: main trace (prover.rs:61-179): pad every table to max(next_pow2(rows), 16) with its padding row
    std::vector<int> claim(cfg.n_slots, -1);
    std::vector<std::vector<Col>> mains(cfg.n_slots);
    pr.trees.emplace_back();
    for (int ti = 0; ti < n_tables; ti++) {
        const TableIn& t = tables[ti];
        if (t.slot < 0 || t.slot >= cfg.n_slots || t.slot >= K_COUNT) throw std::invalid_argument("bad slot");
        const CompInfo& ci = comp_info(t.slot);
        if (t.n_cols != ci.n_main) throw std::invalid_argument("table width does not match the component");
        if (t.n_rows == 0) throw std::invalid_argument("TraceError::EmptyTrace");
        size_t size = 16;
        while (size < t.n_rows) size <<= 1;
        int log = log2_exact(size);
        std::vector<Col> cols(ci.n_main);
        std::vector<uint32_t> pad(ci.n_main, 0);
        for (auto& kv : ci.padding) pad[kv.first] = kv.second;
        for (int c = 0; c < ci.n_main; c++) cols[c].assign(size, pad[c]);
#pragma omp parallel for schedule(static) if (t.n_rows >= 4096)
        for (size_t r = 0; r < t.n_rows; r++)
            for (int c = 0; c < ci.n_main; c++) cols[c][r] = t.rows[r * ci.n_main + c];
        claim[t.slot] = log;
        mains[t.slot] = cols;
        for (auto& c : cols) pr.trees[1].polys.push_back(std::move(c));
    }
    transform_all(pr.trees[1].polys, false);
    for (int c : claim)
        if (c >= 0) pr.ch.mix_u64((uint64_t)c);
    commit_tree(pr.trees[1], cfg.log_blowup, pr.ch);
    double t1 = Prover::now_ms();
    stage_ms[0] = t1 - t0;
    // phase 2: interaction trace (prover.rs:181-298)
    {
        QM za[2];
        pr.ch.draw_secure_felts(za, 2);
        pr.rels[REL_NODE].set(za[0], za[1]);
        if (cfg.draw_lookup)
            for (int r = REL_SIN; r <= REL_RANGE_CHECK; r++) {
                pr.ch.draw_secure_felts(za, 2);
                pr.rels[r].set(za[0], za[1]);
            }
    }
    auto pre_index = [&](int lut, int col) {
        for (size_t i = 0; i < pre.size(); i++)
            if (pre[i].lut == lut && pre[i].col_index == col) return (int)i;
        throw std::invalid_argument("missing preprocessed column");
    };
    std::vector<QM> iclaim(cfg.n_slots);
    pr.trees.emplace_back();
    for (int slot = 0; slot < cfg.n_slots; slot++) {
        if (claim[slot] < 0) continue;
        const CompInfo& ci = comp_info(slot);
        std::vector<const uint32_t*> lut_cols;
        for (int k = 0; k < ci.n_lut_cols; k++) lut_cols.push_back(pre[pre_index(ci.lut, k)].values);
        gen_interaction_trace(slot, mains[slot], claim[slot], pr.rels, lut_cols, pr.trees[2].polys, iclaim[slot]);
    }
    transform_all(pr.trees[2].polys, false);
    for (int slot = 0; slot < cfg.n_slots; slot++)
        if (claim[slot] >= 0) pr.ch.mix_felts(&iclaim[slot], 1);
    commit_tree(pr.trees[2], cfg.log_blowup, pr.ch);
    stage_ms[1] = Prover::now_ms() - t1;
    // LuminairComponents::new (crates/air/src/components/mod.rs:261-527): claim-slot order, consecutive column spans
    {
        int lut_log[REL_COUNT] = {0};
        for (auto& l : pre) lut_log[l.lut] = l.log_size;
        int mo = 0, io = 0;
        for (int slot = 0; slot < cfg.n_slots; slot++) {
            if (claim[slot] < 0) continue;
            const CompInfo& ci = comp_info(slot);
            Component c;
            c.slot = slot;
            c.cx = CompCtx{slot, claim[slot], pr.rels, (cfg.air_era == 1 && slot == K_MUL) ? 1 : 0};
            c.claimed_sum = iclaim[slot];
            InfoEval info;
            evaluate_component(info, c.cx);
            c.n_constraints = info.n_constraints;
            c.n_main = info.n_main;
            c.n_inter = 4 * info.n_ext;
            c.main_off = mo;
            c.inter_off = io;
            mo += c.n_main;
            io += c.n_inter;
            for (int k = 0; k < ci.n_lut_cols; k++) c.pre_idx.push_back(pre_index(ci.lut, k));
            bool consumer = ci.lut >= 0 && ci.n_lut_cols == 0;
            c.bound = (consumer ? std::max(claim[slot], lut_log[ci.lut]) : claim[slot]) + 1;
            pr.comps.push_back(c);
        }
    }
    auto sp = pr.stark_prove();
    for (int k = 0; k < 8; k++)
        if (k >= 2) stage_ms[k] = pr.stage_ms[k];
    // bincode LuminairProof (crates/prover/src/lib.rs:15-32)
    Writer w;
    for (int c : claim) {
        if (c < 0)
            w.u8(0);
        else {
            w.u8(1);
            w.u32((uint32_t)c);
        }
    }
    for (int s = 0; s < cfg.n_slots; s++) {
        if (claim[s] < 0)
            w.u8(0);
        else {
            w.u8(1);
            w.felt(iclaim[s]);
        }
    }
    w.u32(cfg.pow_bits);
    w.u32(cfg.log_blowup);
    w.u32(cfg.log_last);
    w.u64(cfg.n_queries);
    w.u64(pr.trees.size());
    for (auto& t : pr.trees) {
        Hash r = t.merkle.root();
        w.raw(r.data(), 32);
    }
    w.u64(sp.sampled.size());
    for (auto& tv : sp.sampled) {
        w.u64(tv.size());
        for (auto& col : tv) {
            w.u64(col.size());
            for (auto& q : col) w.felt(q);
        }
    }
    w.u64(sp.decommitments.size());
    for (auto& d : sp.decommitments) w.decommitment(d);
    w.u64(sp.queried.size());
    for (auto& qv : sp.queried) {
        w.u64(qv.size());
        for (uint32_t v : qv) w.u32(v);
    }
    w.u64(sp.nonce);
    w.fri_layer(sp.fri.first);
    w.u64(sp.fri.inner.size());
    for (auto& l : sp.fri.inner) w.fri_layer(l);
    w.u64(sp.fri.last_layer_poly.size());
    for (auto& q : sp.fri.last_layer_poly) w.felt(q);
    w.u32(sp.fri.last_log);
    return std::move(w.b);
}

}  // namespace cpu

// ---- C ABI (ctypes: oracle/cpu_prover.py) -----------------------------------------------------------------------------------
extern "C" {
typedef struct {
    int slot, n_cols;
    uint64_t n_rows;
    const uint32_t* rows;
} ocp_table;
typedef struct {
    int lut, col_index, log_size;
    const uint32_t* values;
} ocp_lut_column;
typedef struct {
    uint32_t pow_bits, log_blowup_factor, log_last_layer_degree_bound;
    uint64_t n_queries;
    int channel_variant, n_slots, air_era, draw_lookup_elements;
} ocp_config;

static thread_local std::string g_err;
const char* ocp_last_error(void) { return g_err.c_str(); }
int ocp_lanes(void) { return cpu::W; }
int ocp_max_threads(void) { return omp_get_max_threads(); }
void ocp_set_threads(int n) {
    if (n > 0) omp_set_num_threads(n);
}

// 0 ok; -3 bad argument; -5 ProvingError::ConstraintsNotSatisfied; -1 anything else.  stage_ms: 8 doubles (may be NULL)
int ocp_prove(const ocp_table* tables, int n_tables, const ocp_lut_column* luts, int n_luts, const ocp_config* cfg,
              uint8_t** out, size_t* out_len, double* stage_ms) {
    try {
        cpu::Config c;
        if (cfg) {
            c.pow_bits = cfg->pow_bits;
            c.log_blowup = cfg->log_blowup_factor;
            c.log_last = cfg->log_last_layer_degree_bound;
            c.n_queries = cfg->n_queries;
            c.channel_variant = cfg->channel_variant;
            c.n_slots = cfg->n_slots;
            c.air_era = cfg->air_era;
            c.draw_lookup = cfg->draw_lookup_elements;
        }
        std::vector<cpu::TableIn> t(n_tables);
        for (int i = 0; i < n_tables; i++) t[i] = {tables[i].slot, tables[i].n_cols, tables[i].n_rows, tables[i].rows};
        std::vector<cpu::LutIn> l(n_luts);
        for (int i = 0; i < n_luts; i++) l[i] = {luts[i].lut, luts[i].col_index, luts[i].log_size, luts[i].values};
        double ms[8] = {0};
        auto bytes = cpu::prove(t.data(), n_tables, l.data(), n_luts, c, ms);
        if (stage_ms) memcpy(stage_ms, ms, sizeof ms);
        *out = (uint8_t*)malloc(bytes.size());
        memcpy(*out, bytes.data(), bytes.size());
        *out_len = bytes.size();
        return 0;
    } catch (const cpu::ProvingError& e) {
        g_err = e.what();
        return -5;
    } catch (const std::invalid_argument& e) {
        g_err = e.what();
        return -3;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}
void ocp_free(void* p) { free(p); }

// unit-level entry points (parity with the numpy oracle; CPU baseline of the CFFT bench)
void ocp_cfft(uint32_t* cols, size_t stride, int n_cols, int log_size, int forward) {
    std::vector<uint32_t*> p(n_cols);
    for (int c = 0; c < n_cols; c++) p[c] = cols + (size_t)c * stride;
    cpu::cfft(p.data(), n_cols, log_size, forward != 0);
}
void ocp_merkle_root(const uint32_t* const* cols, const int* logs, int n_cols, uint8_t out[32]) {
    std::vector<cpu::ColRef> r;
    for (int c = 0; c < n_cols; c++) r.push_back({cols[c], logs[c]});
    auto t = cpu::merkle_commit(r);
    auto h = t.root();
    memcpy(out, h.data(), 32);
}
void ocp_eval_at_point(const uint32_t* const* cols, int n_cols, int log_size, const uint32_t point[8], uint32_t* out) {
    cpu::QPt pt{cpu::QM{{point[0], point[1], point[2], point[3]}}, cpu::QM{{point[4], point[5], point[6], point[7]}}};
    std::vector<cpu::QM> res(n_cols);
    cpu::eval_at_point(cols, n_cols, log_size, pt, res.data());
    memcpy(out, res.data(), 16 * (size_t)n_cols);
}
}
