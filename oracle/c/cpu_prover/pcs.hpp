// CPU prover (oracle; TEST INFRASTRUCTURE ONLY).
//
// DEEP quotients and FRI.  Restates stwo core/pcs/quotients.rs (ColumnSampleBatch, quotient_constants,
// accumulate_row_quotients), prover/backend/{cpu,simd}/quotients.rs, core/fri.rs + prover/fri.rs (FriProver::commit /
// decommit, fold_line, fold_circle_into_line, Queries) and prover/line.rs (LineEvaluation::interpolate) @0790eba
// (un-vendored); reached from /root/reference/crates/prover/src/prover.rs:311-312.
#pragma once
#include <algorithm>
#include <map>
#include <tuple>

#include "buf.hpp"
#include "cfft.hpp"
#include "merkle.hpp"

namespace cpu {

struct QCol {  // SecureColumnByCoords
    Col c[4];
    int log = 0;
    void alloc(int l) {
        log = l;
        for (auto& x : c) x.assign((size_t)1 << l, 0);
    }
    QM at(size_t i) const { return QM{{c[0][i], c[1][i], c[2][i], c[3][i]}}; }
};

struct Sample {
    QPt point;
    QM value;
};
struct Batch {
    QPt point;
    std::vector<std::pair<int, QM>> cols;  // (column index within the size class, sampled value)
};
static inline std::array<uint32_t, 8> pt_key(const QPt& p) {
    return {p.x.c[0], p.x.c[1], p.x.c[2], p.x.c[3], p.y.c[0], p.y.c[1], p.y.c[2], p.y.c[3]};
}
// ColumnSampleBatch::new_vec: grouped by point; stwo keeps them in a BTreeMap keyed by the point (x then y, each QM31
// compared as its four u32 coordinates)
static inline std::vector<Batch> column_sample_batches(const std::vector<std::vector<Sample>>& samples) {
    std::map<std::array<uint32_t, 8>, Batch> g;
    for (size_t ci = 0; ci < samples.size(); ci++)
        for (auto& s : samples[ci]) {
            auto& b = g[pt_key(s.point)];
            b.point = s.point;
            b.cols.push_back({(int)ci, s.value});
        }
    std::vector<Batch> out;
    for (auto& kv : g) out.push_back(std::move(kv.second));
    return out;
}

// QuotientOps::accumulate_quotients for the columns of one log size
static inline void accumulate_quotients(int log, const std::vector<const uint32_t*>& cols,
                                        const std::vector<std::vector<Sample>>& samples, QM random_coeff, QCol& out) {
    auto dom = get_domain(log);
    const uint32_t *hx = dom->hx.data(), *hy = dom->hy();
    auto batches = column_sample_batches(samples);
    const int nb = (int)batches.size();
    // quotient_constants: per (batch, column) c_j (scaled by alpha^(j+1)); the a_j y + b_j terms summed per batch
    std::vector<std::vector<QM>> cj(nb);
    std::vector<QM> asum(nb), bsum(nb), brc(nb);
    for (int b = 0; b < nb; b++) {
        QM alpha = qm(1);
        asum[b] = bsum[b] = qm(0);
        const QPt& pt = batches[b].point;
        for (auto& cv : batches[b].cols) {
            alpha = alpha * random_coeff;
            QM a = complex_conjugate(cv.second) - cv.second;
            QM c = complex_conjugate(pt.y) - pt.y;
            QM bb = cv.second * c - a * pt.y;
            cj[b].push_back(alpha * c);
            asum[b] = asum[b] + alpha * a;
            bsum[b] = bsum[b] + alpha * bb;
        }
        brc[b] = qm_pow(random_coeff, batches[b].cols.size());
    }
    out.alloc(log);
    const size_t N = (size_t)1 << log;
    uint32_t* oc[4] = {out.c[0].data(), out.c[1].data(), out.c[2].data(), out.c[3].data()};
    if (N < (size_t)2 * W) {  // tiny domains: scalar
        for (size_t r = 0; r < N; r++) {
            uint32_t x = hx[r >> 1], y = (r & 1) ? m_neg(hy[r >> 1]) : hy[r >> 1];
            QM acc = qm(0);
            for (int b = 0; b < nb; b++) {
                const QPt& pt = batches[b].point;
                CM prx = pt.x.lo(), pix = pt.x.hi(), pry = pt.y.lo(), piy = pt.y.hi();
                CM d = CM{m_sub(prx.a, x), prx.b} * piy - CM{m_sub(pry.a, y), pry.b} * pix;
                QM num = qm(0) - (asum[b] * y + bsum[b]);
                for (size_t j = 0; j < cj[b].size(); j++) num = num + cj[b][j] * cols[batches[b].cols[j].first][r];
                QM q = mul_cm(num, cm_inv(d));
                acc = b == 0 ? q : acc * brc[b] + q;
            }
            for (int k = 0; k < 4; k++) oc[k][r] = acc.c[k];
        }
        return;
    }
    const size_t CHV = 32;  // vectors per batch-inversion chunk
    const size_t nvec = N / W, nchunk = (nvec + CHV - 1) / CHV;
#pragma omp parallel for schedule(static)
    for (size_t ch = 0; ch < nchunk; ch++) {
        size_t v0 = ch * CHV, v1 = std::min(nvec, v0 + CHV), nv = v1 - v0;
        VC den[CHV], pre[CHV];
        VQ acc[CHV];
        for (int b = 0; b < nb; b++) {
            const QPt& pt = batches[b].point;
            VC prx = vc_set1(pt.x.lo()), pix = vc_set1(pt.x.hi()), pry = vc_set1(pt.y.lo()), piy = vc_set1(pt.y.hi());
            VC run = {vset1(1), vzero()};
            for (size_t i = 0; i < nv; i++) {
                size_t r = (v0 + i) * W;
                // domain point of each row: (hx[r >> 1], r odd ? -hy[r >> 1] : hy[r >> 1])
                uint32_t xx[W], yy[W];
                for (int l = 0; l < W; l++) {
                    uint32_t t = hy[(r + l) >> 1];
                    xx[l] = hx[(r + l) >> 1];
                    yy[l] = (l & 1) ? m_neg(t) : t;
                }
                V x = vload(xx), y = vload(yy);
                VC d = VC{prx.a - x, prx.b} * piy - VC{pry.a - y, pry.b} * pix;
                den[i] = d;
                pre[i] = run;
                run = run * d;
            }
            VC inv = vc_inv(run);
            for (size_t i = nv; i-- > 0;) {
                VC di = inv * pre[i];
                inv = inv * den[i];
                den[i] = di;
            }
            VQ as = vq_set1(asum[b]), bs = vq_set1(bsum[b]), br = vq_set1(brc[b]);
            for (size_t i = 0; i < nv; i++) {
                size_t r = (v0 + i) * W;
                uint32_t yy[W];
                for (int l = 0; l < W; l++) {
                    uint32_t t = hy[(r + l) >> 1];
                    yy[l] = (l & 1) ? m_neg(t) : t;
                }
                V y = vload(yy);
                VQ num = vq_zero() - (as * y + bs);
                for (size_t j = 0; j < cj[b].size(); j++) {
                    V f = vload(cols[batches[b].cols[j].first] + r);
                    const QM& c = cj[b][j];
                    for (int k = 0; k < 4; k++) num.c[k] = num.c[k] + vset1(c.c[k]) * f;
                }
                VQ q = vq_mul_cm(num, den[i]);
                acc[i] = b == 0 ? q : acc[i] * br + q;
            }
        }
        for (size_t i = 0; i < nv; i++) vq_store(oc, (v0 + i) * W, acc[i]);
    }
}

// ---- FRI ------------------------------------------------------------------------------------------------------------
// fold_circle_into_line: dst = dst * alpha^2 + (f(p) + f(-p)) + alpha * (f(p) - f(-p)) / p.y
static inline void fold_circle_into_line(QCol& dst, const QCol& src, QM alpha) {
    auto dom = get_domain(src.log);
    const uint32_t* yinv = dom->itw[0].data();
    const size_t H = (size_t)1 << (src.log - 1);
    QM a2 = alpha * alpha;
    if (H >= (size_t)W) {
        const uint32_t* sc[4] = {src.c[0].data(), src.c[1].data(), src.c[2].data(), src.c[3].data()};
        uint32_t* dc[4] = {dst.c[0].data(), dst.c[1].data(), dst.c[2].data(), dst.c[3].data()};
        VQ va = vq_set1(alpha), va2 = vq_set1(a2);
#pragma omp parallel for schedule(static) if (H >= 4096)
        for (size_t k = 0; k < H; k += W) {
            VQ fp, fn;
            for (int c = 0; c < 4; c++) vdeinterleave(vload(sc[c] + 2 * k), vload(sc[c] + 2 * k + W), fp.c[c], fn.c[c]);
            VQ f0 = fp + fn, f1 = (fp - fn) * vload(yinv + k);
            VQ d = vq_load(dc, k) * va2 + (va * f1 + f0);
            vq_store(dc, k, d);
        }
    } else {
        for (size_t k = 0; k < H; k++) {
            QM fp = src.at(2 * k), fn = src.at(2 * k + 1);
            QM f0 = fp + fn, f1 = (fp - fn) * yinv[k];
            QM d = dst.at(k) * a2 + (alpha * f1 + f0);
            for (int c = 0; c < 4; c++) dst.c[c][k] = d.c[c];
        }
    }
}
// fold_line: values on LineDomain(half_odds(log)) -> half size; f0 + alpha * f1, f1 = (f(x) - f(-x)) / x
static inline QCol fold_line(const QCol& src, QM alpha) {
    QCol dst;
    dst.alloc(src.log - 1);
    const size_t H = (size_t)1 << (src.log - 1);
    auto dom = get_domain(src.log + 1);  // its half coset is half_odds(log): x of storage row 2k = tw[1][k]
    const uint32_t* xinv = dom->itw[1].data();
    if (H >= (size_t)W) {
        const uint32_t* sc[4] = {src.c[0].data(), src.c[1].data(), src.c[2].data(), src.c[3].data()};
        uint32_t* dc[4] = {dst.c[0].data(), dst.c[1].data(), dst.c[2].data(), dst.c[3].data()};
        VQ va = vq_set1(alpha);
#pragma omp parallel for schedule(static) if (H >= 4096)
        for (size_t k = 0; k < H; k += W) {
            VQ fx, fn;
            for (int c = 0; c < 4; c++) vdeinterleave(vload(sc[c] + 2 * k), vload(sc[c] + 2 * k + W), fx.c[c], fn.c[c]);
            VQ f0 = fx + fn, f1 = (fx - fn) * vload(xinv + k);
            vq_store(dc, k, f0 + va * f1);
        }
    } else {
        for (size_t k = 0; k < H; k++) {
            QM fx = src.at(2 * k), fn = src.at(2 * k + 1);
            QM d = (fx + fn) + alpha * ((fx - fn) * xinv[k]);
            for (int c = 0; c < 4; c++) dst.c[c][k] = d.c[c];
        }
    }
    return dst;
}

// LineEvaluation::interpolate + into_ordered_coefficients on a (small) last layer; coset = the line domain's coset
static inline std::vector<QM> line_interpolate(const QCol& v, Coset coset) {
    const int logn = v.log;
    const size_t n = (size_t)1 << logn;
    std::vector<QM> vals(n);
    for (size_t i = 0; i < n; i++) vals[i] = v.at(bit_reverse((uint32_t)i, logn));
    Coset dom = coset;
    for (size_t size = n; size > 1; size /= 2) {
        std::vector<uint32_t> xinv(size / 2);
        for (size_t i = 0; i < size / 2; i++) xinv[i] = m_inv(dom.at(i).x);
        for (size_t start = 0; start < n; start += size)
            for (size_t i = 0; i < size / 2; i++) {
                QM l = vals[start + i], r = vals[start + size / 2 + i];
                vals[start + i] = l + r;
                vals[start + size / 2 + i] = (l - r) * xinv[i];
            }
        dom = Coset{(dom.initial * 2) & IDX_MASK, dom.log_size - 1};
    }
    uint32_t inv_n = m_inv((uint32_t)n);
    std::vector<QM> out(n);
    for (size_t i = 0; i < n; i++) out[i] = vals[bit_reverse((uint32_t)i, logn)] * inv_n;
    return out;
}

static inline std::vector<size_t> fold_queries(const std::vector<size_t>& q, int n_folds) {
    std::vector<size_t> out;
    for (size_t x : q) {
        size_t f = x >> n_folds;
        if (out.empty() || out.back() != f) out.push_back(f);
    }
    return out;
}
static inline std::vector<size_t> generate_queries(Channel& ch, int log_domain, uint64_t n_queries) {
    std::vector<size_t> qs;
    uint64_t cnt = 0;
    size_t mask = ((size_t)1 << log_domain) - 1;
    for (;;) {
        Hash rb = ch.draw_random_bytes();
        for (int k = 0; k < 8; k++) {
            uint32_t w;
            memcpy(&w, rb.data() + 4 * k, 4);
            qs.push_back(w & mask);
            if (++cnt == n_queries) {
                std::sort(qs.begin(), qs.end());
                qs.erase(std::unique(qs.begin(), qs.end()), qs.end());
                return qs;
            }
        }
    }
}

struct FriLayerProof {
    std::vector<QM> fri_witness;
    Decommitment decommitment;
    Hash commitment;
};
struct FriProof {
    FriLayerProof first;
    std::vector<FriLayerProof> inner;
    std::vector<QM> last_layer_poly;
    uint32_t last_log;
};

static inline std::vector<ColRef> coord_refs(const QCol& q) {
    return {{q.c[0].data(), q.log}, {q.c[1].data(), q.log}, {q.c[2].data(), q.log}, {q.c[3].data(), q.log}};
}
// compute_decommitment_positions_and_witness_evals (fold step 1)
static inline void decommit_positions(const QCol& col, const std::vector<size_t>& qp, std::vector<size_t>& positions,
                                      std::vector<QM>& witness) {
    size_t i = 0;
    while (i < qp.size()) {
        size_t j = i;
        while (j < qp.size() && (qp[j] >> 1) == (qp[i] >> 1)) j++;
        size_t start = (qp[i] >> 1) << 1, k = i;
        for (size_t pos = start; pos < start + 2; pos++) {
            positions.push_back(pos);
            if (k < j && qp[k] == pos) {
                k++;
                continue;
            }
            witness.push_back(col.at(pos));
        }
        i = j;
    }
}

struct FriProver {
    std::vector<QCol> columns;  // strictly decreasing log sizes
    MerkleTree first_tree;
    struct Inner {
        QCol eval;
        MerkleTree tree;
    };
    std::vector<Inner> inner;
    std::vector<QM> last_layer_poly;

    void commit(Channel& ch, uint32_t log_blowup, uint32_t log_last, std::vector<QCol>&& cols) {
        columns = std::move(cols);
        std::vector<ColRef> coord;
        for (auto& q : columns)
            for (auto& r : coord_refs(q)) coord.push_back(r);
        first_tree = merkle_commit(coord);
        ch.mix_root(first_tree.root());
        QM alpha = ch.draw_secure_felt();
        int llog = columns[0].log - 1;
        Coset line = Coset::half_odds(llog);
        QCol layer;
        layer.alloc(llog);
        size_t ci = 0;
        const size_t last_size = (size_t)1 << (log_last + log_blowup);
        while (((size_t)1 << layer.log) > last_size) {
            while (ci < columns.size() && columns[ci].log - 1 == layer.log) {
                fold_circle_into_line(layer, columns[ci], alpha);
                ci++;
            }
            Inner in;
            in.eval = std::move(layer);
            in.tree = merkle_commit(coord_refs(in.eval));
            ch.mix_root(in.tree.root());
            alpha = ch.draw_secure_felt();
            layer = fold_line(in.eval, alpha);
            inner.push_back(std::move(in));
            // merkle trees hold pointers into in.eval's buffers: moving a vector keeps its heap buffer
            line = Coset{(line.initial * 2) & IDX_MASK, line.log_size - 1};
        }
        if (ci != columns.size()) throw std::runtime_error("fri: columns left unfolded");
        auto coeffs = line_interpolate(layer, line);
        size_t bound = (size_t)1 << log_last;
        for (size_t k = bound; k < coeffs.size(); k++)
            if (!coeffs[k].is_zero()) throw std::runtime_error("fri: invalid degree");
        coeffs.resize(bound);
        last_layer_poly = coeffs;
        ch.mix_felts(last_layer_poly.data(), last_layer_poly.size());
    }

    FriProof decommit(const std::vector<size_t>& queries) {
        FriProof p;
        int max_log = columns[0].log;
        std::map<int, std::vector<size_t>> pos_by_size;
        for (auto& col : columns) {
            auto cq = fold_queries(queries, max_log - col.log);
            decommit_positions(col, cq, pos_by_size[col.log], p.first.fri_witness);
        }
        std::vector<uint32_t> unused;
        merkle_decommit(first_tree, pos_by_size, unused, p.first.decommitment);
        p.first.commitment = first_tree.root();
        auto lq = fold_queries(queries, 1);
        for (auto& in : inner) {
            FriLayerProof lp;
            std::vector<size_t> positions;
            decommit_positions(in.eval, lq, positions, lp.fri_witness);
            std::map<int, std::vector<size_t>> m;
            m[in.eval.log] = positions;
            std::vector<uint32_t> u2;
            merkle_decommit(in.tree, m, u2, lp.decommitment);
            lp.commitment = in.tree.root();
            p.inner.push_back(std::move(lp));
            lq = fold_queries(lq, 1);
        }
        p.last_layer_poly = last_layer_poly;
        uint32_t l = 0;
        while (((size_t)1 << l) < last_layer_poly.size()) l++;
        p.last_log = l;
        return p;
    }
};

}  // namespace cpu
