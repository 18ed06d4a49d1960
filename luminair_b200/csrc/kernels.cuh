// Launchers of the streaming kernels behind the PCS / FRI / AIR stages (pcs_kernels.cu,
// air_kernels.cu).  Everything takes device pointers and a stream; nothing synchronises.
#pragma once
#include "../../include/luminair_b200.h"
#include "air.cuh"
#include "cfft.cuh"

namespace lb {

// eval_at_point works on chunks of 2^EVAL_CHUNK_LOG coefficients (one CTA each, basis entries in registers).  Measured on B200
// (OODS stage of the cfg-3 / wide proof): 12 -> 0.271 / 0.398 ms at 116 registers; 11 -> 0.294 / 0.436 ms at 63 registers
// (more CTAs per SM, but twice the per-chunk reductions and basis loads)
#ifndef LB_EVAL_M
#define LB_EVAL_M 12
#endif
constexpr int EVAL_CHUNK_LOG = LB_EVAL_M;

cudaError_t kernels_init(cudaStream_t stream);  // constant tables (call once per context)

// ---- PolyOps::eval_at_point ------------------------------------------------------------------
// d_cols: device table of n_cols coefficient columns (2^log each).  d_mappings: log QM31 fold
// factors [y, x, pi(x), ...] (host computes them).  d_basis: scratch of 2^min(log,12) QM31.
// d_partials: scratch of n_cols * 2^(log - min(log,12)) QM31.  d_out: n_cols QM31.
cudaError_t eval_at_point(const uint32_t* const* d_cols, int n_cols, int log, const QM31* d_mappings, QM31* d_basis,
                          QM31* d_partials, QM31* d_out, cudaStream_t stream);

// ---- QuotientOps::accumulate_quotients ---------------------------------------------------------
constexpr int MAX_QUOTIENT_BATCHES = 4;  // distinct sample points per column size (LuminAIR: the OODS point and its predecessor)
struct QuotientBatch {
    CM31 prx, pry, pix, piy;  // sample point: real / imaginary CM31 halves of x and y
    QM31 sum_a, sum_b;        // sum_j alpha^j a_j, sum_j alpha^j b_j  (line coefficients, folded on host)
    QM31 rc_pow;              // random_coeff ^ |batch|
    int first, count;         // entries [first, first+count) of the entry table
};
struct QuotientEntry {
    QM31 c;   // alpha^j * c_j
    int col;  // index into d_cols
    int pad[3];
};
struct QuotientParams {
    int n_batches;
    QuotientBatch b[MAX_QUOTIENT_BATCHES];
};
// row0 / n_rows: the row range [row0, row0 + n_rows) of the domain the columns and outputs hold (n_rows = 0: all 2^log rows)
cudaError_t accumulate_quotients(uint32_t* const out[4], const uint32_t* const* d_cols, const QuotientEntry* d_entries,
                                 const QuotientParams& qp, const Twiddles* tw, int log, cudaStream_t stream, uint32_t row0 = 0,
                                 uint32_t n_rows = 0);

// ---- FriOps -------------------------------------------------------------------------------------
// dst (4 coords, n/2) = dst * alpha^2 + fold(src (4 coords, n = 2^log));  itw = inverse y twiddles of the domain
cudaError_t fold_circle_into_line(uint32_t* const dst[4], const uint32_t* const src[4], const uint2* itw, int log,
                                  QM31 alpha, cudaStream_t stream);
// dst (4 coords, n/2) = fold(src (4 coords, n = 2^log)); itw = inverse x twiddles of the line domain
cudaError_t fold_line(uint32_t* const dst[4], const uint32_t* const src[4], const uint2* itw, int log, QM31 alpha,
                      cudaStream_t stream);

// the same with the folding coefficient in device memory
cudaError_t fold_circle_into_line_dev(uint32_t* const dst[4], const uint32_t* const src[4], const uint2* itw, int log,
                                      const QM31* d_alpha, cudaStream_t stream);
cudaError_t fold_line_dev(uint32_t* const dst[4], const uint32_t* const src[4], const uint2* itw, int log, const QM31* d_alpha,
                          cudaStream_t stream);
// device-side Blake2sChannel state for the FRI commit loop: digest <- mix_root(*d_root); *d_alpha_out = draw_secure_felt();
// the new digest is also written to d_digest_log (8 words)
struct DevChannel {
    uint32_t digest[8];
    uint32_t n_sent;
    uint32_t pad[7];
};
cudaError_t channel_mix_root_draw(DevChannel* d_ch, const uint32_t* d_root, int variant, QM31* d_alpha_out, uint32_t* d_digest_log,
                                  cudaStream_t stream);
// the same for a row-sharded layer: d_roots = the all-gathered sub-tree roots (2^logw x 8 words); hashes the top logw levels
// into d_top_out (level k at word 8 * (2^k - 1); 8 * (2^(logw+1) - 1) words), mixes the root, draws
cudaError_t channel_mix_sharded_root_draw(DevChannel* d_ch, const uint32_t* d_roots, int logw, int variant, QM31* d_alpha_out,
                                          uint32_t* d_digest_log, uint32_t* d_top_out, cudaStream_t stream);
// the last FRI layers (line layers of at most 2^FRI_TAIL_MAX_LOG values) in one launch: per layer lg = from_log .. last_log + 1:
// Merkle tree of the layer -> mix_root -> draw -> fold_line into layer lg - 1
constexpr int FRI_TAIL_MAX_LOG = 10;
struct FriTailArgs {
    uint32_t* vals[FRI_TAIL_MAX_LOG + 1];    // vals[lg]: the 4 coordinate columns of layer lg, 2^lg words apart
    uint32_t* tree[FRI_TAIL_MAX_LOG + 1];    // tree[lg]: all levels of the tree over layer lg, level k at word 8 * (2^k - 1)
    const uint2* itw[FRI_TAIL_MAX_LOG + 1];  // inverse twiddles of fold_line at layer lg
    int from_log, last_log, variant;
    DevChannel* ch;
    QM31* alphas;       // alphas[0]: coefficient that produced layer from_log; alphas[s + 1] is drawn after layer from_log - s
    uint32_t* digests;  // 8 words per layer: channel digest after each mix_root
};
cudaError_t fri_tail(const FriTailArgs& a, cudaStream_t stream);

// ---- GrindOps ------------------------------------------------------------------------------------
// tests nonces [base, base + count); *d_found = min nonce that works (or UINT64_MAX)
cudaError_t grind_range(const uint32_t digest[8], int variant, uint32_t pow_bits, uint64_t base, uint64_t count,
                        unsigned long long* d_found, cudaStream_t stream);

// ---- small column utilities -----------------------------------------------------------------------
cudaError_t add_inplace(uint32_t* dst, const uint32_t* src, size_t n, cudaStream_t stream);  // dst += src (M31)
cudaError_t bit_reverse(uint32_t* col, int log, cudaStream_t stream);  // in place
cudaError_t canonical_to_storage(uint32_t* out, const uint32_t* in, int log, cudaStream_t stream);
// out = 1 / in element-wise; *d_flag is set to 1 when an input is zero (its output is then unspecified)
cudaError_t batch_inverse_m31(uint32_t* out, const uint32_t* in, size_t n, int* d_flag, cudaStream_t stream);
cudaError_t batch_inverse_qm31(uint32_t* const out[4], const uint32_t* const in[4], size_t n, int* d_flag, cudaStream_t stream);
cudaError_t gather_words(uint32_t* d_out, const uint32_t* const* d_addrs, int n, cudaStream_t stream);
// rows (row-major n_rows x n_cols) -> n_cols columns of 2^log at `stride`, padded with the component's
// `padding()` row (write_trace, e.g. add/witness.rs:43-46; air.cuh padding_value)
cudaError_t transpose_pad(uint32_t* d_cols, size_t stride, const uint32_t* d_rows, uint64_t n_rows, int n_cols, int log,
                          int kind, cudaStream_t stream);

// ---- AIR: LogUp interaction trace + constraint quotients -----------------------------------------
// main: component's n_main trace columns (2^log at main_stride) ; pre: the preprocessed LUT columns (values on
// CanonicCoset(log)) a lookup-table component tabulates, else unused ; inter: 4*n_fracs columns out.
// d_scan_tmp: 4 * 2^log u32 ; d_block_sums: 4 * (2^log / 1024 + 1) u32 ; d_claimed: 4 u32 out.
struct PreCols {
    const uint32_t* p[2];
};
cudaError_t logup_interaction_trace(int kind, const uint32_t* main, size_t main_stride, PreCols pre, uint32_t* inter,
                                    size_t inter_stride, int log, const Relations& rels, uint32_t* d_scan_tmp,
                                    uint32_t* d_block_sums, uint32_t* d_claimed, cudaStream_t stream);

struct ConstraintParams {
    const uint32_t* main;
    size_t main_stride;  // columns evaluated on CanonicCoset(eval_log)
    const uint32_t* inter;
    size_t inter_stride;
    PreCols pre;         // preprocessed columns on CanonicCoset(eval_log) (lookup-table components)
    uint32_t* acc[4];
    int accumulate;  // 0: store, 1: add to acc
    int log_size, eval_log;
    Relations rels;
    QM31 cumsum_shift;
    QM31 pows[MAX_CONSTRAINTS];  // this component's random-coefficient powers, first constraint first
    const uint32_t* denom_inv;   // DEVICE: 1 / Z_H at eval_domain.at(bitrev(i)), i < 2^(eval_log - log_size)
    // row range [row0, row0 + n_rows) of the evaluation domain held by main / inter / pre / acc (n_rows = 0: the whole
    // domain).  inter_prev: for a row shard, the [-1]-offset mask values of the last LogUp column as 4 coordinate columns
    // at inter_stride (the predecessor rows live on other ranks; the column's owner ships them, see shifted_prev_column);
    // NULL: read them from `inter` at the predecessor row.
    uint32_t row0, n_rows;
    const uint32_t* inter_prev;
};
cudaError_t constraint_quotients(int kind, const ConstraintParams& p, cudaStream_t stream);
// out[j] = col[offset_bit_reversed_circle_domain_index(j, domain_log, eval_log, -1)], j < 2^eval_log; the source column is held
// as n_src equal row ranges (src_shards[s] = rows [s R/n_src, ...)), the result is written as n_out equal row ranges; both
// counts are powers of two <= 8 (1 = one plain column)
cudaError_t shifted_prev_column(uint32_t* const out_shards[8], int n_out, const uint32_t* const src_shards[8], int n_src,
                                int domain_log, int eval_log, cudaStream_t stream);



// ---- every operator's process_trace (trace_kernels.cu; include/luminair_b200.h lb_trace_op) ---------------------------
struct LookupRanges {
    int n;
    int32_t lo[LB_MAX_LOOKUP_RANGES], hi[LB_MAX_LOOKUP_RANGES];
    uint32_t base[LB_MAX_LOOKUP_RANGES];  // table index of lo[k]
};
struct TraceOp {
    int op;
    uint32_t node_id, lhs_id, rhs_id;
    const int32_t* lhs;
    const uint32_t* lhs_idx;
    const int32_t* rhs;
    const uint32_t* rhs_idx;
    uint64_t n;
    uint32_t group;
    const uint32_t* out_mult;  // per output element, or NULL: out_mult_all for every element
    uint32_t out_mult_all;
    int32_t* out;
    uint32_t* rows;  // already offset to the node's first row
    LookupRanges lut;
    const uint32_t* lut_vals;
    uint32_t* lut_mult;
    int* err;  // device flag: set when a lookup input is outside every range
};
cudaError_t trace_op(const TraceOp& p, cudaStream_t stream);
cudaError_t trace_count_uses(uint32_t* uses, const uint32_t* idx, uint64_t n, cudaStream_t stream);

}  // namespace lb
