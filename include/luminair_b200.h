/*
 * luminair_b200 — C ABI of the B200-native Circle-STARK prover backend.
 *
 * This is the drop-in boundary for LuminAIR's hot path: every entry point replaces one of the
 * stwo `SimdBackend` trait methods that /root/reference reaches from
 * crates/prover/src/prover.rs:28-319 (`prove`) and from the witness writers in crates/air
 * (through `TreeBuilder<SimdBackend>`, crates/air/src/utils.rs:112-128).  A Rust `CudaBackend`
 * shim binds these with `extern "C"`.
 *
 * Conventions
 *   - all functions return 0 on success, a negative LB_ERR_* otherwise; lb_last_error() gives text
 *   - no exceptions, no callbacks, no globals: one lb_ctx per device/stream, re-entrant per ctx
 *   - device buffers are plain device pointers to u32 (M31 values, canonical [0, 2^31-1))
 *   - a "column batch" is n_cols columns of 2^log_size u32 at a fixed element stride
 *   - evaluations are in stwo's bit-reversed circle-domain order; QM31 columns are 4 coordinate
 *     columns (SecureColumnByCoords)
 *   - host pointers are caller-owned and only touched during the call
 */
#ifndef LUMINAIR_B200_H
#define LUMINAIR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LB_OK 0
#define LB_ERR_CUDA (-1)
#define LB_ERR_OOM (-2)
#define LB_ERR_BAD_ARG (-3)
#define LB_ERR_NCCL (-4)
#define LB_ERR_CONSTRAINTS (-5) /* ProvingError::ConstraintsNotSatisfied (prover.rs:312) */

typedef struct lb_ctx lb_ctx;

/* ---- context / memory ---------------------------------------------------------------- */
int lb_version(void);
/* A context belongs to one GPU and owns a stream, the twiddle tables, a memory pool for its stream-ordered allocations and a
 * pinned staging buffer.  Calls on one context are serial (one stream); several contexts - on different GPUs or on the same
 * one - can be driven from different host threads at the same time and do not wait on each other.  Every entry point makes
 * the context's device current for the call and restores the caller's. */
int lb_ctx_create(int device, lb_ctx** out);
void lb_ctx_destroy(lb_ctx* ctx);
const char* lb_last_error(lb_ctx* ctx);
int lb_sync(lb_ctx* ctx);
/* Column<M31>::zeros / uninitialized / to_cpu  (stwo ColumnOps) */
int lb_alloc(lb_ctx* ctx, size_t n_u32, uint32_t** d_out);
int lb_free(lb_ctx* ctx, uint32_t* d_ptr);
/* stream-ordered variants (cudaMallocAsync / cudaFreeAsync on the context's stream; memory stays in the device pool): no
 * device-wide synchronisation.  For short-lived buffers; not valid for lb_ipc_export. */
int lb_alloc_pooled(lb_ctx* ctx, size_t n_u32, uint32_t** d_out);
int lb_free_pooled(lb_ctx* ctx, uint32_t* d_ptr);
int lb_memset_zero(lb_ctx* ctx, uint32_t* d_ptr, size_t n_u32);
int lb_upload(lb_ctx* ctx, uint32_t* d_dst, const uint32_t* h_src, size_t n_u32);
int lb_download(lb_ctx* ctx, uint32_t* h_dst, const uint32_t* d_src, size_t n_u32);
int lb_copy(lb_ctx* ctx, uint32_t* d_dst, const uint32_t* d_src, size_t n_u32);
/* CUDA-event timing on the context's stream (bench / profiling) */
int lb_timer_start(lb_ctx* ctx);
int lb_timer_stop_ms(lb_ctx* ctx, float* ms_out);
int lb_device_info(lb_ctx* ctx, int* sm_count, size_t* total_mem_bytes);

/* ---- PolyOps ----------------------------------------------------------------------------- */
/* PolyOps::precompute_twiddles (prover.rs:38-42): tables for every domain of log size <= max_log */
int lb_twiddles_ensure(lb_ctx* ctx, int max_log);
/* stwo-format twiddle tree of the root half coset half_odds(root_log) (2^root_log u32), for parity tests */
int lb_twiddles_export(lb_ctx* ctx, int root_log, uint32_t* d_out);
/* PolyOps::interpolate_columns: evaluations -> coefficients, in place (extend_evals, add/witness.rs:51) */
int lb_interpolate_batch(lb_ctx* ctx, uint32_t* d_cols, size_t stride, int n_cols, int log_size);
/* PolyOps::evaluate_polynomials: coefficients (2^log_in) -> evaluations on CanonicCoset(log_out) */
int lb_evaluate_batch(lb_ctx* ctx, const uint32_t* d_coeffs, size_t src_stride, int log_in, uint32_t* d_out,
                      size_t dst_stride, int log_out, int n_cols);

/* evaluate_polynomials fused with the column->row all-to-all of a column-sharded commit: the final pass stores
 * rows [s*R/W, (s+1)*R/W) (R = 2^log_out, W = n_peers) of column c into h_peers[s] + (peer_col0 + c) * (R/W).
 * h_peers: HOST array of W DEVICE pointers (this GPU's own buffer or lb_ipc_open'ed NVLink peer buffers).
 * d_scratch (n_cols x 2^log_out at dst_stride) holds the intermediate passes.  Needs log_out >= 16 and R/W >= 4096. */
int lb_evaluate_batch_scatter(lb_ctx* ctx, const uint32_t* d_coeffs, size_t src_stride, int log_in, uint32_t* d_scratch,
                              size_t dst_stride, int log_out, int n_cols, uint32_t* const* h_peers, int n_peers,
                              size_t peer_col0);
/* CUDA IPC plumbing for the above: export an lb_alloc'ed buffer, open a peer's handle, close it */
int lb_ipc_export(lb_ctx* ctx, const uint32_t* d_ptr, uint8_t handle_out[64]);
int lb_ipc_open(lb_ctx* ctx, const uint8_t handle[64], uint32_t** d_ptr_out);
int lb_ipc_close(lb_ctx* ctx, uint32_t* d_ptr);

/* Host-buffer form of extend_evals + commit's evaluate (values -> coefficients -> evaluations on CanonicCoset(log_out)):
 * h_values: n_cols x 2^log_in, h_evals: n_cols x 2^log_out, h_coeffs (optional, may be NULL): n_cols x 2^log_in, all HOST.
 * Columns are processed in chunks (chunk_cols, 0 = automatic) over three streams, so the upload of chunk k+1, the two
 * transforms of chunk k and the download of chunk k-1 overlap; the host buffers should be page-locked for that. */
int lb_lde_host(lb_ctx* ctx, const uint32_t* h_values, uint32_t* h_evals, int n_cols, int log_in, int log_out,
                uint32_t* h_coeffs, int chunk_cols);

/* ---- ColumnOps / FieldOps / AccumulationOps: the small trait methods stwo::prover::prove and the LogUp generator reach ---- */
/* ColumnOps<BaseField>::bit_reverse_column, in place (a SecureField column is its 4 coordinate columns: 4 calls) */
int lb_bit_reverse(lb_ctx* ctx, uint32_t* d_col, int log_size);
/* PolyOps::new_canonical_ordered: 2^log_size values in canonic-coset order -> bit-reversed circle-domain order (d_out != d_in) */
int lb_new_canonical_ordered(lb_ctx* ctx, const uint32_t* d_in, uint32_t* d_out, int log_size);
/* FieldOps<BaseField>::batch_inverse / FieldOps<SecureField>::batch_inverse (4 coordinate columns); d_out may equal d_in.
 * A zero input fails with LB_ERR_BAD_ARG "0 has no inverse" (stwo panics with that text). */
int lb_batch_inverse_m31(lb_ctx* ctx, const uint32_t* d_in, uint32_t* d_out, size_t n);
int lb_batch_inverse_qm31(lb_ctx* ctx, const uint32_t* const d_in[4], uint32_t* const d_out[4], size_t n);
/* AccumulationOps::accumulate: column += other, coordinate-wise over SecureColumnByCoords of n elements */
int lb_accumulate(lb_ctx* ctx, uint32_t* const d_column[4], const uint32_t* const d_other[4], size_t n);
/* AccumulationOps::generate_secure_powers: h_out[4 * k .. 4 * k + 3] = felt^k for k < n_powers (host; the sequence is serial) */
int lb_generate_secure_powers(const uint32_t felt[4], int n_powers, uint32_t* h_out);

/* ---- MerkleOps<Blake2sMerkleHasher> ------------------------------------------------------- */
/* commit_on_layer: d_out = 2^log_size digests (8 u32 each); d_prev = child layer or NULL;
 * h_cols = HOST array of n_cols DEVICE column pointers (columns of exactly this log size) */
int lb_merkle_commit_layer(lb_ctx* ctx, int log_size, const uint32_t* d_prev, const uint32_t* const* h_cols,
                           int n_cols, uint32_t* d_out);
/* h_out[k*n_cols + c] = cols[c][h_idx[k]]  (decommitment gathers) */
int lb_gather_rows(lb_ctx* ctx, const uint32_t* const* h_cols, int n_cols, const uint32_t* h_idx, int n_idx,
                   uint32_t* h_out);

/* ---- PolyOps::eval_at_point (OODS sampling inside stwo::prover::prove, prover.rs:311-312) ---- */
/* h_cols: HOST array of n_cols DEVICE coefficient columns (2^log_size each); point = x (4 u32) then y (4 u32),
 * QM31 coordinates; h_out: HOST, 4 u32 per column */
int lb_eval_at_point(lb_ctx* ctx, const uint32_t* const* h_cols, int n_cols, int log_size, const uint32_t point[8],
                     uint32_t* h_out);

/* ---- QuotientOps::accumulate_quotients (DEEP quotients) --------------------------------------- */
/* one ColumnSampleBatch: a sample point and the (column index, sampled value) pairs taken there */
typedef struct {
    uint32_t point[8];      /* x (4 u32), y (4 u32) */
    int n_cols;
    const int* col_idx;     /* indices into h_cols */
    const uint32_t* values; /* 4 u32 per column */
} lb_sample_batch;
/* h_cols: HOST array of DEVICE LDE columns of 2^log_size; d_out: 4 DEVICE coordinate columns of 2^log_size */
int lb_accumulate_quotients(lb_ctx* ctx, int log_size, const uint32_t* const* h_cols, int n_cols,
                            const lb_sample_batch* batches, int n_batches, const uint32_t random_coeff[4],
                            uint32_t* const d_out[4]);

/* Column-sharded form (BASELINE cfg 5): this device holds columns [col_offset, col_offset + batch.n_cols) of a
 * ColumnSampleBatch that has n_cols_global columns across all devices.  d_out receives this device's PARTIAL sum (the
 * line-coefficient terms of its own columns only, weighted with the global powers of random_coeff); the coordinate-wise
 * M31 sum of all devices' outputs equals lb_accumulate_quotients over all columns. */
typedef struct {
    int col_offset, n_cols_global;
} lb_batch_shard;
int lb_accumulate_quotients_shard(lb_ctx* ctx, int log_size, const uint32_t* const* h_cols, int n_cols,
                                  const lb_sample_batch* batches, const lb_batch_shard* shards, int n_batches,
                                  const uint32_t random_coeff[4], uint32_t* const d_out[4]);

/* ---- FriOps ------------------------------------------------------------------------------------- */
/* dst (4 coords of 2^(log_size-1)) = dst * alpha^2 + fold(src (4 coords of 2^log_size on CanonicCoset(log_size))) */
int lb_fold_circle_into_line(lb_ctx* ctx, uint32_t* const d_dst[4], const uint32_t* const d_src[4], int log_size,
                             const uint32_t alpha[4]);
/* dst (4 coords of 2^(log_size-1)) = fold(src on LineDomain(half_odds(log_size))) */
int lb_fold_line(lb_ctx* ctx, uint32_t* const d_dst[4], const uint32_t* const d_src[4], int log_size,
                 const uint32_t alpha[4]);

/* ---- GrindOps<Blake2sChannel>::grind -------------------------------------------------------------- */
/* smallest nonce whose mix_u64 into `digest` leaves >= pow_bits trailing zero bits; channel_variant as in lb_prove_config */
int lb_grind(lb_ctx* ctx, const uint32_t digest[8], int channel_variant, uint32_t pow_bits, uint64_t* nonce_out);

/* ---- AIR kernels (crates/air): component ids ------------------------------------------------------- */
#define LB_COMP_ADD 0    /* components/add    (component.rs:38-116, witness.rs:126-167) */
#define LB_COMP_MUL 1    /* components/mul    */
#define LB_COMP_INPUTS 2 /* components/inputs */
#define LB_COMP_SUM_REDUCE 4   /* components/sum_reduce */
#define LB_COMP_MAX_REDUCE 5   /* components/max_reduce */
#define LB_COMP_CONTIGUOUS 6   /* components/contiguous */
#define LB_COMP_MUL_ARTIFACT 3 /* Mul AIR of the revision that produced ui/demo/public/proof (KAT only) */
#define LB_COMP_RECIP 7        /* components/recip   (component.rs:39-107) */
#define LB_COMP_SQRT 8         /* components/sqrt    */
#define LB_COMP_REM 9          /* components/rem     */
#define LB_COMP_SIN 10         /* components/sin     (component.rs:51-123; third relation use = the sin LUT) */
#define LB_COMP_EXP2 11        /* components/exp2    */
#define LB_COMP_LOG2 12        /* components/log2    */
#define LB_COMP_SIN_LOOKUP 13  /* components/lookups/sin  (component.rs:41-59: -multiplicity per table row) */
#define LB_COMP_EXP2_LOOKUP 14 /* components/lookups/exp2 */
#define LB_COMP_LOG2_LOOKUP 15 /* components/lookups/log2 */
#define LB_COMP_LESS_THAN 16   /* components/less_than (component.rs:49-184: 4 limbs range-checked) */
#define LB_COMP_RANGE_CHECK_LOOKUP 17 /* components/lookups/range_check */
/* LuminairInteractionElements (components/mod.rs:220-236, lookups/mod.rs:31-51), in draw order */
#define LB_REL_NODE 0
#define LB_REL_SIN 1
#define LB_REL_EXP2 2
#define LB_REL_LOG2 3
#define LB_REL_RANGE_CHECK 4
#define LB_REL_COUNT 5
typedef struct {
    uint32_t z[4], alpha[4]; /* relation!(X, N): combine(v) = sum_i alpha^i v_i - z */
} lb_relation;
/* InteractionClaimGenerator::write_interaction_trace: LogUp columns from the main trace (values, not
 * coefficients).  d_inter receives 4*k columns of 2^log_size; claimed_out = claimed sum (4 u32, HOST). */
int lb_logup_interaction_trace(lb_ctx* ctx, int component, const uint32_t* d_main, size_t main_stride, uint32_t* d_inter,
                               size_t inter_stride, int log_size, const uint32_t z[4], const uint32_t alpha[4],
                               uint32_t claimed_out[4]);
/* the same for every component: rels = all LB_REL_COUNT relations; d_lut = the preprocessed LUT column values
 * (2^log_size each) a lookup-table component tabulates (lookups/exp2/witness.rs:117-144), NULL otherwise */
int lb_logup_interaction_trace_lut(lb_ctx* ctx, int component, const uint32_t* d_main, size_t main_stride,
                                   const uint32_t* const d_lut[2], uint32_t* d_inter, size_t inter_stride, int log_size,
                                   const lb_relation rels[LB_REL_COUNT], uint32_t claimed_out[4]);
/* ComponentProver::evaluate_constraint_quotients_on_domain for FrameworkComponent<XEval>
 * (crates/air/src/components/mod.rs:530-601): d_main / d_inter are the LDE columns on CanonicCoset(log_size+1);
 * pows = this component's random-coefficient powers, first constraint first (4 u32 each);
 * d_acc = 4 coordinate columns of 2^(log_size+1), overwritten (accumulate = 0) or added to (1). */
int lb_constraint_quotients(lb_ctx* ctx, int component, const uint32_t* d_main, size_t main_stride, const uint32_t* d_inter,
                            size_t inter_stride, int log_size, const uint32_t z[4], const uint32_t alpha[4],
                            const uint32_t claimed_sum[4], const uint32_t* pows, int n_pows, uint32_t* const d_acc[4],
                            int accumulate);

/* the same for every component, on the evaluation domain CanonicCoset(eval_log_size) =
 * max_constraint_log_degree_bound (LUT consumers: max(log_size, lut_log_size) + 1, exp2/component.rs:41-43): every
 * column (d_main, d_inter, d_lut) holds 2^eval_log_size values; d_lut = the LUT columns a lookup-table component reads */
int lb_constraint_quotients_lut(lb_ctx* ctx, int component, const uint32_t* d_main, size_t main_stride, const uint32_t* d_inter,
                                size_t inter_stride, const uint32_t* const d_lut[2], int log_size, int eval_log_size,
                                const lb_relation rels[LB_REL_COUNT], const uint32_t claimed_sum[4], const uint32_t* pows,
                                int n_pows, uint32_t* const d_acc[4], int accumulate);

/* ---- luminair_prover::prover::prove (crates/prover/src/prover.rs:28-319) ---------------------------- */
/* one `TraceTable` of the LuminairPie (crates/air/src/pie.rs:143-148): row-major rows of the component's
 * main-trace columns, canonical M31 values */
typedef struct {
    int slot;             /* field index in LuminairClaim (crates/air/src/lib.rs:30-48): add 0, mul 1, recip 2, sin 3, sin_lookup 4,
                             sum_reduce 5, max_reduce 6, sqrt 7, rem 8, exp2 9, exp2_lookup 10, log2 11, log2_lookup 12,
                             less_than 13, range_check_lookup 14, inputs 15, contiguous 16 */
    int n_cols;
    uint64_t n_rows;      /* unpadded */
    const uint32_t* rows; /* n_rows x n_cols, HOST (or DEVICE when rows_on_device != 0) */
    int rows_on_device;
} lb_trace_table;
typedef struct {
    uint32_t pow_bits, log_blowup_factor, log_last_layer_degree_bound; /* PcsConfig::default(): 5, 1, 0 */
    uint64_t n_queries;                                                  /* 3 */
    int channel_variant; /* 0: Blake2sChannel as pinned by the reference's committed proof; 1: "v2" mixing */
    int n_slots;         /* Option fields in LuminairClaim: 17 */
    int air_era;         /* 0: AIRs of the tree at /root/reference; 1: the revision that produced ui/demo/public/proof
                            (8 claim slots, Mul AIR with one extra, identically-zero constraint) - known-answer test only */
    int draw_lookup_elements; /* 1: LuminairInteractionElements::draw also draws the 4 LUT relations (components/mod.rs:227-235) */
} lb_prove_config;
/* one preprocessed (LUT) column of `lookups_to_preprocessed_column(&settings.lookups)`
 * (crates/air/src/preprocessed.rs:181-206), in that order: sin 0/1, exp2 0/1, log2 0/1, range_check 0.  The values
 * (2^log_size canonical M31, storage order) are generated by the caller exactly as the reference does on the host
 * (preprocessed.rs:233-248, 351-383); lb_prove_with_lookups sorts the columns as PreProcessedTrace::new does. */
typedef struct {
    int lut;                /* LB_REL_SIN .. LB_REL_RANGE_CHECK */
    int col_index;          /* 0: input values, 1: function values */
    int log_size;
    const uint32_t* values; /* HOST (or DEVICE when on_device != 0) */
    int on_device;
} lb_preprocessed_column;
/* Tables in pie order.  On success *proof_out is a malloc'd bincode `LuminairProof` (free with lb_free_host).
 * cfg == NULL: the reference's defaults.  Errors: LB_ERR_BAD_ARG ("TraceError::EmptyTrace", unsupported
 * component), LB_ERR_CONSTRAINTS (ProvingError::ConstraintsNotSatisfied). */
int lb_prove(lb_ctx* ctx, const lb_trace_table* tables, int n_tables, const lb_prove_config* cfg, uint8_t** proof_out,
             size_t* proof_len);
/* prove() for graphs with lookup tables: the preprocessed trace (prover.rs:52-59) holds `lut_columns` */
int lb_prove_with_lookups(lb_ctx* ctx, const lb_trace_table* tables, int n_tables, const lb_preprocessed_column* lut_columns,
                          int n_lut_columns, const lb_prove_config* cfg, uint8_t** proof_out, size_t* proof_len);
void lb_free_host(void* p);

/* ---- multi-GPU prove(): one rank (process or thread) per GPU, NCCL over NVLink / NVSwitch ----------------------
 * The reference has ONE tree per phase holding every column (crates/prover/src/prover.rs:56-59,68,179,298) and one
 * stwo::prover::prove (prover.rs:311-312); lb_prove_sharded runs that same protocol over `world` GPUs and returns, on every
 * rank, the SAME bytes lb_prove returns on one GPU:
 *   - column-wise steps (interpolate, LDE, eval_at_point) are split by column: every column has one owner rank;
 *   - the last pass of an owner's LDE stores every tile straight into the row shard of the rank that owns those rows (rank r
 *     holds rows [r R/W, (r+1) R/W) of every column): the shards live in a symmetric heap mapped from every peer over NVLink
 *     (CUDA IPC; ranks must be separate processes - otherwise, or with LB_SHARD_IPC=0, one NCCL send/recv per peer and run);
 *     everything row-wise - Merkle leaves and sub-trees, constraint quotients, DEEP quotients, FRI folds and layer trees -
 *     then runs on the row shard with no further data exchange, only the 32-byte sub-tree roots are all-gathered (the top
 *     log2 W tree levels are hashed from them);
 *   - the composition polynomial goes rows -> columns once (16 B per row), sampled values and the decommitted words are
 *     summed over the ranks (each has exactly one owner); the last, latency-bound FRI layers are all-gathered and replicated;
 *   - the Fiat-Shamir channel runs replicated on every rank.
 * Every rank passes the same tables / LUT columns / configuration (the trace is replicated input: on the device it is
 * generated by lb_trace_* in microseconds).  NCCL is bound at run time (libnccl.so.2); without it these entry points return
 * LB_ERR_NCCL and the single-GPU entry points are unaffected.  Every committed column needs at least 4 rows per rank. */
#define LB_COMM_ID_BYTES 128
typedef struct lb_comm lb_comm;
/* rank 0 creates the id (ncclGetUniqueId); the host hands it to the other ranks by its own means (MPI, a file, a socket) */
int lb_comm_unique_id(uint8_t id_out[LB_COMM_ID_BYTES]);
/* collective: every rank calls it with its own context (one GPU each); world must be a power of two */
int lb_comm_init(lb_ctx* ctx, const uint8_t id[LB_COMM_ID_BYTES], int rank, int world, lb_comm** out);
void lb_comm_destroy(lb_comm* comm);
/* rank / world and the traffic of the last lb_prove_sharded on this rank: bytes_sent = NCCL sends + direct NVLink peer stores
 * (any pointer may be NULL) */
int lb_comm_stats(const lb_comm* comm, int* rank, int* world, uint64_t* bytes_sent, uint64_t* bytes_received,
                  int* n_collectives);
int lb_prove_sharded(lb_ctx* ctx, lb_comm* comm, const lb_trace_table* tables, int n_tables,
                     const lb_preprocessed_column* lut_columns, int n_lut_columns, const lb_prove_config* cfg,
                     uint8_t** proof_out, size_t* proof_len);
/* diagnostics of the last lb_prove: channel digest after every mix (32 B each) and per-stage wall-clock ms */
int lb_prove_transcript(lb_ctx* ctx, uint8_t* out, size_t cap_hashes, size_t* n_hashes);
int lb_prove_stage_ms(lb_ctx* ctx, float* out, int cap, int* n);

/* ---- trace emitters: LuminairOperator::process_trace on the device ------------------------------
 * CopyToStwo (crates/graph/src/op/prim.rs:72-84), LuminairAdd (:919-1013), LuminairMul: element-wise nodes over device
 * tensors of raw Fixed<12> values (int32: LuminAIR's values are M31 elements, |v| < 2^31).  Each call computes the node's
 * output tensor (d_out, may be NULL for inputs) and appends its n rows, in the column order of the component's
 * *TraceTableRow (components/{inputs,add,mul}/table.rs), to a row-major device table at row `row0` - the format lb_prove takes
 * with rows_on_device = 1.  Consumers take their operands with multiplicity -1 (as the reference does); `out_mult` is what the
 * node yields: its number of consumers, 0 for a final output.  Fixed-point product: q = floor(x*y / 2^12), rem = x*y - q*2^12. */
int lb_trace_inputs(lb_ctx* ctx, uint32_t node_id, const int32_t* d_vals, uint64_t n, uint32_t out_mult, uint32_t* d_rows,
                    uint64_t row0);
int lb_trace_add(lb_ctx* ctx, uint32_t node_id, uint32_t lhs_id, uint32_t rhs_id, const int32_t* d_lhs, const int32_t* d_rhs,
                 uint64_t n, uint32_t out_mult, int32_t* d_out, uint32_t* d_rows, uint64_t row0);
int lb_trace_mul(lb_ctx* ctx, uint32_t node_id, uint32_t lhs_id, uint32_t rhs_id, const int32_t* d_lhs, const int32_t* d_rhs,
                 uint64_t n, uint32_t out_mult, int32_t* d_out, uint32_t* d_rows, uint64_t row0);

/* ---- every operator of the graph: LuminairOperator::process_trace (crates/graph/src/op/prim.rs) -------------------------
 * lb_trace_op is the general form of the three entry points above: operands are device tensors of raw Fixed<12> values plus an
 * optional gather index (u32, one entry per row: the element the row reads - the operator's index expression for
 * broadcasts / expands, flattened by the caller; NULL = identity).  The operator kind is named by the claim slot of its
 * component.  Rows are appended at `row0` of a row-major device table in the component's *TraceTableRow column order
 * (crates/air/src/components/<name>/table.rs); the node's output tensor is written to d_out.
 *   LB_OP_INPUTS      CopyToStwo / LuminairConstant  prim.rs:52-84,151-188 (d_lhs = the tensor; d_out may be NULL)
 *   LB_OP_CONTIGUOUS  prim.rs:229-298     LB_OP_RECIP  :388-428     LB_OP_SQRT :617-657     LB_OP_ADD :967-1013
 *   LB_OP_MUL :1090-1136   LB_OP_REM :1372-1418   LB_OP_LESS_THAN :1225-1292 (counts its four 8-bit limbs in
 *   lookup->d_multiplicities, 256 counters: the RangeCheckLookup<1> table)
 *   LB_OP_SIN :496-543, LB_OP_EXP2 :725-772, LB_OP_LOG2 :840-887: f(x) is read from the LUT column the caller generated on the
 *   host exactly as the reference does (preprocessed.rs:351-383) at LookupLayout::find_index(x) (preprocessed.rs:96-116), and
 *   that entry's counter in lookup->d_multiplicities (the *_lookup table) is incremented; a value outside every range fails
 *   with LB_ERR_BAD_ARG "Value should fit in range." (the reference panics with that text).
 *   LB_OP_SUM_REDUCE :1517-1562, LB_OP_MAX_REDUCE :1685-1731: output element i folds the `group` gathered inputs
 *   [i*group, (i+1)*group); one row per step, n * group rows.
 * d_out_mult: per output element, the multiplicity the node yields it with = the number of rows that consume it
 * (node_info.num_consumers, prim.rs:947-951; 0 for a graph output; NULL = all 0) - count it with lb_trace_count_uses. */
#define LB_OP_ADD 0
#define LB_OP_MUL 1
#define LB_OP_RECIP 2
#define LB_OP_SIN 3
#define LB_OP_SUM_REDUCE 5
#define LB_OP_MAX_REDUCE 6
#define LB_OP_SQRT 7
#define LB_OP_REM 8
#define LB_OP_EXP2 9
#define LB_OP_LOG2 11
#define LB_OP_LESS_THAN 13
#define LB_OP_INPUTS 15
#define LB_OP_CONTIGUOUS 16
#define LB_MAX_LOOKUP_RANGES 8
typedef struct {
    int n_ranges;                       /* LookupLayout.ranges (preprocessed.rs:41-60): sorted, disjoint, inclusive */
    int32_t lo[8], hi[8];               /* LB_MAX_LOOKUP_RANGES entries, raw Fixed<12> values */
    const uint32_t* d_values;           /* DEVICE: LUT column 1 (function values, canonical M31); unused by less_than */
    uint32_t* d_multiplicities;         /* DEVICE: one counter per table entry, incremented */
} lb_lookup;
typedef struct {
    int op;                             /* LB_OP_* */
    uint32_t node_id, lhs_id, rhs_id;   /* rhs_id unused by unary operators */
    const int32_t* d_lhs;
    const uint32_t* d_lhs_idx;
    const int32_t* d_rhs;
    const uint32_t* d_rhs_idx;
    uint64_t n;                         /* output elements */
    uint32_t group;                     /* reductions: steps per output element; else ignored */
    const uint32_t* d_out_mult;
    int32_t* d_out;
    uint32_t* d_rows;
    uint64_t row0;
    const lb_lookup* lookup;            /* sin / exp2 / log2 / less_than */
} lb_trace_op_desc;
int lb_trace_op(lb_ctx* ctx, const lb_trace_op_desc* op);
/* d_uses[d_idx[i]] += 1 for i < n_reads (d_idx NULL: d_uses[i] += 1): one call per operand of every consumer gives the
 * per-element multiplicities of a node (graph.rs:161-604 keeps the same count in node_info) */
int lb_trace_count_uses(lb_ctx* ctx, uint32_t* d_uses, const uint32_t* d_idx, uint64_t n_reads);

#ifdef __cplusplus
}
#endif
#endif /* LUMINAIR_B200_H */
