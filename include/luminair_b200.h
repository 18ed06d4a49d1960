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

typedef struct lb_ctx lb_ctx;

/* ---- context / memory ---------------------------------------------------------------- */
int lb_version(void);
/* one context per GPU; owns a stream, scratch and the twiddle tables */
int lb_ctx_create(int device, lb_ctx** out);
void lb_ctx_destroy(lb_ctx* ctx);
const char* lb_last_error(lb_ctx* ctx);
int lb_sync(lb_ctx* ctx);
/* Column<M31>::zeros / uninitialized / to_cpu  (stwo ColumnOps) */
int lb_alloc(lb_ctx* ctx, size_t n_u32, uint32_t** d_out);
int lb_free(lb_ctx* ctx, uint32_t* d_ptr);
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

/* ---- MerkleOps<Blake2sMerkleHasher> ------------------------------------------------------- */
/* commit_on_layer: d_out = 2^log_size digests (8 u32 each); d_prev = child layer or NULL;
 * h_cols = HOST array of n_cols DEVICE column pointers (columns of exactly this log size) */
int lb_merkle_commit_layer(lb_ctx* ctx, int log_size, const uint32_t* d_prev, const uint32_t* const* h_cols,
                           int n_cols, uint32_t* d_out);
/* h_out[k*n_cols + c] = cols[c][h_idx[k]]  (decommitment gathers) */
int lb_gather_rows(lb_ctx* ctx, const uint32_t* const* h_cols, int n_cols, const uint32_t* h_idx, int n_idx,
                   uint32_t* h_out);

#ifdef __cplusplus
}
#endif
#endif /* LUMINAIR_B200_H */
