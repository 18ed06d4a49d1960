// A compiled-language caller of the C ABI (what the reference's Rust host side would do through FFI):
// builds the trace tables of examples/simple (/root/reference/examples/simple/src/main.rs:15-22:
// c = a * b; d = c + w; e = c * d over 2x2 tensors, Fixed<12>) the way LuminairGraph::gen_trace emits them
// (crates/graph/src/op/prim.rs:72-84 CopyToStwo, :919-1013 Add/Mul; node ids of ui/demo/public/graph.dot),
// calls lb_prove, and compares the bincode LuminairProof with a fixture file when one is given.
//
//   g++ -O2 -std=c++17 -Iinclude examples/prove_simple.cpp -Lluminair_b200 -lluminair_b200 -Wl,-rpath,$PWD/luminair_b200 -o examples/prove_simple
//   examples/prove_simple [tests/golden/simple_current.proof.bin] [--device-trace]
// With --device-trace the tables are not built here: the three input tensors are uploaded and every operator's process_trace
// runs on the device (lb_trace_count_uses + lb_trace_op), the way a CudaBackend build of LuminairGraph::gen_trace would.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "luminair_b200.h"

static const uint32_t P = 0x7FFFFFFFu;
static uint32_t m31(int64_t v) { return (uint32_t)(((v % (int64_t)P) + (int64_t)P) % (int64_t)P); }

struct Table {
    int n_cols;
    std::vector<uint32_t> rows;
    void push(std::initializer_list<int64_t> r) {
        for (int64_t v : r) rows.push_back(m31(v));
    }
};

int main(int argc, char** argv) {
    const int64_t S = 4096;  // Fixed<12>
    int64_t a[4] = {1 * S, 2 * S, 3 * S, 4 * S}, b[4] = {10 * S, 20 * S, 30 * S, 40 * S}, w[4] = {-S, -S, -S, -S};
    int64_t c[4], c_rem[4], d[4], e[4], e_rem[4];
    for (int i = 0; i < 4; ++i) {
        int64_t p = a[i] * b[i];
        c[i] = p >> 12;
        c_rem[i] = p - (c[i] << 12);
        d[i] = c[i] + w[i];
        int64_t q = c[i] * d[i];
        e[i] = q >> 12;
        e_rem[i] = q - (e[i] << 12);
    }
    // node ids: mul c = 3 (inputs 6, 7), add d = 4 (inputs 3, 8), mul e = 5 (inputs 3, 4); CopyToStwo nodes 6, 7, 8
    Table add{15, {}}, mul{16, {}}, inp{7, {}};
    for (int i = 0; i < 4; ++i) add.push({4, 3, 8, i, i == 3, 4, 3, 8, i + 1, c[i], w[i], d[i], -1, -1, 1});
    for (int i = 0; i < 4; ++i) mul.push({3, 6, 7, i, i == 3, 3, 6, 7, i + 1, a[i], b[i], c[i], c_rem[i], -1, -1, 2});
    for (int i = 0; i < 4; ++i) mul.push({5, 3, 4, i, i == 3, 5, 3, 4, i + 1, c[i], d[i], e[i], e_rem[i], -1, -1, 0});
    const int64_t* ins[3] = {a, b, w};
    for (int t = 0; t < 3; ++t)
        for (int i = 0; i < 4; ++i) inp.push({6 + t, i, i == 3, 6 + t, i + 1, ins[t][i], 1});

    lb_ctx* ctx = nullptr;
    if (lb_ctx_create(0, &ctx) != LB_OK) {
        std::fprintf(stderr, "lb_ctx_create failed (no CUDA device?)\n");
        return 2;
    }
    lb_trace_table tables[3] = {
        {0, add.n_cols, add.rows.size() / add.n_cols, add.rows.data(), 0},    // LuminairClaim slot 0: add
        {1, mul.n_cols, mul.rows.size() / mul.n_cols, mul.rows.data(), 0},    // slot 1: mul
        {15, inp.n_cols, inp.rows.size() / inp.n_cols, inp.rows.data(), 0},   // slot 15: inputs
    };
    bool device_trace = false;
    const char* fixture = nullptr;
    for (int k = 1; k < argc; ++k) {
        if (!std::strcmp(argv[k], "--device-trace")) device_trace = true;
        else fixture = argv[k];
    }
    if (device_trace) {
        // tensors (raw Fixed<12>, int32) of nodes 3..8 and their per-element consumer counts, all on the device
        uint32_t *val[9] = {}, *uses[9] = {}, *t_add = nullptr, *t_mul = nullptr, *t_inp = nullptr;
        int ok = LB_OK;
        for (int node = 3; node <= 8 && ok == LB_OK; ++node) {
            ok = lb_alloc(ctx, 4, &val[node]);
            if (ok == LB_OK) ok = lb_alloc(ctx, 4, &uses[node]);
            if (ok == LB_OK) ok = lb_memset_zero(ctx, uses[node], 4);
        }
        for (int t = 0; t < 3 && ok == LB_OK; ++t) {
            int32_t raw[4];
            for (int i = 0; i < 4; ++i) raw[i] = (int32_t)ins[t][i];
            ok = lb_upload(ctx, val[6 + t], reinterpret_cast<const uint32_t*>(raw), 4);
        }
        if (ok == LB_OK) ok = lb_alloc(ctx, 4 * 15, &t_add);
        if (ok == LB_OK) ok = lb_alloc(ctx, 8 * 16, &t_mul);
        if (ok == LB_OK) ok = lb_alloc(ctx, 12 * 7, &t_inp);
        // consumers: c = a * b reads 6, 7; d = c + w reads 3, 8; e = c * d reads 3, 4 (identity index expressions)
        const int reads[6] = {6, 7, 3, 8, 3, 4};
        for (int k = 0; k < 6 && ok == LB_OK; ++k) ok = lb_trace_count_uses(ctx, uses[reads[k]], nullptr, 4);
        auto op = [&](int kind, uint32_t node, uint32_t lhs, uint32_t rhs, uint32_t* rows, uint64_t row0) {
            lb_trace_op_desc d{};
            d.op = kind;
            d.node_id = node;
            d.n = 4;
            d.d_out_mult = uses[node];
            d.d_rows = rows;
            d.row0 = row0;
            if (kind == LB_OP_INPUTS) {
                d.d_lhs = reinterpret_cast<const int32_t*>(val[node]);
            } else {
                d.lhs_id = lhs;
                d.rhs_id = rhs;
                d.d_lhs = reinterpret_cast<const int32_t*>(val[lhs]);
                d.d_rhs = reinterpret_cast<const int32_t*>(val[rhs]);
                d.d_out = reinterpret_cast<int32_t*>(val[node]);
            }
            return lb_trace_op(ctx, &d);
        };
        for (uint32_t t = 0; t < 3 && ok == LB_OK; ++t) ok = op(LB_OP_INPUTS, 6 + t, 0, 0, t_inp, 4 * t);
        if (ok == LB_OK) ok = op(LB_OP_MUL, 3, 6, 7, t_mul, 0);
        if (ok == LB_OK) ok = op(LB_OP_ADD, 4, 3, 8, t_add, 0);
        if (ok == LB_OK) ok = op(LB_OP_MUL, 5, 3, 4, t_mul, 4);
        if (ok != LB_OK) {
            std::fprintf(stderr, "device gen_trace: %d: %s\n", ok, lb_last_error(ctx));
            lb_ctx_destroy(ctx);
            return 1;
        }
        tables[0].rows = t_add; tables[1].rows = t_mul; tables[2].rows = t_inp;
        for (auto& t : tables) t.rows_on_device = 1;
        std::printf("trace tables generated on the device\n");
    }
    uint8_t* proof = nullptr;
    size_t len = 0;
    int rc = lb_prove(ctx, tables, 3, nullptr, &proof, &len);  // NULL config: PcsConfig::default(), 17 claim slots
    if (rc != LB_OK) {
        std::fprintf(stderr, "lb_prove: %d: %s\n", rc, lb_last_error(ctx));
        lb_ctx_destroy(ctx);
        return 1;
    }
    std::printf("proof: %zu bytes\n", len);
    int status = 0;
    if (fixture) {
        FILE* f = std::fopen(fixture, "rb");
        if (!f) {
            std::fprintf(stderr, "cannot open %s\n", fixture);
            status = 3;
        } else {
            std::vector<uint8_t> want;
            uint8_t buf[4096];
            size_t n;
            while ((n = std::fread(buf, 1, sizeof buf, f)) > 0) want.insert(want.end(), buf, buf + n);
            std::fclose(f);
            bool same = want.size() == len && std::memcmp(want.data(), proof, len) == 0;
            std::printf("fixture %s: %s\n", fixture, same ? "identical" : "DIFFERENT");
            status = same ? 0 : 4;
        }
    }
    lb_free_host(proof);
    lb_ctx_destroy(ctx);
    return status;
}
