"""GPU parity for the components beyond Add/Mul/Inputs/reductions (SURVEY 8 a6/a7 over all 17 components):
Recip, Sqrt, Rem, Sin/Exp2/Log2 with their lookup-table components, LessThan with the range-check table.
Kernel level (LogUp columns, constraint quotients on the evaluation domain - including the extended domain of a
LUT consumer whose table is larger than its trace) and whole proofs, bytes equal to the oracle prover's."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pie as piemod
from oracle import air, cfft as ocfft, prover as oprover, verifier as overifier
from oracle.circle import CanonicCoset
from oracle.fields import P, QM31
from oracle.proof import from_bincode, to_bincode
from test_gpu_prover import _assert_same_proof, _oracle_transcript, _rand_q

KIND = {"recip": 7, "sqrt": 8, "rem": 9, "sin": 10, "exp2": 11, "log2": 12, "sin_lookup": 13, "exp2_lookup": 14,
        "log2_lookup": 15, "less_than": 16, "range_check_lookup": 17}
LUT_NAMES = ["sin", "exp2", "log2", "range_check"]


@pytest.fixture(scope="module")
def be():
    from luminair_b200.backend import CudaBackend
    b = CudaBackend(0)
    yield b
    b.close()


def _meta(pre):
    return [(cid, len(v).bit_length() - 1) for cid, v in pre]


@pytest.fixture(scope="module")
def graph():
    return piemod.all_components_graph(n=40, seed=9)


@pytest.mark.parametrize("name", sorted(KIND))
def test_logup_and_constraint_quotients(be, graph, name):
    from luminair_b200.backend import ColumnBatch
    pie, pre = graph
    pre = overifier.sort_preprocessed(pre)
    pre_vals = {cid: np.asarray(v, dtype=np.uint64) for cid, v in pre}
    rng = np.random.Generator(np.random.PCG64(77))
    node = air.RelationElements(_rand_q(rng), _rand_q(rng), 2)
    lookups = {n: air.RelationElements(_rand_q(rng), _rand_q(rng), 1 if n == "range_check" else 2) for n in LUT_NAMES}
    rels = [(node.z.tup(), node.alpha.tup())] + [(lookups[n].z.tup(), lookups[n].alpha.tup()) for n in LUT_NAMES]
    cls = overifier.SLOT_EVALS[oprover.SLOT_OF[name]]
    padded = oprover.pad_table(np.asarray(dict(pie)[name], dtype=np.uint64), cls.padding_row())
    lg = padded.shape[0].bit_length() - 1
    cols = [np.ascontiguousarray(padded[:, j]) for j in range(padded.shape[1])]
    want_cols, want_claim = air.gen_interaction_trace(cls, cols, lg, node, lookups, pre_vals)

    main = ColumnBatch(be.upload(np.stack(cols).astype(np.uint32).reshape(-1)), len(cols), lg)
    n_ic = 4 * cls.n_interaction
    inter = ColumnBatch(be.alloc(n_ic << lg), n_ic, lg)
    lut_ids = cls.preprocessed_ids() if hasattr(cls, "preprocessed_ids") else []
    lut_bufs = [be.upload(pre_vals[i].astype(np.uint32)) for i in lut_ids]
    claimed = be.logup_interaction_trace_lut(KIND[name], main, inter, rels, [b.ptr for b in lut_bufs])
    assert tuple(int(x) for x in claimed) == want_claim.tup()
    got = be.download(inter.buf).reshape(n_ic, -1)
    for k in range(n_ic):
        assert np.array_equal(got[k], want_cols[k].astype(np.uint32)), f"interaction column {k}"

    # constraint quotients on the component's evaluation domain
    lut_logs = overifier.lut_log_sizes(_meta(pre))
    alloc = air.TraceLocationAllocator([cid for cid, _ in pre])
    comp = air.FrameworkComponent(alloc, cls(lg, node, lookups, lut_logs), want_claim)
    eval_log = comp.max_constraint_log_degree_bound()
    if name in ("sin", "exp2", "log2", "less_than"):
        assert eval_log == max(lg, lut_logs[cls.lut if name != "less_than" else "range_check"]) + 1
    dom = CanonicCoset(eval_log).circle_domain()

    def ext(c):
        c = np.asarray(c, dtype=np.uint64)
        return ocfft.evaluate(ocfft.interpolate(c, CanonicCoset(len(c).bit_length() - 1).circle_domain()), dom)

    main_ev = [ext(c) for c in cols]
    inter_ev = [ext(c) for c in want_cols]
    pre_polys = [ocfft.interpolate(v, CanonicCoset(len(v).bit_length() - 1).circle_domain()) for _, v in [(c, pre_vals[c]) for c, _ in pre]]
    pre_lde = [ocfft.evaluate(p, CanonicCoset(len(p).bit_length()).circle_domain()) for p in pre_polys]
    pows = [_rand_q(rng) for _ in range(comp.n_constraints)]
    main_polys = [ocfft.interpolate(c, CanonicCoset(lg).circle_domain()) for c in cols]
    inter_polys = [ocfft.interpolate(np.asarray(c, dtype=np.uint64), CanonicCoset(lg).circle_domain()) for c in want_cols]
    lde = {0: pre_lde, 1: [ocfft.evaluate(p, CanonicCoset(lg + 1).circle_domain()) for p in main_polys],
           2: [ocfft.evaluate(p, CanonicCoset(lg + 1).circle_domain()) for p in inter_polys]}
    want = comp.evaluate_constraint_quotients_on_domain(lde, pows, {0: pre_polys, 1: main_polys, 2: inter_polys})

    ml = ColumnBatch(be.upload(np.stack(main_ev).astype(np.uint32).reshape(-1)), len(cols), eval_log)
    il = ColumnBatch(be.upload(np.stack(inter_ev).astype(np.uint32).reshape(-1)), n_ic, eval_log)
    lut_ev = [be.upload(ext(pre_vals[i]).astype(np.uint32)) for i in lut_ids]
    acc = be.alloc(4 << eval_log)
    be.constraint_quotients_lut(KIND[name], ml, il, lg, eval_log, rels, want_claim.tup(), [p.tup() for p in pows],
                                [acc.at(k << eval_log) for k in range(4)], [b.ptr for b in lut_ev])
    got = be.download(acc).reshape(4, -1)
    for k in range(4):
        assert np.array_equal(got[k], np.asarray(want.c[k], dtype=np.uint32)), f"coordinate {k}"


def _prove_both(be, pie, pre):
    from luminair_b200.prover import prove
    lp, digests = _oracle_transcript(lambda: oprover.prove(pie, preprocessed=pre))
    got = prove(pie, backend=be, preprocessed=pre)
    _assert_same_proof(be, got, to_bincode(lp), digests)
    overifier.verify(from_bincode(got), preprocessed=_meta(pre))
    return got


def test_all_components_proof_bytes(be, graph):
    """One proof over all 17 components: LUT tables of 2^8 .. 2^15 rows next to 64-row traces."""
    pie, pre = graph
    assert len(pie) == 17
    _prove_both(be, pie, pre)


def test_lut_smaller_than_trace_proof_bytes(be):
    g = piemod.GraphTrace()
    rng = np.random.Generator(np.random.PCG64(11))
    x = g.input(rng.integers(-8, 8, 300))
    y = g.exp2(x)
    g.less_than(y, x)
    pie, pre = g.finish()
    _prove_both(be, pie, pre)


def test_mlp_graph_proof_bytes(be):
    """BASELINE cfg 4 shape: Linear 2-64-64-1 with tanh (Mul/SumReduce/Add/Exp2+LUT/Recip), synthetic weights."""
    pie, pre = piemod.mlp_graph()
    assert [k for k, _ in pie] == ["add", "mul", "recip", "sum_reduce", "exp2", "exp2_lookup", "inputs"]
    assert dict(pie)["mul"].shape[0] == 2 * 64 + 64 * 64 + 64 + 4 * 64
    _prove_both(be, pie, pre)


@pytest.mark.parametrize("kind,col", [("recip", 9), ("sqrt", 8), ("rem", 12), ("less_than", 13)])
def test_broken_witness_is_rejected(be, kind, col):
    from luminair_b200.prover import ProvingError, prove
    pie, pre = piemod.all_components_graph(n=16, seed=5)
    pie = [(k, v.copy()) for k, v in pie]
    tbl = dict(pie)[kind]
    tbl[1, col] = (int(tbl[1, col]) + 1) % P
    with pytest.raises(ProvingError):
        prove(pie, backend=be, preprocessed=pre)


def test_missing_lut_is_an_error(be):
    from luminair_b200._lib import LuminairB200Error
    from luminair_b200.prover import prove
    pie, pre = piemod.all_components_graph(n=16, seed=5)
    with pytest.raises(LuminairB200Error):
        prove(pie, backend=be, preprocessed=[c for c in pre if not c[0].startswith("exp2")])


def test_wide_graph_proof_bytes_and_acceptance(be):
    """The headline trace shape (Add + Mul + Rem + SumReduce tables over two inputs): bytes vs the oracle at 2^6 rows,
    oracle-verifier acceptance at 2^16 rows."""
    from luminair_b200.prover import prove
    small = piemod.wide_graph(6)
    _prove_both(be, small, ())
    big = piemod.wide_graph(16)
    assert sum(v.shape[1] for k, v in big if k != "inputs") == 61
    overifier.verify(from_bincode(prove(big, backend=be)))
