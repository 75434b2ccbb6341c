"""More of the reference's own deterministic tests, transcribed as data (no math/rand inputs): every case below is a
test the reference runs against its OWN optimised kernels, so the oracle (CPU) and the CUDA path (GPU, through the C
ABI mirrors) must pass it as well.  They widen the oracle's pin beyond tests/golden/reference_kats.json.

* internal/simd/simd_unroll_test.go:13-118   odd / exact / 16 / 128-element vectors vs the sequential scalar loop, 1e-4
* internal/simd/parallel_reduction_test.go:13-186  batch cosine / dot vs scalar (1e-4), 768-dim (i % 10) / 10 fixtures
                                                   (1e-3), the 4-accumulator diagonal pattern (dot == 4 +- 1e-6)
* internal/simd/coverage_increase_test.go:74-176   EuclideanBatchInto: exact 5.0 / 5.0 / ~1.732, 128-d ramp == 0,
                                                   results beyond len(vectors) untouched
* internal/simd/simd_avx2_test.go:118-162          cosine edge cases on makeTestVector data (0, exactly 1.0, 1, 2)
* internal/simd/static_dispatch_test.go:10-26, dispatch_test.go:205-257   sqrt(20), dot 20, ranges
* internal/simd/simd_compare_test.go:7-72, branchless_test.go:55-82        MatchInt64 / MatchFloat32 truth tables
* internal/simd/sq8_extra_test.go:9-46             QuantizeSQ8 {0,127,255,0,255}, ComputeBounds table

* internal/store/result_merger_test.go:10-119      MergeSortedStreams: three streams, empty channel, limit k, interleaved
* internal/store/adaptive_index_test.go:105-165,280-306   BruteForceIndex on createTestDatasetWithVectors: k = 10 of 100
                                                   ascending (closed-form answer), k > size, empty index

The scalar loops restated here (seq_*) are the reference's generic functions (internal/simd/simd.go:131-163) in
float32 arithmetic, the yardstick its tests use.
"""
import math

import numpy as np
import pytest

f32 = np.float32
L2, COS, DOT = 0, 1, 2


# ---- the reference's scalar yardsticks (simd.go:131-163, simd_test.go:13-33) in float32 arithmetic ----
def seq_l2(a, b):
    s = f32(0)
    for x, y in zip(a, b):
        d = f32(f32(x) - f32(y))
        s = f32(s + f32(d * d))
    return f32(math.sqrt(float(s)))


def seq_dot(a, b):
    s = f32(0)
    for x, y in zip(a, b):
        s = f32(s + f32(f32(x) * f32(y)))
    return s


def seq_cos(a, b):
    dot, na, nb = f32(0), f32(0), f32(0)
    for x, y in zip(a, b):
        x, y = f32(x), f32(y)
        dot = f32(dot + f32(x * y))
        na = f32(na + f32(x * x))
        nb = f32(nb + f32(y * y))
    if na == 0 or nb == 0:
        return f32(1.0)
    return f32(f32(1.0) - f32(dot / f32(math.sqrt(float(na) * float(nb)))))


SEQ = {L2: seq_l2, COS: seq_cos, DOT: seq_dot}


def make16():           # simd_unroll_test.go:243-249
    return [f32(i + 1) for i in range(16)]


def make16_offset():    # :251-257
    return [f32(i + 17) for i in range(16)]


def make128():          # :259-265   float32(i) * 0.1 in float32
    return [f32(f32(i) * f32(0.1)) for i in range(128)]


def make128_offset():   # :267-273
    return [f32(f32(f32(i) * f32(0.1)) + f32(1.0)) for i in range(128)]


def make128_ones():     # :275-281
    return [f32(1.0)] * 128


def make_test_vector(dim, seed):   # simd_test.go:366-372: seed * float32(i+1) * 0.1
    return [f32(f32(f32(seed) * f32(i + 1)) * f32(0.1)) for i in range(dim)]


UNROLL = [  # (metric, name, a, b)
    (L2, "empty", [], []), (L2, "single", [1.0], [2.0]), (L2, "two", [1, 2], [3, 4]), (L2, "three", [1, 2, 3], [4, 5, 6]),
    (L2, "four_exact", [1, 2, 3, 4], [5, 6, 7, 8]), (L2, "five", [1, 2, 3, 4, 5], [6, 7, 8, 9, 10]),
    (L2, "eight_exact", list(range(1, 9)), list(range(9, 17))), (L2, "nine", list(range(1, 10)), list(range(10, 19))),
    (L2, "sixteen", make16(), make16_offset()), (L2, "128_dim", make128(), make128_offset()),
    (L2, "zeros", [0, 0, 0, 0], [0, 0, 0, 0]), (L2, "negative", [-1, -2, -3, -4], [1, 2, 3, 4]),
    (COS, "empty", [], []), (COS, "single", [1.0], [2.0]), (COS, "four_exact", [1, 2, 3, 4], [5, 6, 7, 8]),
    (COS, "five", [1, 2, 3, 4, 5], [6, 7, 8, 9, 10]), (COS, "eight_exact", list(range(1, 9)), list(range(9, 17))),
    (COS, "128_dim", make128(), make128_offset()), (COS, "orthogonal", [1, 0, 0, 0], [0, 1, 0, 0]),
    (COS, "parallel", [1, 2, 3, 4], [2, 4, 6, 8]), (COS, "antiparallel", [1, 2, 3, 4], [-1, -2, -3, -4]),
    (DOT, "empty", [], []), (DOT, "single", [3.0], [4.0]), (DOT, "four_exact", [1, 2, 3, 4], [5, 6, 7, 8]),
    (DOT, "five", [1, 2, 3, 4, 5], [1, 1, 1, 1, 1]), (DOT, "eight_exact", list(range(1, 9)), [1] * 8),
    (DOT, "128_dim", make128(), make128_ones()), (DOT, "zeros", [0, 0, 0, 0], [1, 2, 3, 4]),
    (DOT, "negative", [-1, -2, -3, -4], [1, 2, 3, 4]),
]
RAW = {L2: "lbo_euclid_f32", COS: "lbo_cosine_f32", DOT: "lbo_dot_f32"}

# (metric, query, vectors, tolerance) -- parallel_reduction_test.go:13-69 and simd_unroll_test.go:120-147
Q8 = list(range(1, 9))
BATCHES = [
    (COS, Q8, [[8, 7, 6, 5, 4, 3, 2, 1], [1] * 8, [0, 0, 0, 0, 0, 0, 0, 1], Q8], 1e-4),
    (DOT, Q8, [[8, 7, 6, 5, 4, 3, 2, 1], [1] * 8, [0, 0, 0, 0, 0, 0, 0, 1], [2] * 8], 1e-4),
    (L2, make128(), [make128_offset(), make128(), make128_ones()], 1e-4),
]


def high_dimension_fixture():   # parallel_reduction_test.go:116-128
    dim = 768
    q = [f32(f32(i % 10) / f32(10.0)) for i in range(dim)]
    vs = [[f32(f32((i + j) % 10) / f32(10.0)) for i in range(dim)] for j in range(10)]
    return q, vs


def diagonal_pattern():         # parallel_reduction_test.go:168-186
    a, b = [0.0] * 16, [1.0] * 16
    for i in (0, 5, 10, 15):
        a[i] = 1.0
    return a, b


# literal expectations: (name, ref, metric, a, b, expected or None, tol, (lo, hi) or None)
LITERALS = [
    ("static_dispatch_sqrt20", "internal/simd/static_dispatch_test.go:10-26", L2, [1, 2, 3, 4], [4, 3, 2, 1], float(f32(math.sqrt(20))), 1e-4, None),
    ("function_dispatch_dot20", "internal/simd/dispatch_test.go:205-225", DOT, [1, 2, 3, 4], [4, 3, 2, 1], 20.0, 0.01, None),
    ("function_dispatch_l2_positive", "internal/simd/dispatch_test.go:205-225", L2, [1, 2, 3, 4], [4, 3, 2, 1], None, 0, (1e-30, 1e30)),
    ("function_dispatch_cos_range", "internal/simd/dispatch_test.go:205-225", COS, [1, 2, 3, 4], [4, 3, 2, 1], None, 0, (-1.0, 1.0)),
    ("batch_into_3_4_0", "internal/simd/coverage_increase_test.go:96-124", L2, [0, 0, 0], [3, 4, 0], 5.0, 0.0, None),
    ("batch_into_0_0_5", "internal/simd/coverage_increase_test.go:96-124", L2, [0, 0, 0], [0, 0, 5], 5.0, 0.0, None),
    ("batch_into_1_1_1", "internal/simd/coverage_increase_test.go:96-124", L2, [0, 0, 0], [1, 1, 1], None, 0, (1.732, 1.733)),
    ("batch_into_identical", "internal/simd/coverage_increase_test.go:74-94", L2, [1, 0, 0], [1, 0, 0], 0.0, 0.0, None),
    ("batch_into_128_ramp", "internal/simd/coverage_increase_test.go:126-152", L2,
     [f32(f32(j) * f32(0.01)) for j in range(128)], [f32(f32(j) * f32(0.01)) for j in range(128)], 0.0, 0.0, None),
    ("dot_batch_dispatch_1", "internal/simd/parallel_reduction_test.go:93-113", DOT, [1, 2, 3, 4], [1, 0, 0, 0], 1.0, 1e-4, None),
    ("dot_batch_dispatch_2", "internal/simd/parallel_reduction_test.go:93-113", DOT, [1, 2, 3, 4], [0, 1, 0, 0], 2.0, 1e-4, None),
    ("dot_batch_dispatch_10", "internal/simd/parallel_reduction_test.go:93-113", DOT, [1, 2, 3, 4], [1, 1, 1, 1], 10.0, 1e-4, None),
    ("cos_batch_dispatch_identical", "internal/simd/parallel_reduction_test.go:71-90", COS, [1, 0, 0, 0], [1, 0, 0, 0], None, 0, (-1e-4, 1e-4)),
    ("cos_batch_dispatch_orthogonal", "internal/simd/parallel_reduction_test.go:71-90", COS, [1, 0, 0, 0], [0, 1, 0, 0], 1.0, 1e-4, None),
    ("accumulator_independence", "internal/simd/parallel_reduction_test.go:168-186", DOT, *diagonal_pattern(), 4.0, 1e-6, None),
    ("avx2_cos_identical_128", "internal/simd/simd_avx2_test.go:122-129", COS, make_test_vector(128, 1.0), make_test_vector(128, 1.0), 0.0, 1e-5, None),
    ("avx2_cos_zero_vector_exact", "internal/simd/simd_avx2_test.go:131-138", COS, [0.0] * 128, make_test_vector(128, 1.0), 1.0, 0.0, None),
    ("avx2_cos_orthogonal_8", "internal/simd/simd_avx2_test.go:140-148", COS, [1, 0, 0, 0, 0, 0, 0, 0], [0, 1, 0, 0, 0, 0, 0, 0], 1.0, 1e-5, None),
    ("avx2_cos_opposite_64", "internal/simd/simd_avx2_test.go:150-161", COS, make_test_vector(64, 1.0), [-x for x in make_test_vector(64, 1.0)], 2.0, 2e-5, None),
]

# simd_compare_test.go:7-72 and branchless_test.go:55-82; op = simd.CompareOp (simd.go:38-45): Eq, Neq, Gt, Ge, Lt, Le
EQ, NEQ, GT, GE, LT, LE = range(6)
I64_SRC = [10, 20, 30, 40, 50, 10, 50]
F32_SRC = [1.5, 2.5, 3.5, 4.5, 5.5]
EXTREME = [0, 1, -1, 42, -9223372036854775808, 9223372036854775807]
MATCH = [
    ("i64", I64_SRC, 10, EQ, [1, 0, 0, 0, 0, 1, 0]), ("i64", I64_SRC, 10, NEQ, [0, 1, 1, 1, 1, 0, 1]),
    ("i64", I64_SRC, 25, GT, [0, 0, 1, 1, 1, 0, 1]), ("i64", I64_SRC, 30, GE, [0, 0, 1, 1, 1, 0, 1]),
    ("i64", I64_SRC, 30, LT, [1, 1, 0, 0, 0, 1, 0]), ("i64", I64_SRC, 30, LE, [1, 1, 1, 0, 0, 1, 0]),
    ("f32", F32_SRC, 2.5, EQ, [0, 1, 0, 0, 0]), ("f32", F32_SRC, 2.5, NEQ, [1, 0, 1, 1, 1]),
    ("f32", F32_SRC, 3.0, GT, [0, 0, 1, 1, 1]), ("f32", F32_SRC, 3.5, GE, [0, 0, 1, 1, 1]),
    ("f32", F32_SRC, 3.5, LT, [1, 1, 0, 0, 0]), ("f32", F32_SRC, 3.5, LE, [1, 1, 1, 0, 0]),
    ("i64", EXTREME, 42, EQ, [0, 0, 0, 1, 0, 0]), ("i64", EXTREME, 42, NEQ, [1, 1, 1, 0, 1, 1]),
]
SQ8_QUANT = ([0.0, 0.5, 1.0, -1.0, 2.0], 0.0, 1.0, [0, 127, 255, 0, 255])          # sq8_extra_test.go:9-22
BOUNDS = [([1.0, 5.0, -2.0, 3.0], -2.0, 5.0), ([42.0], 42.0, 42.0), ([], 0.0, 0.0), ([1.0, 1.0, 1.0], 1.0, 1.0)]  # :24-46


def _check_literal(got, expected, tol, rng, name):
    if expected is None:
        assert rng[0] <= got <= rng[1], (name, got)
    else:
        assert abs(got - expected) <= tol, (name, got, expected)


# ------------------------------------------------------------------------------------------ CPU: the oracle
def test_citations_present():
    """Every cited range exists in the reference tree when it is mounted (build container only)."""
    import os
    if not os.path.isdir("/root/reference"):
        pytest.skip("reference tree absent")
    for _n, ref, *_ in LITERALS:
        path, rng = ref.split(":")
        lo, hi = (int(x) for x in rng.split("-"))
        lines = open(os.path.join("/root/reference", path)).read().split("\n")[lo - 1:hi]
        assert any(s.strip().startswith("func Test") or "t.Run(" in s for s in lines), ref


@pytest.mark.parametrize("metric,name,a,b", UNROLL, ids=[f"{'l2 cos dot'.split()[m]}_{n}" for m, n, _a, _b in UNROLL])
def test_oracle_unrolled_cases(oracle, metric, name, a, b):
    a, b = np.array(a, f32), np.array(b, f32)
    got = float(oracle.raw(RAW[metric], a, b))
    assert abs(got - float(SEQ[metric](a, b))) < 1e-4     # almostEqual, simd_unroll_test.go:238-241


@pytest.mark.parametrize("case", range(len(BATCHES)))
def test_oracle_batch_unrolled_cases(oracle, case):
    metric, q, vs, tol = BATCHES[case]
    q = np.array(q, f32)
    for v in vs:
        v = np.array(v, f32)
        assert abs(float(oracle.raw(RAW[metric], q, v)) - float(SEQ[metric](q, v))) <= tol


def test_oracle_high_dimension_fixture(oracle):
    q, vs = high_dimension_fixture()
    q = np.array(q, f32)
    flat = np.array(vs, f32)
    for metric in (L2, COS, DOT):
        got = oracle.batch_flat(metric, q, flat)
        for i, v in enumerate(flat):
            want = float(SEQ[metric](q, v))
            g = float(got[i]) if metric != DOT else float(oracle.raw(RAW[DOT], q, v))
            assert abs(g - want) <= 1e-3, (metric, i, g, want)


@pytest.mark.parametrize("case", LITERALS, ids=[c[0] for c in LITERALS])
def test_oracle_literal_kats(oracle, case):
    name, _ref, metric, a, b, expected, tol, rng = case
    got = float(oracle.raw(RAW[metric], np.array(a, f32), np.array(b, f32)))
    _check_literal(got, expected, tol, rng, name)


def _match(src, val, op):
    """simd.go:585-690 restated: one byte per element."""
    s = np.asarray(src)
    return [int(x) for x in {EQ: s == val, NEQ: s != val, GT: s > val, GE: s >= val, LT: s < val, LE: s <= val}[op]]


@pytest.mark.parametrize("case", range(len(MATCH)))
def test_match_truth_tables_restated(case):
    kind, src, val, op, want = MATCH[case]
    src = np.array(src, np.int64 if kind == "i64" else f32)
    assert _match(src, val, op) == want


def test_oracle_sq8_reference_kats(oracle):
    src, lo, hi, want = SQ8_QUANT
    assert oracle.quantize_sq8(np.array(src, f32), lo, hi).tolist() == want
    for vec, mn, mx in BOUNDS:
        assert oracle.compute_bounds(np.array(vec, f32)) == (mn, mx)


# ------------------------------------------------------------------------------------------ GPU: the CUDA path
@pytest.mark.gpu
def test_gpu_unrolled_and_literal_cases():
    from longbow_b200 import simd
    fn = {L2: simd.EuclideanDistance, COS: simd.CosineDistance, DOT: simd.DotProduct}
    for metric, name, a, b in UNROLL:
        got = float(fn[metric](np.array(a, f32), np.array(b, f32)))
        assert abs(got - float(SEQ[metric](a, b))) < 1e-4, (metric, name, got)
    for name, _ref, metric, a, b, expected, tol, rng in LITERALS:
        _check_literal(float(fn[metric](np.array(a, f32), np.array(b, f32))), expected, tol, rng, name)


@pytest.mark.gpu
def test_gpu_batch_cases():
    from longbow_b200 import simd
    fn = {L2: simd.EuclideanDistanceBatch, COS: simd.CosineDistanceBatch, DOT: simd.DotProductBatch}
    q768, v768 = high_dimension_fixture()
    for metric, q, vs, tol in BATCHES + [(m, q768, v768, 1e-3) for m in (L2, COS, DOT)]:
        q = np.array(q, f32)
        rows = [np.array(v, f32) for v in vs]
        res = np.full(len(rows), 999.0, f32)
        fn[metric](q, rows, res)
        for i, v in enumerate(rows):
            assert abs(float(res[i]) - float(SEQ[metric](q, v))) <= tol, (metric, i, res[i])
    # coverage_increase_test.go:154-176: a results slot beyond len(vectors) is not touched (shown on a non-strict
    # batch form; EuclideanDistanceBatch itself requires equal lengths, batch_operations.go:30-32)
    res = np.array([999.0, 888.0], f32)
    simd.DotProductF16Batch(np.array([1, 2, 3], np.float16), [np.array([1, 2, 3], np.float16)], res)
    assert res[0] == 14.0 and res[1] == 888.0


@pytest.mark.gpu
def test_gpu_match_truth_tables():
    from longbow_b200 import store
    for kind, src, val, op, want in MATCH:
        col = np.array(src, np.int64 if kind == "i64" else f32)
        bm = store.GenerateFilterBitset(col, op, val)
        got = [(int(bm[i // 64]) >> (i % 64)) & 1 for i in range(col.size)]
        assert got == want, (kind, val, op, got)


@pytest.mark.gpu
def test_gpu_sq8_reference_kats():
    from longbow_b200 import simd
    src, lo, hi, want = SQ8_QUANT
    dst = np.zeros(len(src), np.uint8)
    simd.QuantizeSQ8(np.array(src, f32), dst, lo, hi)
    assert dst.tolist() == want
    for vec, mn, mx in BOUNDS:
        assert simd.ComputeBounds(np.array(vec, f32)) == (mn, mx)


# ------------------------------------------------------------------ shard merge + brute-force index reference cases
# internal/store/result_merger_test.go:10-119: sorted per-stream results merged by score.  Streams as padded
# [parts, 1, k_in] lists with label -1 / MaxFloat32 padding, the layout of the per-shard top lists.
MAXF = float(np.finfo(np.float32).max)
MERGE_CASES = [
    # (name, streams [(id, score), ...], k, expected ids)
    ("three_streams", [[(1, 0.1), (2, 0.4), (3, 0.7)], [(4, 0.2), (5, 0.5)], [(6, 0.3), (7, 0.6), (8, 0.9)]], 10,
     [1, 4, 6, 2, 5, 7, 3, 8]),
    ("empty_channel", [[], [(1, 0.5)]], 5, [1]),
    ("limit_k", [[(1, 0.1), (2, 0.2), (3, 0.3)]], 2, [1, 2]),
    ("interleaved", [[(1, 0.1), (2, 0.9)], [(3, 0.2), (4, 0.5)]], 10, [1, 3, 4, 2]),
]


def _merge_inputs(streams):
    k_in = max(1, max(len(s) for s in streams))
    d = np.full((len(streams), 1, k_in), MAXF, f32)
    l = np.full((len(streams), 1, k_in), -1, np.int64)
    for p, s in enumerate(streams):
        for j, (i, sc) in enumerate(s):
            d[p, 0, j], l[p, 0, j] = f32(sc), i
    return d, l


def _check_merge(od, ol, streams, k, want_ids):
    score = {i: f32(sc) for s in streams for i, sc in s}
    got = [int(x) for x in ol[0] if x >= 0]
    assert got == want_ids[:k]
    assert [f32(x) for x in od[0][:len(got)]] == [score[i] for i in got]      # scores travel unchanged
    assert all(int(x) == -1 for x in ol[0][len(got):])


@pytest.mark.parametrize("case", MERGE_CASES, ids=[c[0] for c in MERGE_CASES])
def test_oracle_merge_sorted_streams(oracle, case):
    _name, streams, k, want = case
    d, l = _merge_inputs(streams)
    od, ol = oracle.merge(d, l, k)
    _check_merge(od, ol, streams, k, want)


def bf_dataset(n):   # createTestDatasetWithVectors, internal/store/adaptive_index_test.go:280-306: float32(i*4+d) * 0.01
    return np.array([[f32(f32(i * 4 + d) * f32(0.01)) for d in range(4)] for i in range(n)], f32).reshape(n, 4)


# the distance to query (0.1, 0.2, 0.3, 0.4) is a parabola in the row number with its vertex at 5.875: no ties
BF_WANT = [6, 5, 7, 4, 8, 3, 9, 2, 10, 1]


def test_oracle_brute_force_reference_cases(oracle):
    """adaptive_index_test.go:105-152: k = 10 over 100 rows comes back ascending; k = 100 over 5 rows gives 5 results."""
    q = np.array([[0.1, 0.2, 0.3, 0.4]], f32)
    d, l = oracle.search(L2, bf_dataset(100), q, 10)
    assert l[0].tolist() == BF_WANT and np.all(np.diff(d[0]) >= 0)
    d, l = oracle.search(L2, bf_dataset(5), np.array([[1.0, 0.0, 0.0, 0.0]], f32), 100)
    assert sorted(int(x) for x in l[0] if x >= 0) == [0, 1, 2, 3, 4] and int((l[0] >= 0).sum()) == 5


@pytest.mark.gpu
def test_gpu_merge_and_brute_force_reference_cases():
    from longbow_b200 import store
    for _name, streams, k, want in MERGE_CASES:
        d, l = _merge_inputs(streams)
        od, ol = store.MergeShardResults(d, l, k)
        _check_merge(od, ol, streams, k, want)
    bf = store.BruteForceIndex(4)
    assert bf.Len() == 0 and not bf.SearchVectors(np.array([1.0, 2.0, 3.0, 4.0], f32), 10)   # :154-165 empty index
    bf.AddBatch(bf_dataset(100))
    res = bf.SearchVectors(np.array([0.1, 0.2, 0.3, 0.4], f32), 10)
    assert [r.ID for r in res] == BF_WANT and all(a.Score <= b.Score for a, b in zip(res, res[1:]))
    bf.Close()
    bf = store.BruteForceIndex(4)
    bf.AddBatch(bf_dataset(5))
    res = bf.SearchVectors(np.array([1.0, 0.0, 0.0, 0.0], f32), 100)                        # :133-152 k > index size
    assert len(res) == 5 and sorted(r.ID for r in res) == [0, 1, 2, 3, 4]
    bf.Close()


# ------------------------------------------------------------------ a few more literals + the mirrors' error behaviour
MORE_LITERALS = [
    # internal/simd/jit_test.go:26-43: q = {1,1,1} against three rows: 0, 1.73205, 13 (+- 1e-4)
    ("jit_batch_0", L2, [1, 1, 1], [1, 1, 1], 0.0, 1e-4), ("jit_batch_sqrt3", L2, [1, 1, 1], [2, 2, 2], 1.73205, 1e-4),
    ("jit_batch_13", L2, [1, 1, 1], [4, 5, 13], 13.0, 1e-4),
    # internal/simd/jit_test.go:10-24 / docs: sqrt(27)
    ("jit_sqrt27", L2, [1, 2, 3], [4, 5, 6], float(f32(math.sqrt(27))), 1e-4),
    # internal/simd/simd_fma_portable_test.go:57-70,166-179: dot 70, orthogonal cosine 1.0 (+- 1e-5)
    ("fma_dot_70", DOT, [1, 2, 3, 4], [5, 6, 7, 8], 70.0, 1e-5), ("fma_cos_orthogonal", COS, [1, 0, 0], [0, 1, 0], 1.0, 1e-5),
]


@pytest.mark.parametrize("case", MORE_LITERALS, ids=[c[0] for c in MORE_LITERALS])
def test_oracle_more_literals(oracle, case):
    _name, metric, a, b, expected, tol = case
    assert abs(float(oracle.raw(RAW[metric], np.array(a, f32), np.array(b, f32))) - expected) <= tol


def test_mirror_argument_errors_need_no_gpu():
    """The argument checks of the simd mirrors are host logic and match the reference's error behaviour: a length
    mismatch is an error (simd_test.go:124-129,215-220; distance_functions.go:18-20), results / vectors mismatch is
    an error for the Euclidean batch forms (batch_operations.go:18-20,30-32,92-94), a too small results slice for the
    cosine / dot forms (:131-157), empty inputs return 0 / 1 / 0 without a kernel (distance_functions.go:21-23,51-53,63-65)."""
    from longbow_b200 import simd
    a, b = np.array([1, 2, 3], f32), np.array([1, 2], f32)
    for fn in (simd.EuclideanDistance, simd.CosineDistance, simd.DotProduct, simd.EuclideanDistanceF16,
               simd.CosineDistanceF16, simd.DotProductF16):
        with pytest.raises(simd.SimdError, match="length mismatch"):
            fn(a, b)
    e = np.zeros(0, f32)
    assert simd.EuclideanDistance(e, e) == 0.0 and simd.CosineDistance(e, e) == 1.0 and simd.DotProduct(e, e) == 0.0
    rows = [np.array([1, 2, 3], f32)] * 2
    for fn in (simd.EuclideanDistanceBatch, simd.EuclideanDistanceVerticalBatch):
        with pytest.raises(simd.SimdError, match="vectors and results length mismatch"):
            fn(a, rows, np.zeros(3, f32))
        fn(a, [], np.zeros(0, f32))                      # nothing to do, no kernel
    for fn in (simd.CosineDistanceBatch, simd.DotProductBatch):   # these two only need room (:131-157)
        with pytest.raises(simd.SimdError, match="results slice too small"):
            fn(a, rows, np.zeros(1, f32))
        fn(a, [], np.zeros(0, f32))
    with pytest.raises(simd.SimdError, match="results slice too small"):
        simd.DotProductF16Batch(np.array([1, 2, 3], np.float16), [np.array([1, 2, 3], np.float16)] * 2, np.zeros(1, f32))
    with pytest.raises(simd.SimdError, match="results length mismatch"):
        simd.EuclideanDistanceBatchFlat(a, np.zeros(6, f32), 2, 3, np.zeros(3, f32))
    with pytest.raises(simd.SimdError, match="flatVectors too small"):
        simd.EuclideanDistanceBatchFlat(a, np.zeros(5, f32), 2, 3, np.zeros(2, f32))
    with pytest.raises(simd.SimdError, match="query dimension mismatch"):
        simd.EuclideanDistanceBatchFlat(b, np.zeros(6, f32), 2, 3, np.zeros(2, f32))
    simd.EuclideanDistanceBatchFlat(a, np.zeros(0, f32), 0, 3, np.zeros(0, f32))   # numVectors == 0: no-op


# ------------------------------------------------------------------ conjunctions of predicates (FilterEvaluator)
# internal/query/filter_evaluator_test.go:70-193: record batches of int64 / float32 columns, filters AND-combined,
# expected row indices as literals.  Operators as the parser maps them onto simd.CompareOp.
FE_OPS = {"=": EQ, "!=": NEQ, ">": GT, ">=": GE, "<": LT, "<=": LE}
FE_BATCH8 = {"id": np.arange(1, 9, dtype=np.int64), "category": np.array([1, 1, 2, 2, 1, 1, 2, 2], np.int64)}
FE_BATCH10 = {"id": np.arange(1, 11, dtype=np.int64), "category": np.array([1, 1, 2, 2, 1, 1, 2, 2, 1, 1], np.int64),
              "value": np.arange(1, 11).astype(f32)}
FE_CASES = [
    (FE_BATCH8, [("id", ">=", 3), ("category", "=", 1)], [4, 5]),                                  # :70-106
    (FE_BATCH10, [("id", ">=", 5)], [4, 5, 6, 7, 8, 9]),                                           # :124-134
    (FE_BATCH10, [("id", ">=", 3), ("category", "=", 1)], [4, 5, 8, 9]),                           # :136-147
    (FE_BATCH10, [("id", ">=", 3), ("category", "=", 1), ("value", "<=", 6.0)], [4, 5]),           # :149-161
    (FE_BATCH10, [], list(range(10))),                                                             # :174-181 no filters
    (FE_BATCH10, [("id", ">", 100)], []),                                                          # :183-192
]


def _numpy_filter(column, op, value, bitmap=None):
    """Same contract as longbow_b200.store.GenerateFilterBitset, on the CPU (validates the test body without a GPU)."""
    col = np.asarray(column)
    hit = {EQ: col == value, NEQ: col != value, GT: col > value, GE: col >= value, LT: col < value, LE: col <= value}[op]
    words = np.packbits(np.pad(hit, (0, (-col.size) % 64)), bitorder="little").view(np.uint64).copy()
    return words if bitmap is None else (words & np.asarray(bitmap, np.uint64))


def _run_filter_cases(filter_fn):
    for batch, filters, want in FE_CASES:
        n = len(next(iter(batch.values())))
        bm = None
        for field, op, value in filters:
            bm = filter_fn(batch[field], FE_OPS[op], value, bm)
        rows = list(range(n)) if bm is None else [i for i in range(n) if (int(bm[i // 64]) >> (i % 64)) & 1]
        assert rows == want, (filters, rows)


def test_filter_evaluator_cases_restated():
    _run_filter_cases(_numpy_filter)


@pytest.mark.gpu
def test_gpu_filter_evaluator_cases():
    from longbow_b200 import store
    _run_filter_cases(lambda col, op, value, bm: store.GenerateFilterBitset(col, op, value, bitmap=bm))
