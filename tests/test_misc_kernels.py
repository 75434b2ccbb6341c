"""SQ8 quantisation, FindNearestCentroid, the single-pair / F16 / vertical simd mirrors and the multi-batch filter
scatter (SURVEY.md 8 f1 / f2 remainder): oracle restatements checked on the CPU, GPU kernels against them."""
import numpy as np
import pytest


def test_oracle_sq8_known_answers(oracle):
    # sq8.go:70-86 by hand: min=0, max=255 -> scale 1: identity with truncation and clamping
    src = np.array([-3.0, 0.0, 0.4, 0.6, 127.9, 255.0, 300.0], np.float32)
    assert oracle.quantize_sq8(src, 0.0, 255.0).tolist() == [0, 0, 0, 0, 127, 255, 255]
    assert oracle.quantize_sq8(src, 5.0, 5.0).tolist() == [0] * 7          # max == min -> scale 0
    assert oracle.compute_bounds(np.array([], np.float32)) == (0.0, 0.0)    # sq8.go:89-91
    assert oracle.compute_bounds(src) == (-3.0, 300.0)
    q = oracle.quantize_sq8(np.linspace(-1, 1, 11, dtype=np.float32), -1.0, 1.0)
    back = oracle.dequantize_sq8(q, -1.0, 1.0)
    assert np.abs(back - np.linspace(-1, 1, 11)).max() <= 2.0 / 255 + 1e-6


def test_oracle_find_nearest_centroid_rules(oracle):
    rng = np.random.default_rng(0)
    q = rng.standard_normal(8).astype(np.float32)
    for k in (1, 8, 9, 256):
        cent = rng.standard_normal((k, 8)).astype(np.float32)
        cent[k // 2] = cent[0]  # a tie: the first minimum wins
        i, d = oracle.find_nearest_centroid(q, cent)
        d2 = ((cent - q) ** 2).sum(1)
        assert i == int(np.argmin(d2))
        want = d2.min() if k <= 8 else np.sqrt(d2.min())   # k <= 8 returns the SQUARED distance (simd.go:283-296)
        assert abs(d - want) <= 1e-4 * max(1.0, want)


@pytest.mark.gpu
def test_sq8_and_centroid_kernels_bit_exact(oracle):
    from longbow_b200 import simd
    rng = np.random.default_rng(1)
    src = (rng.standard_normal(10007) * 3).astype(np.float32)
    mn, mx = simd.ComputeBounds(src)
    assert (mn, mx) == oracle.compute_bounds(src)
    for lo, hi in ((mn, mx), (-1.0, 1.0), (2.0, 2.0)):
        dst = np.zeros(src.size, np.uint8)
        simd.QuantizeSQ8(src, dst, lo, hi)
        assert np.array_equal(dst, oracle.quantize_sq8(src, lo, hi))
        assert np.array_equal(simd.DequantizeSQ8(dst, lo, hi), oracle.dequantize_sq8(dst, lo, hi))
    rows = rng.integers(0, 256, (3000, 96), dtype=np.uint8)
    q = rng.standard_normal(96).astype(np.float32)
    out = np.empty(3000, np.float32)
    simd.SQ8DequantDistanceBatch(q, rows, mn, mx, out)
    assert np.array_equal(out, oracle.sq8_dequant_distance_batch(q, rows, mn, mx))
    for k, sub in ((4, 8), (8, 3), (9, 8), (256, 8), (16, 33)):
        cent = rng.standard_normal((k, sub)).astype(np.float32)
        cent[k - 1] = cent[0]
        qq = rng.standard_normal(sub).astype(np.float32)
        assert simd.FindNearestCentroid(qq, cent, sub, k) == oracle.find_nearest_centroid(qq, cent)
    assert simd.FindNearestCentroid(q[:8], np.zeros(4, np.float32), 8, 4) == (0, simd.MaxFloat32)


@pytest.mark.gpu
def test_pair_and_f16_mirrors(oracle):
    from longbow_b200 import simd
    rng = np.random.default_rng(2)
    for dim in (1, 7, 128, 384, 769):
        a, b = rng.standard_normal(dim).astype(np.float32), rng.standard_normal(dim).astype(np.float32)
        assert simd.EuclideanDistance(a, b) == np.float32(oracle.raw("lbo_euclid_f32", a, b))
        assert simd.CosineDistance(a, b) == np.float32(oracle.raw("lbo_cosine_f32", a, b))
        assert simd.DotProduct(a, b) == np.float32(oracle.raw("lbo_dot_f32", a, b))
        ah, bh = a.astype(np.float16), b.astype(np.float16)
        assert simd.EuclideanDistanceF16(ah, bh) == np.float32(oracle.raw("lbo_euclid_f16", ah.view(np.uint16), bh.view(np.uint16)))
        assert simd.CosineDistanceF16(ah, bh) == np.float32(oracle.raw("lbo_cosine_f16", ah.view(np.uint16), bh.view(np.uint16)))
        assert simd.DotProductF16(ah, bh) == np.float32(oracle.raw("lbo_dot_f16", ah.view(np.uint16), bh.view(np.uint16)))
    assert simd.EuclideanDistance(np.zeros(0, np.float32), np.zeros(0, np.float32)) == 0.0
    assert simd.CosineDistance(np.zeros(0, np.float32), np.zeros(0, np.float32)) == 1.0
    with pytest.raises(simd.SimdError):
        simd.EuclideanDistance(np.zeros(3, np.float32), np.zeros(4, np.float32))
    q = rng.standard_normal(100).astype(np.float32)
    vecs = [rng.standard_normal(100).astype(np.float32) for _ in range(9)]
    o1, o2 = np.empty(9, np.float32), np.empty(9, np.float32)
    simd.EuclideanDistanceVerticalBatch(q, vecs, o1)
    simd.EuclideanDistanceBatch(q, vecs, o2)
    assert np.array_equal(o1, o2)


@pytest.mark.gpu
def test_filter_scatter_multi_batch(oracle):
    """dataset.go:247-300: three record batches, one indexed contiguously, one through an id map with misses."""
    from longbow_b200 import store
    rng = np.random.default_rng(3)
    n_vids = 5000
    cols = [rng.integers(0, 10, 1000).astype(np.int64), rng.integers(0, 10, 1500).astype(np.int64),
            rng.integers(0, 10, 777).astype(np.int64)]
    ids1 = rng.permutation(np.arange(2500, 4000)).astype(np.uint32)
    ids1[::7] = 0xFFFFFFFF                      # GetVectorID miss
    ids2 = rng.choice(np.arange(4000, n_vids), 777, replace=False).astype(np.uint32)
    got = store.GenerateFilterBitsetBatches([(cols[0], 100), (cols[1], ids1), (cols[2], ids2)], 0, 3, n_vids)
    want = np.zeros(n_vids, bool)
    want[100 + np.nonzero(cols[0] == 3)[0]] = True
    m1 = (cols[1] == 3) & (ids1 != 0xFFFFFFFF)
    want[ids1[m1]] = True
    want[ids2[cols[2] == 3]] = True
    from longbow_b200.gpu import pack_bitmap
    assert np.array_equal(got, pack_bitmap(want))
