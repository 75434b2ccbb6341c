"""The round-2 PQ ADC scan (csrc/pq_scan.cu): coarse integer pass over a bank-conflict-free quantised LUT ->
exact sequential fp32 sums of the survivors -> certification.  Every returned (id, distance) must be bit-equal
to the oracle's restatement of simd.ADCDistanceBatch (internal/simd/simd.go:345-355) + top-k, in every scan mode
(1 = exhaustive fp32 kernel, 2 = coarse one query per pass, 3 = coarse four queries per pass)."""
import numpy as np
import pytest

from tests.util import assert_topk_equal, random_bitmap

pytestmark = pytest.mark.gpu
L2 = 0


def _setup(rng, n, M, sub):
    cb = rng.standard_normal((M, 256, sub)).astype(np.float32)
    codes = rng.integers(0, 256, (n, M), dtype=np.uint8)
    return cb, codes


@pytest.fixture
def pq_mode():
    from longbow_b200 import _lib
    yield lambda m: _lib.set_option("pq_scan", m)
    _lib.set_option("pq_scan", 0)


@pytest.mark.parametrize("mode", [2, 3, 0])
@pytest.mark.parametrize("M,sub,n,nq", [(96, 8, 70001, 9), (16, 4, 20000, 5), (40, 2, 33333, 4), (64, 4, 9000, 1),
                                        (7, 3, 5000, 3), (33, 1, 4100, 6)])
def test_coarse_scan_bit_exact(oracle, pq_mode, mode, M, sub, n, nq):
    from longbow_b200 import gpu, pq
    rng = np.random.default_rng(7000 + M + mode)
    cb, codes = _setup(rng, n, M, sub)
    dim = M * sub
    q = rng.standard_normal((nq, dim)).astype(np.float32)
    enc = pq.PQEncoder(dim, M, 256, cb)
    # appended in uneven pieces: the tiled mirror must place rows at any offset (not only multiples of 32)
    cuts = [0, 17, 1000, n // 2 + 5, n]
    for a, b in zip(cuts[:-1], cuts[1:]):
        enc.add_codes(codes[a:b])
    pq_mode(mode)
    gd, gl = enc.search(q, 10)
    wd, wl = oracle.pq_search(cb, codes, None, q, 10, 0)
    assert_topk_equal(gd, gl, wd, wl, 0.0, f"adc top-10 mode {mode}")
    gd, gl = enc.search(q, 100)
    wd, wl = oracle.pq_search(cb, codes, None, q, 100, 0)
    assert_topk_equal(gd, gl, wd, wl, 0.0, f"adc top-100 mode {mode}")
    # bitmaps + fp32 re-rank of k' = 100
    raw = oracle.pq_decode(codes, cb) + rng.normal(0, 0.05, (n, dim)).astype(np.float32)
    rawidx = gpu.DenseIndex(dim, np.float32, L2)
    rawidx.add(raw)
    enc.attach_raw(rawidx)
    tomb, allow = random_bitmap(rng, n, 0.05), random_bitmap(rng, n, 0.3)
    enc.set_tombstones(tomb)
    gd, gl = enc.search(q, 10, 100, allow=allow)
    wd, wl = oracle.pq_search(cb, codes, raw, q, 10, 100, tomb=gpu.pack_bitmap(tomb), allow=gpu.pack_bitmap(allow))
    assert_topk_equal(gd, gl, wd, wl, 0.0, f"adc + rerank + bitmaps mode {mode}")
    enc.close(); rawidx.close()


@pytest.mark.parametrize("mode", [2, 3])
def test_coarse_scan_ties_by_id(oracle, pq_mode, mode):
    """Duplicate code rows: equal ADC distances, the lower id must win, also across CTA parts."""
    from longbow_b200 import pq
    rng = np.random.default_rng(42)
    M, sub, n = 32, 4, 50000
    cb, codes = _setup(rng, n, M, sub)
    codes[40000:40040] = codes[123]      # 41 identical rows spread over the scan
    codes[777] = codes[123]
    q = (oracle.pq_decode(codes[123:124], cb)[0] + 0.01).astype(np.float32)[None, :]
    enc = pq.PQEncoder(M * sub, M, 256, cb)
    enc.add_codes(codes)
    pq_mode(mode)
    gd, gl = enc.search(np.repeat(q, 5, axis=0), 20)
    wd, wl = oracle.pq_search(cb, codes, None, np.repeat(q, 5, axis=0), 20, 0)
    assert_topk_equal(gd, gl, wd, wl, 0.0, "duplicate rows")
    enc.close()


@pytest.mark.parametrize("mode", [2, 3])
def test_coarse_scan_certification_repairs(oracle, pq_mode, mode):
    """A codebook whose centroids differ by ~1e-7 relative: thousands of rows fall inside one quantisation step of
    the k-th distance, far more than the candidate margin.  The scan must flag the query and the host call must
    return the exhaustive answer."""
    from longbow_b200 import pq
    rng = np.random.default_rng(5)
    M, sub, n = 16, 4, 30000
    base = rng.standard_normal((M, 1, sub)).astype(np.float32)
    cb = np.repeat(base, 256, axis=1)
    cb += (rng.standard_normal(cb.shape) * 1e-6).astype(np.float32)   # all centroids of a sub-quantiser nearly equal
    cb[:, 0] += 5.0                                                      # one far centroid sets the LUT range
    codes = rng.integers(1, 256, (n, M), dtype=np.uint8)
    q = rng.standard_normal((3, M * sub)).astype(np.float32)
    enc = pq.PQEncoder(M * sub, M, 256, cb)
    enc.add_codes(codes)
    pq_mode(mode)
    gd, gl = enc.search(q, 10)
    wd, wl = oracle.pq_search(cb, codes, None, q, 10, 0)
    assert_topk_equal(gd, gl, wd, wl, 0.0, "near-equal sums")
    assert enc.last_uncertified() >= 1
    enc.close()


def test_coarse_scan_ring_kernel(oracle, pq_mode):
    """The warp-specialised bulk-copy ring variant of the single-query pass (off by default): same answers."""
    from longbow_b200 import _lib, pq
    rng = np.random.default_rng(77)
    M, sub, n = 96, 8, 90001
    cb, codes = _setup(rng, n, M, sub)
    q = rng.standard_normal((3, M * sub)).astype(np.float32)
    enc = pq.PQEncoder(M * sub, M, 256, cb)
    enc.add_codes(codes)
    pq_mode(2)
    _lib.set_option("pq_ring", 1)
    try:
        gd, gl = enc.search(q, 100)
    finally:
        _lib.set_option("pq_ring", 0)
    wd, wl = oracle.pq_search(cb, codes, None, q, 100, 0)
    assert_topk_equal(gd, gl, wd, wl, 0.0, "ring kernel")
    assert enc.last_uncertified() == 0
    enc.close()


def test_coarse_scan_small_and_empty(oracle, pq_mode):
    from longbow_b200 import pq
    rng = np.random.default_rng(9)
    M, sub = 8, 4
    cb, codes = _setup(rng, 100, M, sub)
    q = rng.standard_normal((2, M * sub)).astype(np.float32)
    enc = pq.PQEncoder(M * sub, M, 256, cb)
    d, l = enc.search(q, 5)
    assert (l == -1).all()
    enc.add_codes(codes)          # below the coarse path's minimum size: exhaustive kernel
    gd, gl = enc.search(q, 200)   # k > n: padding
    wd, wl = oracle.pq_search(cb, codes, None, q, 200, 0)
    assert_topk_equal(gd, gl, wd, wl, 0.0, "tiny")
    enc.close()


@pytest.mark.parametrize("M,sub,n,nq", [(96, 8, 70001, 130), (16, 4, 20000, 70), (64, 4, 9000, 64), (32, 16, 30000, 96),
                                        (40, 2, 33333, 65)])
def test_gemm_coarse_bit_exact(oracle, pq_mode, M, sub, n, nq):
    """Batches (>= 64 queries): coarse stage = decode to fp16 + dense tensor-core scan (csrc/pq_gemm.cu), then the same
    exact fp32 table sums + certification.  Forced (mode 4) and by the automatic policy; bitmaps + re-rank too."""
    from longbow_b200 import gpu, pq
    rng = np.random.default_rng(9100 + M)
    cb, codes = _setup(rng, n, M, sub)
    dim = M * sub
    q = rng.standard_normal((nq, dim)).astype(np.float32)
    q[:5] = oracle.pq_decode(codes[[0, 5, 4000, n - 1, n // 2]], cb)   # queries that are (decoded) rows: distance 0
    enc = pq.PQEncoder(dim, M, 256, cb)
    cuts = [0, 17, 1000, n // 2 + 5, n]
    for a, b in zip(cuts[:-1], cuts[1:]):
        enc.add_codes(codes[a:b])
    for mode in (4, 0):
        pq_mode(mode)
        gd, gl = enc.search(q, 10)
        wd, wl = oracle.pq_search(cb, codes, None, q, 10, 0)
        assert_topk_equal(gd, gl, wd, wl, 0.0, f"gemm adc top-10 mode {mode}")
    pq_mode(4)
    gd, gl = enc.search(q, 100)
    wd, wl = oracle.pq_search(cb, codes, None, q, 100, 0)
    assert_topk_equal(gd, gl, wd, wl, 0.0, "gemm adc top-100")
    raw = oracle.pq_decode(codes, cb) + rng.normal(0, 0.05, (n, dim)).astype(np.float32)
    rawidx = gpu.DenseIndex(dim, np.float32, L2)
    rawidx.add(raw)
    enc.attach_raw(rawidx)
    tomb, allow = random_bitmap(rng, n, 0.05), random_bitmap(rng, n, 0.3)
    enc.set_tombstones(tomb)
    gd, gl = enc.search(q, 10, 100, allow=allow)
    wd, wl = oracle.pq_search(cb, codes, raw, q, 10, 100, tomb=gpu.pack_bitmap(tomb), allow=gpu.pack_bitmap(allow))
    assert_topk_equal(gd, gl, wd, wl, 0.0, "gemm adc + rerank + bitmaps")
    enc.close(); rawidx.close()


def test_gemm_coarse_ties_and_certification(oracle, pq_mode):
    """Duplicate rows (equal distances: lower id wins) and a codebook whose centroids differ by ~1e-6: fp16 decoding
    cannot separate them, the query must be flagged and the host call must return the exhaustive answer."""
    from longbow_b200 import pq
    rng = np.random.default_rng(43)
    M, sub, n = 32, 4, 50000
    cb, codes = _setup(rng, n, M, sub)
    codes[40000:40040] = codes[123]
    codes[777] = codes[123]
    q = rng.standard_normal((70, M * sub)).astype(np.float32)
    q[0] = oracle.pq_decode(codes[123:124], cb)[0] + 0.01
    enc = pq.PQEncoder(M * sub, M, 256, cb)
    enc.add_codes(codes)
    pq_mode(4)
    gd, gl = enc.search(q, 20)
    wd, wl = oracle.pq_search(cb, codes, None, q, 20, 0)
    assert_topk_equal(gd, gl, wd, wl, 0.0, "gemm duplicate rows")
    enc.close()
    M, sub, n = 16, 4, 30000
    base = rng.standard_normal((M, 1, sub)).astype(np.float32)
    cb = np.repeat(base, 256, axis=1)
    cb += (rng.standard_normal(cb.shape) * 1e-6).astype(np.float32)
    cb[:, 0] += 5.0
    codes = rng.integers(1, 256, (n, M), dtype=np.uint8)
    q = rng.standard_normal((66, M * sub)).astype(np.float32)
    enc = pq.PQEncoder(M * sub, M, 256, cb)
    enc.add_codes(codes)
    gd, gl = enc.search(q, 10)
    wd, wl = oracle.pq_search(cb, codes, None, q, 10, 0)
    assert_topk_equal(gd, gl, wd, wl, 0.0, "gemm near-equal sums")
    assert enc.last_uncertified() >= 1
    enc.close()


def test_gemm_path_declines_out_of_range_codebooks(oracle, pq_mode):
    """Centroid components beyond the fp16 comfort zone: no fp16 codebook is built, a 70-query batch takes the look-up
    passes (forcing mode 4 is refused) and the answers are the oracle's."""
    from longbow_b200 import _lib, pq
    rng = np.random.default_rng(11)
    M, sub, n = 16, 4, 20000
    cb, codes = _setup(rng, n, M, sub)
    cb = (cb * 3.0e4).astype(np.float32)
    q = (rng.standard_normal((70, M * sub)) * 3.0e4).astype(np.float32)
    enc = pq.PQEncoder(M * sub, M, 256, cb)
    enc.add_codes(codes)
    gd, gl = enc.search(q, 10)
    wd, wl = oracle.pq_search(cb, codes, None, q, 10, 0)
    assert_topk_equal(gd, gl, wd, wl, 0.0, "large centroids")
    pq_mode(4)
    with pytest.raises(_lib.LongbowError):
        enc.search(q, 10)
    enc.close()


def test_gemm_path_can_be_switched_off(oracle):
    """lb_set_option("pq_gemm", 0): batches stay on the look-up passes (no decode scratch); same answers."""
    from longbow_b200 import _lib, pq
    rng = np.random.default_rng(12)
    M, sub, n = 16, 4, 20000
    cb, codes = _setup(rng, n, M, sub)
    q = rng.standard_normal((70, M * sub)).astype(np.float32)
    enc = pq.PQEncoder(M * sub, M, 256, cb)
    enc.add_codes(codes)
    wd, wl = oracle.pq_search(cb, codes, None, q, 10, 0)
    _lib.set_option("pq_gemm", 0)
    try:
        l0 = _lib.launch_count()
        gd, gl = enc.search(q, 10)
        n_off = _lib.launch_count() - l0
    finally:
        _lib.set_option("pq_gemm", 1)
    assert_topk_equal(gd, gl, wd, wl, 0.0, "pq_gemm off")
    l0 = _lib.launch_count()
    gd, gl = enc.search(q, 10)
    n_on = _lib.launch_count() - l0
    assert_topk_equal(gd, gl, wd, wl, 0.0, "pq_gemm on")
    assert n_on != n_off   # different kernel chains actually ran
    enc.close()
