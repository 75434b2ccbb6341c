"""Shared helpers for the parity tests: seeded synthetic inputs and comparison rules
(north star): int8 / PQ bit-exact ids + distances; fp32 same ids, distances <= 1e-5 relative; fp16
<= 1e-2 relative; ties broken by id.  The CUDA path re-scores candidates with the reference's own
arithmetic, so in practice it is bit-exact for every dtype; the tests assert exactness and report the
looser contractual bound in the failure message."""
import numpy as np


def make_db(rng, n, dim, dtype):
    if dtype == np.float32:
        return rng.random((n, dim), dtype=np.float32)
    if dtype == np.float16:
        x = rng.standard_normal((n, dim)).astype(np.float32)
        x /= np.maximum(np.linalg.norm(x, axis=1, keepdims=True), 1e-12)
        return x.astype(np.float16)
    if dtype == np.int8:
        return rng.integers(-128, 128, (n, dim), dtype=np.int8)
    raise ValueError(dtype)


def random_bitmap(rng, n, frac):
    return rng.random(n) < frac


def assert_topk_equal(got_d, got_l, want_d, want_l, rtol, what=""):
    assert got_l.shape == want_l.shape and got_d.shape == want_d.shape
    if not np.array_equal(got_l, want_l):
        bad = np.argwhere(got_l != want_l)
        q, j = bad[0]
        raise AssertionError(f"{what}: labels differ at query {q} rank {j}: got {got_l[q]} / {got_d[q]} "
                             f"want {want_l[q]} / {want_d[q]} ({len(bad)} mismatches)")
    valid = want_l >= 0
    denom = np.maximum(np.abs(want_d[valid]), 1e-30)
    rel = np.abs(got_d[valid] - want_d[valid]) / denom
    assert rel.size == 0 or rel.max() <= rtol, f"{what}: distance rel err {rel.max()} > {rtol}"
    assert np.array_equal(got_d[~valid], want_d[~valid]), f"{what}: padding distances differ"
