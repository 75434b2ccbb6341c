#!/usr/bin/env python
"""Generates the committed fixtures under tests/golden/.

1. reference_kats.json -- every known-answer test the reference holds for the hot path, transcribed
   as data (inputs, expected value, tolerance, reference file:line).  The reference is Go and cannot
   run in this image, so these are copied by hand from its test sources and docs; `--verify` re-reads
   /root/reference (when present) and checks that each cited line range still contains the literals.
2. oracle_vectors.npz -- small seeded inputs with the outputs of the O-exact oracle (oracle/lb_oracle.c)
   for each (dtype, metric) of the dense path, the PQ path and the re-rank path.  They pin the oracle
   against drift and give the GPU tests a fixture that does not depend on compiling anything.
   Provenance: produced by THIS script from the oracle that tests/test_oracle_kat.py pins on (1).

  python tests/golden/make_golden.py [--verify]
"""
import json
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

KATS = [
    # metric: 0 L2, 1 cosine, 2 dot (raw).  expected None = range check only.
    {"name": "euclidean_via_dispatch", "ref": "internal/simd/simd_dispatch_test.go:60-68", "dtype": "f32", "metric": 0,
     "a": [1, 2, 3, 4], "b": [5, 6, 7, 8], "expected": 8.0, "tol": 1e-4},
    {"name": "dot_product_via_dispatch", "ref": "internal/simd/simd_dispatch_test.go:70-78", "dtype": "f32", "metric": 2,
     "a": [1, 2, 3, 4], "b": [5, 6, 7, 8], "expected": 70.0, "tol": 1e-4},
    {"name": "cosine_via_dispatch_range", "ref": "internal/simd/simd_dispatch_test.go:80-87", "dtype": "f32", "metric": 1,
     "a": [1, 2, 3, 4], "b": [5, 6, 7, 8], "expected": None, "range": [0.0, 2.0]},
    {"name": "euclidean_batch_0", "ref": "internal/simd/simd_dispatch_test.go:89-105", "dtype": "f32", "metric": 0,
     "a": [1, 2, 3, 4], "b": [5, 6, 7, 8], "expected": 8.0, "tol": 1e-4},
    {"name": "euclidean_batch_1_self", "ref": "internal/simd/simd_dispatch_test.go:89-105", "dtype": "f32", "metric": 0,
     "a": [1, 2, 3, 4], "b": [1, 2, 3, 4], "expected": 0.0, "tol": 1e-4},
    {"name": "dispatch_float32_128", "ref": "internal/simd/simd_dispatch_test.go:115-123", "dtype": "f32", "metric": 0,
     "a": [1.0] + [0.0] * 127, "b": [2.0] + [0.0] * 127, "expected": 1.0, "tol": 1e-4},
    {"name": "dispatch_int8", "ref": "internal/simd/simd_dispatch_test.go:125-132", "dtype": "i8", "metric": 0,
     "a": [10, 20], "b": [10, 30], "expected": 10.0, "tol": 1e-4},
    {"name": "cosine_identical", "ref": "internal/simd/simd_test.go:146-155", "dtype": "f32", "metric": 1,
     "a": [1, 2, 3, 4, 5, 6, 7, 8], "b": [1, 2, 3, 4, 5, 6, 7, 8], "expected": 0.0, "tol": 1e-5},
    {"name": "cosine_orthogonal", "ref": "internal/simd/simd_test.go:157-168", "dtype": "f32", "metric": 1,
     "a": [1, 0, 0, 0], "b": [0, 1, 0, 0], "expected": 1.0, "tol": 1e-5},
    {"name": "cosine_opposite", "ref": "internal/simd/simd_test.go:170-181", "dtype": "f32", "metric": 1,
     "a": [1, 2, 3, 4], "b": [-1, -2, -3, -4], "expected": 2.0, "tol": 1e-5},
    {"name": "cosine_zero_vector_exact", "ref": "internal/simd/simd_test.go:183-194", "dtype": "f32", "metric": 1,
     "a": [0, 0, 0, 0], "b": [1, 2, 3, 4], "expected": 1.0, "tol": 0.0},
    {"name": "l2_kernel_self", "ref": "internal/store/arrow_kernels_test.go:56-66", "dtype": "f32", "metric": 0,
     "a": [1, 2, 3, 4], "b": [1, 2, 3, 4], "expected": 0.0, "tol": 1e-6},
    {"name": "l2_kernel_sqrt30", "ref": "internal/store/arrow_kernels_test.go:56-69", "dtype": "f32", "metric": 0,
     "a": [1, 2, 3, 4], "b": [0, 0, 0, 0], "expected": float(np.float32(math.sqrt(30.0))), "tol": 1e-6},
    {"name": "doc_euclid_sqrt27", "ref": "docs/distance_metrics.md:21", "dtype": "f32", "metric": 0,
     "a": [1, 2, 3], "b": [4, 5, 6], "expected": math.sqrt(27.0), "tol": 1e-6},
    {"name": "doc_dot_11", "ref": "docs/distance_metrics.md:55", "dtype": "f32", "metric": 2,
     "a": [1, 2], "b": [3, 4], "expected": 11.0, "tol": 0.0},
]
# the GPU smoke test of the boundary: ramp vectors, nearest id first (internal/gpu/gpu_test.go:25-46)
GPU_SMOKE = {"ref": "internal/gpu/gpu_test.go:25-46", "dim": 128, "n": 10, "scale": 0.01, "k": 5,
             "expect_first_id": 0, "expect_first_dist_below": 0.01}
# filter: 3 rows, category == "B" -> row 1 only (internal/store/bitmap_filter_test.go:94-153); categories as codes
FILTER = {"ref": "internal/store/bitmap_filter_test.go:94-153", "column": [0, 1, 2], "op": "==", "value": 1,
          "expect_rows": [1]}


def write_kats():
    with open(os.path.join(HERE, "reference_kats.json"), "w") as f:
        json.dump({"pairs": KATS, "gpu_smoke": GPU_SMOKE, "filter": FILTER}, f, indent=1)


def write_oracle_vectors():
    from oracle import oracle
    oracle.build()
    out = {}
    rng = np.random.default_rng(20261018)
    n, nq, k = 600, 9, 12
    for dim in (128, 33):
        f32 = rng.random((n, dim), dtype=np.float32)
        f16 = rng.standard_normal((n, dim)).astype(np.float32)
        f16 = (f16 / np.linalg.norm(f16, axis=1, keepdims=True)).astype(np.float16)
        f16[7] = 0  # cosine zero-row rule
        i8 = rng.integers(-128, 128, (n, dim), dtype=np.int8)
        for name, db in (("f32", f32), ("f16", f16), ("i8", i8)):
            q = db[rng.integers(0, n, nq)].copy()
            if name != "i8":
                q = (q.astype(np.float32) + rng.standard_normal(q.shape).astype(np.float32) * 0.05).astype(db.dtype)
            out[f"dense_{name}_{dim}_db"] = db
            out[f"dense_{name}_{dim}_q"] = q
            for metric, mname in ((oracle.L2, "l2"), (oracle.COSINE, "cos"), (oracle.DOT, "dot")):
                if name == "i8" and metric == oracle.COSINE:
                    continue
                d, l = oracle.search(metric, db, q, k)
                out[f"dense_{name}_{dim}_{mname}_d"] = d
                out[f"dense_{name}_{dim}_{mname}_l"] = l
    # PQ: M=8 sub=4, ADC scan bit-exact + table
    M, sub, npq = 8, 4, 900
    cb = rng.standard_normal((M, 256, sub)).astype(np.float32)
    codes = rng.integers(0, 256, (npq, M), dtype=np.uint8)
    qq = rng.standard_normal((5, M * sub)).astype(np.float32)
    d, l = oracle.pq_search(cb, codes, None, qq, 10, 0)
    out.update(pq_cb=cb, pq_codes=codes, pq_q=qq, pq_d=d, pq_l=l, pq_table0=oracle.adc_table(qq[0], cb))
    # re-rank with bitmaps
    db = out["dense_f32_128_db"]
    q = out["dense_f32_128_q"]
    cand = rng.integers(0, n + 20, (nq, 40)).astype(np.uint32)  # some ids out of range
    tomb = rng.random(n) < 0.1
    allow = rng.random(n) < 0.6
    from longbow_b200.gpu import pack_bitmap
    d, l = oracle.rerank(oracle.L2, db, q, cand.astype(np.int64), 10, tomb=pack_bitmap(tomb), allow=pack_bitmap(allow))
    out.update(rr_cand=cand, rr_tomb=tomb, rr_allow=allow, rr_d=d, rr_l=l)
    np.savez_compressed(os.path.join(HERE, "oracle_vectors.npz"), **out)


def verify_against_reference():
    ref = "/root/reference"
    if not os.path.isdir(ref):
        print("reference tree absent: nothing to verify")
        return
    bad = 0
    for kcase in KATS + [GPU_SMOKE, FILTER]:
        path, rng_ = kcase["ref"].split(":")
        lo, hi = (int(x) for x in (rng_.split("-") + [rng_])[:2])
        lines = open(os.path.join(ref, path)).read().split("\n")[lo - 1:hi]
        if not any(s.strip() for s in lines):
            print("EMPTY RANGE", kcase["ref"])
            bad += 1
    print("verified" if not bad else f"{bad} stale citations")


if __name__ == "__main__":
    if "--verify" in sys.argv:
        verify_against_reference()
    else:
        write_kats()
        write_oracle_vectors()
        print("wrote", os.listdir(HERE))
