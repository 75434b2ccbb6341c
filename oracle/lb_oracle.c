/*
 * lb_oracle.c -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 *
 * CPU restatement ("O-exact") of the arithmetic on Longbow's vector-distance /
 * k-NN hot path.  The reference is Go and cannot be built in this image (no Go
 * toolchain), so every function below restates the *portable Go* that the
 * reference falls back to on any CPU (the "default" branch of
 * internal/simd/dispatch.go:184-213 wires the Unrolled4x kernels), citing the
 * file:line it follows.  Lane order, operation order and rounding points are the
 * reference's: fp32 multiply and add are separate roundings (Go does not fuse
 * on amd64), so this file MUST be compiled with -ffp-contract=off and without
 * -ffast-math (see oracle/Makefile).
 *
 * Pinning: the reference ships no golden vectors for this path;
 * the oracle is pinned against the reference's own known-answer tests
 * (internal/simd/simd_dispatch_test.go:56-142, internal/simd/simd_test.go:146-194,
 * internal/store/arrow_kernels_test.go:56-69, internal/gpu/gpu_test.go:25-46,
 * docs/distance_metrics.md:21,55) in tests/test_oracle_kat.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define LBO_API __attribute__((visibility("default")))

/* metric / dtype enums: internal/simd/registry.go:8-14, 31-47 */
enum { LBO_L2 = 0, LBO_COSINE = 1, LBO_DOT = 2 };
enum { LBO_F32 = 0, LBO_F16 = 1, LBO_I8 = 2, LBO_U8 = 3 };

/* ---------------------------------------------------------------------------
 * fp16 -> fp32 widening.  arrow-go v18.5.1 arrow/float16 Num.Float32() (go.mod:6)
 * is an exact IEEE-754 binary16 -> binary32 conversion (subnormals, inf, nan
 * preserved); restated bit-wise so it does not depend on F16C being present.
 * ------------------------------------------------------------------------- */
static inline float h2f(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1fu;
    uint32_t man = h & 0x3ffu;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else { /* subnormal: normalise */
            int e = -1;
            do { man <<= 1; e++; } while (!(man & 0x400u));
            man &= 0x3ffu;
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | (man << 13);
        }
    } else if (exp == 31) {
        bits = sign | 0x7f800000u | (man << 13);
    } else {
        bits = sign | ((exp + 127 - 15) << 23) | (man << 13);
    }
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

LBO_API float lbo_h2f(uint16_t h) { return h2f(h); }

/* Go: float32(math.Sqrt(float64(x))).  sqrt in double then narrowed equals the
 * correctly rounded single-precision sqrt (double has > 2*24+2 bits). */
static inline float sqrt_via_f64(float x) { return (float)sqrt((double)x); }

/* ---------------------------------------------------------------------------
 * fp32 kernels
 * ------------------------------------------------------------------------- */

/* internal/simd/distance_functions.go:195-227 (L2SquaredFloat32) and
 * internal/simd/simd.go:365-396 (euclideanUnrolled4x): four accumulators,
 * lane l sums elements i == l (mod 4) in increasing i, remainder into lane 0,
 * final (s0+s1)+s2)+s3. */
LBO_API float lbo_l2sq_f32(const float *a, const float *b, int n) {
    float s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    int i = 0;
    for (; i <= n - 4; i += 4) {
        float d0 = a[i] - b[i], d1 = a[i + 1] - b[i + 1];
        float d2 = a[i + 2] - b[i + 2], d3 = a[i + 3] - b[i + 3];
        s0 += d0 * d0; s1 += d1 * d1; s2 += d2 * d2; s3 += d3 * d3;
    }
    for (; i < n; i++) { float d = a[i] - b[i]; s0 += d * d; }
    return s0 + s1 + s2 + s3;
}

/* internal/simd/distance_functions.go:17-31 -> simd.go:131-134 / :365-396.
 * Empty vectors -> 0 (distance_functions.go:21-23). */
LBO_API float lbo_euclid_f32(const float *a, const float *b, int n) {
    if (n == 0) return 0.0f;
    return sqrt_via_f64(lbo_l2sq_f32(a, b, n));
}

/* internal/simd/simd.go:399-450 (cosineUnrolled4x).  Exactly 1.0 when either
 * squared norm is 0 (:446-448); empty -> 1.0 (distance_functions.go:51-53). */
LBO_API float lbo_cosine_f32(const float *a, const float *b, int n) {
    if (n == 0) return 1.0f;
    float d0 = 0, d1 = 0, d2 = 0, d3 = 0;
    float a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    float b0 = 0, b1 = 0, b2 = 0, b3 = 0;
    int i = 0;
    for (; i <= n - 4; i += 4) {
        float x0 = a[i], x1 = a[i + 1], x2 = a[i + 2], x3 = a[i + 3];
        float y0 = b[i], y1 = b[i + 1], y2 = b[i + 2], y3 = b[i + 3];
        d0 += x0 * y0; d1 += x1 * y1; d2 += x2 * y2; d3 += x3 * y3;
        a0 += x0 * x0; a1 += x1 * x1; a2 += x2 * x2; a3 += x3 * x3;
        b0 += y0 * y0; b1 += y1 * y1; b2 += y2 * y2; b3 += y3 * y3;
    }
    for (; i < n; i++) { d0 += a[i] * b[i]; a0 += a[i] * a[i]; b0 += b[i] * b[i]; }
    float dot = d0 + d1 + d2 + d3;
    float na = a0 + a1 + a2 + a3;
    float nb = b0 + b1 + b2 + b3;
    if (na == 0 || nb == 0) return 1.0f;
    return 1.0f - (dot / (float)sqrt((double)na * (double)nb));
}

/* internal/simd/simd.go:453-479 (dotUnrolled4x); empty -> 0. */
LBO_API float lbo_dot_f32(const float *a, const float *b, int n) {
    float s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    int i = 0;
    for (; i <= n - 4; i += 4) {
        s0 += a[i] * b[i]; s1 += a[i + 1] * b[i + 1];
        s2 += a[i + 2] * b[i + 2]; s3 += a[i + 3] * b[i + 3];
    }
    for (; i < n; i++) s0 += a[i] * b[i];
    return s0 + s1 + s2 + s3;
}

/* Single-accumulator "generic" variants, internal/simd/simd.go:138-163
 * (cosineGeneric, dotGeneric) -- kept so tests can bound the spread between
 * the reference's own CPU variants (they differ only in summation order). */
LBO_API float lbo_cosine_f32_seq(const float *a, const float *b, int n) {
    float dot = 0, na = 0, nb = 0;
    for (int i = 0; i < n; i++) { dot += a[i] * b[i]; na += a[i] * a[i]; nb += b[i] * b[i]; }
    if (na == 0 || nb == 0) return 1.0f;
    return 1.0f - (dot / (float)sqrt((double)na * (double)nb));
}
LBO_API float lbo_dot_f32_seq(const float *a, const float *b, int n) {
    float s = 0;
    for (int i = 0; i < n; i++) s += a[i] * b[i];
    return s;
}

/* ---------------------------------------------------------------------------
 * fp16 kernels: internal/simd/simd.go:767-848.  Each element widened exactly,
 * fp32 arithmetic, same 4-lane order.
 * ------------------------------------------------------------------------- */
LBO_API float lbo_euclid_f16(const uint16_t *a, const uint16_t *b, int n) {
    if (n == 0) return 0.0f; /* distance_functions.go:80-82 */
    float s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    int i = 0;
    for (; i <= n - 4; i += 4) {
        float d0 = h2f(a[i]) - h2f(b[i]), d1 = h2f(a[i + 1]) - h2f(b[i + 1]);
        float d2 = h2f(a[i + 2]) - h2f(b[i + 2]), d3 = h2f(a[i + 3]) - h2f(b[i + 3]);
        s0 += d0 * d0; s1 += d1 * d1; s2 += d2 * d2; s3 += d3 * d3;
    }
    for (; i < n; i++) { float d = h2f(a[i]) - h2f(b[i]); s0 += d * d; }
    return sqrt_via_f64(s0 + s1 + s2 + s3);
}

LBO_API float lbo_dot_f16(const uint16_t *a, const uint16_t *b, int n) {
    float s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    int i = 0;
    for (; i <= n - 4; i += 4) {
        s0 += h2f(a[i]) * h2f(b[i]); s1 += h2f(a[i + 1]) * h2f(b[i + 1]);
        s2 += h2f(a[i + 2]) * h2f(b[i + 2]); s3 += h2f(a[i + 3]) * h2f(b[i + 3]);
    }
    for (; i < n; i++) s0 += h2f(a[i]) * h2f(b[i]);
    return s0 + s1 + s2 + s3;
}

LBO_API float lbo_cosine_f16(const uint16_t *a, const uint16_t *b, int n) {
    if (n == 0) return 1.0f; /* distance_functions.go:92-94 */
    float d0 = 0, d1 = 0, d2 = 0, d3 = 0;
    float a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    float b0 = 0, b1 = 0, b2 = 0, b3 = 0;
    int i = 0;
    for (; i <= n - 4; i += 4) {
        float x0 = h2f(a[i]), x1 = h2f(a[i + 1]), x2 = h2f(a[i + 2]), x3 = h2f(a[i + 3]);
        float y0 = h2f(b[i]), y1 = h2f(b[i + 1]), y2 = h2f(b[i + 2]), y3 = h2f(b[i + 3]);
        d0 += x0 * y0; d1 += x1 * y1; d2 += x2 * y2; d3 += x3 * y3;
        a0 += x0 * x0; a1 += x1 * x1; a2 += x2 * x2; a3 += x3 * x3;
        b0 += y0 * y0; b1 += y1 * y1; b2 += y2 * y2; b3 += y3 * y3;
    }
    for (; i < n; i++) {
        float x = h2f(a[i]), y = h2f(b[i]);
        d0 += x * y; a0 += x * x; b0 += y * y;
    }
    float dot = d0 + d1 + d2 + d3;
    float na = a0 + a1 + a2 + a3;
    float nb = b0 + b1 + b2 + b3;
    if (na == 0 || nb == 0) return 1.0f;
    return 1.0f - (dot / (float)sqrt((double)na * (double)nb));
}

/* ---------------------------------------------------------------------------
 * int8 kernels: internal/simd/simd_baseline.go:13-54.  Elements converted to
 * fp32, fp32 accumulate in 4 lanes.  (No cosine kernel is registered for int8:
 * internal/simd/dispatch.go:241-242.)
 * ------------------------------------------------------------------------- */
LBO_API float lbo_euclid_i8(const int8_t *a, const int8_t *b, int n) {
    if (n == 0) return 0.0f; /* dispatch.go:268-270 */
    float s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    int i = 0;
    for (; i <= n - 4; i += 4) {
        float d0 = (float)a[i] - (float)b[i], d1 = (float)a[i + 1] - (float)b[i + 1];
        float d2 = (float)a[i + 2] - (float)b[i + 2], d3 = (float)a[i + 3] - (float)b[i + 3];
        s0 += d0 * d0; s1 += d1 * d1; s2 += d2 * d2; s3 += d3 * d3;
    }
    for (; i < n; i++) { float d = (float)a[i] - (float)b[i]; s0 += d * d; }
    return sqrt_via_f64(s0 + s1 + s2 + s3);
}

LBO_API float lbo_dot_i8(const int8_t *a, const int8_t *b, int n) {
    float s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    int i = 0;
    for (; i <= n - 4; i += 4) {
        s0 += (float)a[i] * (float)b[i]; s1 += (float)a[i + 1] * (float)b[i + 1];
        s2 += (float)a[i + 2] * (float)b[i + 2]; s3 += (float)a[i + 3] * (float)b[i + 3];
    }
    for (; i < n; i++) s0 += (float)a[i] * (float)b[i];
    return s0 + s1 + s2 + s3;
}

/* SQ8: internal/simd/sq8.go:45-66 -- squared L2 over uint8 as int32 (no sqrt);
 * the batch form returns float32(int32) (simd.go:166-178). */
LBO_API int32_t lbo_l2sq_u8(const uint8_t *a, const uint8_t *b, int n) {
    int32_t sum = 0;
    for (int i = 0; i < n; i++) { int32_t d = (int32_t)a[i] - (int32_t)b[i]; sum += d * d; }
    return sum;
}

/* ---------------------------------------------------------------------------
 * Metric as a *distance to minimise*.  Dot product is negated
 * (internal/store/distance_resolvers.go:11-16,74-84 (commented intent),
 * docs/distance_metrics.md:42-55).
 * ------------------------------------------------------------------------- */
static float pair_distance(int metric, int dtype, const void *a, const void *b, int n) {
    switch (dtype) {
    case LBO_F32:
        if (metric == LBO_L2) return lbo_euclid_f32(a, b, n);
        if (metric == LBO_COSINE) return lbo_cosine_f32(a, b, n);
        return -lbo_dot_f32(a, b, n);
    case LBO_F16:
        if (metric == LBO_L2) return lbo_euclid_f16(a, b, n);
        if (metric == LBO_COSINE) return lbo_cosine_f16(a, b, n);
        return -lbo_dot_f16(a, b, n);
    case LBO_I8:
        if (metric == LBO_L2) return lbo_euclid_i8(a, b, n);
        if (metric == LBO_DOT) return -lbo_dot_i8(a, b, n);
        return NAN;
    case LBO_U8:
        if (metric == LBO_L2) return (float)lbo_l2sq_u8(a, b, n);
        return NAN;
    }
    return NAN;
}

static size_t elem_size(int dtype) {
    return dtype == LBO_F32 ? 4 : dtype == LBO_F16 ? 2 : 1;
}

LBO_API float lbo_distance(int metric, int dtype, const void *a, const void *b, int n) {
    return pair_distance(metric, dtype, a, b, n);
}

/* One query x n rows in a flat row-major buffer:
 * internal/simd/batch_operations.go:64-87, simd.go:203-229. */
LBO_API void lbo_batch_flat(int metric, int dtype, const void *q, const void *flat,
                            int64_t n, int dim, float *out) {
    size_t stride = (size_t)dim * elem_size(dtype);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++)
        out[i] = pair_distance(metric, dtype, q, (const char *)flat + (size_t)i * stride, dim);
}

/* ---------------------------------------------------------------------------
 * Top-k with the reference's brute-force rule, made deterministic:
 * internal/store/adaptive_index.go:176-222 keeps a size-k max-heap and replaces
 * only on dist < top (strict), scanning ids in increasing order, so among
 * equal distances the lowest ids survive; output ascending by distance.  The
 * order *inside* ties is unpinned in the reference (heap pop / sort.Slice);
 * we fix it to (distance, id) ascending.
 * ------------------------------------------------------------------------- */
typedef struct { float d; int64_t id; } lbo_pair;

static inline int pair_less(lbo_pair a, lbo_pair b) {
    return a.d < b.d || (a.d == b.d && a.id < b.id);
}

/* bounded insertion into a sorted array of at most k (k is small) */
static inline void topk_push(lbo_pair *best, int *cnt, int k, lbo_pair p) {
    if (*cnt == k && !pair_less(p, best[k - 1])) return;
    int i = (*cnt < k) ? (*cnt)++ : k - 1;
    while (i > 0 && pair_less(p, best[i - 1])) { best[i] = best[i - 1]; i--; }
    best[i] = p;
}

static inline int bit_test(const uint64_t *bm, int64_t i) {
    return (int)((bm[i >> 6] >> (i & 63)) & 1u);
}

/* Brute-force k-NN over a flat DB for nq queries.
 * tomb: optional dense bitmap, bit set = row deleted (ArrowHNSW.deleted,
 *       internal/store/arrow_hnsw.go:147,468-472).
 * allow: optional dense bitmap, bit set = row passes the predicate
 *       (query.Bitset, internal/query/bitmap.go:13-120).
 * NaN / +Inf distances are never selected.  Slots without a result are padded
 * with label -1, distance FLT_MAX (FAISS convention). */
LBO_API int lbo_search(int metric, int dtype, const void *db, int64_t n, int dim,
                       const void *queries, int64_t nq, int k,
                       const uint64_t *tomb, const uint64_t *allow, int64_t id_base,
                       float *out_d, int64_t *out_id) {
    if (k <= 0 || dim <= 0) return 1;
    size_t stride = (size_t)dim * elem_size(dtype);
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t qi = 0; qi < nq; qi++) {
        lbo_pair *best = (lbo_pair *)malloc(sizeof(lbo_pair) * (size_t)k);
        int cnt = 0;
        const char *q = (const char *)queries + (size_t)qi * stride;
        for (int64_t i = 0; i < n; i++) {
            if (tomb && bit_test(tomb, i)) continue;
            if (allow && !bit_test(allow, i)) continue;
            float d = pair_distance(metric, dtype, q, (const char *)db + (size_t)i * stride, dim);
            if (!(d < INFINITY)) continue;
            lbo_pair p = { d, i };
            topk_push(best, &cnt, k, p);
        }
        for (int j = 0; j < k; j++) {
            out_d[qi * k + j] = j < cnt ? best[j].d : FLT_MAX;
            out_id[qi * k + j] = j < cnt ? best[j].id + id_base : -1;
        }
        free(best);
    }
    return 0;
}

/* Re-rank: internal/store/parallel_search.go:147-365 + :124-130, and
 * internal/store/hnsw_batch.go:206-245 (RerankBatch).  Per query, a list of c
 * candidate ids: drop ids failing the predicate bitmap (:183), drop deleted /
 * out-of-range ids (location miss, :186-229), Euclidean distance of the rest
 * (:347 -> EuclideanDistanceBatchFlat), sort ascending, first k.  Ids < 0 or
 * >= n are location misses.  Duplicate ids are NOT removed (RerankBatch does
 * not remove them either; searchLayer never produces them). */
LBO_API int lbo_rerank(int metric, int dtype, const void *db, int64_t n, int dim,
                       const void *queries, int64_t nq, const int64_t *cand, int c, int k,
                       const uint64_t *tomb, const uint64_t *allow,
                       float *out_d, int64_t *out_id) {
    if (k <= 0 || dim <= 0 || c < 0) return 1;
    size_t stride = (size_t)dim * elem_size(dtype);
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t qi = 0; qi < nq; qi++) {
        lbo_pair *best = (lbo_pair *)malloc(sizeof(lbo_pair) * (size_t)k);
        int cnt = 0;
        const char *q = (const char *)queries + (size_t)qi * stride;
        const int64_t *ids = cand + (size_t)qi * c;
        for (int j = 0; j < c; j++) {
            int64_t id = ids[j];
            if (id < 0 || id >= n) continue;
            if (allow && !bit_test(allow, id)) continue;
            if (tomb && bit_test(tomb, id)) continue;
            float d = pair_distance(metric, dtype, q, (const char *)db + (size_t)id * stride, dim);
            if (!(d < INFINITY)) continue;
            lbo_pair p = { d, id };
            topk_push(best, &cnt, k, p);
        }
        for (int j = 0; j < k; j++) {
            out_d[qi * k + j] = j < cnt ? best[j].d : FLT_MAX;
            out_id[qi * k + j] = j < cnt ? best[j].id : -1;
        }
        free(best);
    }
    return 0;
}

/* Shard merge: internal/store/sharded_hnsw.go:432-503 (concat, sort asc,
 * truncate) / internal/store/result_merger.go:34-100 (k-way merge).  Keyed on
 * (distance, id); label -1 entries are padding and ignored.
 * in_d/in_id: [parts][nq][k_in]; out: [nq][k]. */
LBO_API int lbo_merge(const float *in_d, const int64_t *in_id, int parts, int64_t nq, int k_in,
                      int k, float *out_d, int64_t *out_id) {
    if (k <= 0) return 1;
    for (int64_t qi = 0; qi < nq; qi++) {
        lbo_pair *best = (lbo_pair *)malloc(sizeof(lbo_pair) * (size_t)k);
        int cnt = 0;
        for (int p = 0; p < parts; p++)
            for (int j = 0; j < k_in; j++) {
                size_t o = ((size_t)p * nq + qi) * k_in + j;
                if (in_id[o] < 0) continue;
                lbo_pair pr = { in_d[o], in_id[o] };
                topk_push(best, &cnt, k, pr);
            }
        for (int j = 0; j < k; j++) {
            out_d[qi * k + j] = j < cnt ? best[j].d : FLT_MAX;
            out_id[qi * k + j] = j < cnt ? best[j].id : -1;
        }
        free(best);
    }
    return 0;
}

/* select_k_neighbors operator: internal/store/arrow_kernels.go:230-345 sorts
 * all n (distance, index) pairs ascending and returns the first k indices.
 * Tie order fixed to index ascending. */
LBO_API int lbo_select_k(const float *d, int64_t n, int k, int64_t *out_idx, float *out_d) {
    lbo_pair *best = (lbo_pair *)malloc(sizeof(lbo_pair) * (size_t)(k > 0 ? k : 1));
    int cnt = 0;
    for (int64_t i = 0; i < n; i++) {
        if (d[i] != d[i]) continue;
        lbo_pair p = { d[i], i };
        topk_push(best, &cnt, k, p);
    }
    for (int j = 0; j < k; j++) {
        out_idx[j] = j < cnt ? best[j].id : -1;
        if (out_d) out_d[j] = j < cnt ? best[j].d : FLT_MAX;
    }
    free(best);
    return cnt;
}

/* ---------------------------------------------------------------------------
 * Product quantisation
 * ------------------------------------------------------------------------- */

/* internal/pq/adc_table.go:15-51: table[m*K+k] = L2Squared(q_m, c_mk), squared,
 * via l2SquaredImpl = L2SquaredFloat32 (dispatch.go:199).
 * codebooks: flat [M][K*sub] (encoder.go:17, persistence.go:25-33). */
LBO_API void lbo_adc_table(const float *q, const float *codebooks, int M, int K, int sub,
                           float *table) {
    for (int m = 0; m < M; m++)
        for (int k = 0; k < K; k++)
            table[m * K + k] = lbo_l2sq_f32(q + m * sub, codebooks + ((size_t)m * K + k) * sub, sub);
}

/* internal/simd/simd.go:345-355 (adcBatchGeneric): sequential fp32 sum over
 * j = 0..M-1 of table[j*256 + code], then sqrt via float64.  Stride is the
 * literal 256 (only K = 256 is coherent). */
LBO_API void lbo_adc_batch(const float *table, const uint8_t *codes, int M, int64_t n, float *out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        float sum = 0;
        const uint8_t *c = codes + (size_t)i * M;
        for (int j = 0; j < M; j++) sum += table[j * 256 + c[j]];
        out[i] = sqrt_via_f64(sum);
    }
}

/* internal/pq/adc_table.go:77-92 (ADCDistance): un-sqrt'd sum, stride K. */
LBO_API float lbo_adc_single(const float *table, const uint8_t *code, int M, int K) {
    float sum = 0;
    for (int m = 0; m < M; m++) sum += table[m * K + code[m]];
    return sum;
}

/* internal/pq/encoder.go:76-136 + internal/simd/simd.go:278-326: per subspace
 * the first centroid with the strictly smallest distance.  K <= 16 compares
 * L2Squared (encoder.go:96-119); larger K compares the sqrt'd
 * EuclideanDistanceBatchFlat results (simd.go:300-325). */
LBO_API void lbo_pq_encode(const float *vec, const float *codebooks, int M, int K, int sub,
                           uint8_t *codes) {
    for (int m = 0; m < M; m++) {
        const float *cb = codebooks + (size_t)m * K * sub;
        const float *v = vec + m * sub;
        int best = 0;
        float bd;
        if (K <= 16) {
            bd = FLT_MAX;
            for (int k = 0; k < K; k++) {
                float d = lbo_l2sq_f32(v, cb + (size_t)k * sub, sub);
                if (d < bd) { bd = d; best = k; }
            }
        } else {
            bd = lbo_euclid_f32(v, cb, sub);
            for (int k = 1; k < K; k++) {
                float d = lbo_euclid_f32(v, cb + (size_t)k * sub, sub);
                if (d < bd) { bd = d; best = k; }
            }
        }
        codes[m] = (uint8_t)best;
    }
}

LBO_API void lbo_pq_encode_batch(const float *vecs, int64_t n, const float *codebooks, int M, int K,
                                 int sub, uint8_t *codes) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++)
        lbo_pq_encode(vecs + (size_t)i * M * sub, codebooks, M, K, sub, codes + (size_t)i * M);
}

/* internal/pq/encoder.go:139-160 */
LBO_API void lbo_pq_decode(const uint8_t *codes, const float *codebooks, int M, int K, int sub,
                           float *vec) {
    for (int m = 0; m < M; m++)
        memcpy(vec + m * sub, codebooks + ((size_t)m * K + codes[m]) * sub, sizeof(float) * sub);
}

/* PQ search as the re-rank stage composes it (parallel_search.go:292-345 for
 * the ADC distances; north star: ADC scan -> top-k' -> fp32 re-rank -> top-k).
 * raw may be NULL: then the ADC top-k itself is returned (kprime ignored). */
LBO_API int lbo_pq_search(const float *codebooks, int M, int K, int sub, const uint8_t *codes,
                          int64_t n, const float *raw, const float *queries, int64_t nq, int k,
                          int kprime, const uint64_t *tomb, const uint64_t *allow,
                          float *out_d, int64_t *out_id) {
    if (K != 256 || k <= 0) return 1;
    int dim = M * sub;
    int kk = raw ? (kprime > k ? kprime : k) : k;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t qi = 0; qi < nq; qi++) {
        const float *q = queries + (size_t)qi * dim;
        float *table = (float *)malloc(sizeof(float) * (size_t)M * 256);
        lbo_adc_table(q, codebooks, M, K, sub, table);
        lbo_pair *best = (lbo_pair *)malloc(sizeof(lbo_pair) * (size_t)kk);
        int cnt = 0;
        for (int64_t i = 0; i < n; i++) {
            if (tomb && bit_test(tomb, i)) continue;
            if (allow && !bit_test(allow, i)) continue;
            float sum = 0;
            const uint8_t *c = codes + (size_t)i * M;
            for (int j = 0; j < M; j++) sum += table[j * 256 + c[j]];
            float d = sqrt_via_f64(sum);
            if (!(d < INFINITY)) continue;
            lbo_pair p = { d, i };
            topk_push(best, &cnt, kk, p);
        }
        if (raw) {
            lbo_pair *fin = (lbo_pair *)malloc(sizeof(lbo_pair) * (size_t)k);
            int fc = 0;
            for (int j = 0; j < cnt; j++) {
                lbo_pair p = { lbo_euclid_f32(q, raw + (size_t)best[j].id * dim, dim), best[j].id };
                if (!(p.d < INFINITY)) continue;
                topk_push(fin, &fc, k, p);
            }
            for (int j = 0; j < k; j++) {
                out_d[qi * k + j] = j < fc ? fin[j].d : FLT_MAX;
                out_id[qi * k + j] = j < fc ? fin[j].id : -1;
            }
            free(fin);
        } else {
            for (int j = 0; j < k; j++) {
                out_d[qi * k + j] = j < cnt ? best[j].d : FLT_MAX;
                out_id[qi * k + j] = j < cnt ? best[j].id : -1;
            }
        }
        free(best);
        free(table);
    }
    return 0;
}

/* Column compare -> 1 byte / row mask: internal/simd/simd.go:572-761
 * (matchInt64Generic / matchFloat32Generic).  op: 0 ==, 1 !=, 2 >, 3 >=, 4 <, 5 <=.
 * Then AND (simd.go:119-126) and pack to a dense allow-bitmap keyed by row. */
LBO_API void lbo_match_i64(const int64_t *col, int64_t n, int64_t val, int op, uint8_t *mask) {
    for (int64_t i = 0; i < n; i++) {
        int64_t v = col[i]; int r;
        switch (op) { case 0: r = v == val; break; case 1: r = v != val; break; case 2: r = v > val; break;
                      case 3: r = v >= val; break; case 4: r = v < val; break; default: r = v <= val; }
        mask[i] = (uint8_t)r;
    }
}
LBO_API void lbo_match_f32(const float *col, int64_t n, float val, int op, uint8_t *mask) {
    for (int64_t i = 0; i < n; i++) {
        float v = col[i]; int r;
        switch (op) { case 0: r = v == val; break; case 1: r = v != val; break; case 2: r = v > val; break;
                      case 3: r = v >= val; break; case 4: r = v < val; break; default: r = v <= val; }
        mask[i] = (uint8_t)r;
    }
}
LBO_API void lbo_mask_to_bitmap(const uint8_t *mask, int64_t n, uint64_t *bm) {
    memset(bm, 0, (size_t)((n + 63) / 64) * 8);
    for (int64_t i = 0; i < n; i++) if (mask[i]) bm[i >> 6] |= (uint64_t)1 << (i & 63);
}

/* ---------------------------------------------------------------------------
 * PQ training: TrainKMeans (internal/pq/kmeans.go:64-151) per subspace, as PQEncoder.Train
 * drives it (internal/pq/encoder.go:39-73: 20 iterations).
 *   init      : centroid c = data row init_idx[c]   (reference: rand.Perm(n)[:k] -- Go's math/rand
 *               stream is not reproducible here, so the caller supplies the indices; UNPINNED)
 *   E-step    : argmin_c L2Squared(vec, cent_c), strict '<' (first lowest index), kmeans.go:101-118
 *   sums      : centSum[j] += vec[j] in row order, fp32 (kmeans.go:126-130)
 *   M-step    : cent[j] = sum[j] / float32(count) (kmeans.go:134-141); an empty cluster is re-seeded
 *               from row (c*7919 + iter*104729) mod n  (reference: rand.Intn(n); UNPINNED stand-in)
 *   early stop: iter > 0 && changed < n/1000 + 1 (kmeans.go:146-148)
 * data: [n][dims] row-major, subspace m uses columns [m*sub, (m+1)*sub).  out: [M][K][sub].
 * ------------------------------------------------------------------------- */
LBO_API int lbo_pq_train(const float *data, int64_t n, int dims, int M, int K, int max_iter,
                         const int32_t *init_idx /* [M][K] */, float *out, int32_t *iters_run /* [M] or NULL */) {
    if (n < K || dims % M != 0) return 1;
    const int sub = dims / M;
    int32_t *assign = (int32_t *)malloc((size_t)n * sizeof(int32_t));
    int32_t *counts = (int32_t *)malloc((size_t)K * sizeof(int32_t));
    float *sums = (float *)malloc((size_t)K * sub * sizeof(float));
    for (int m = 0; m < M; m++) {
        float *cent = out + (size_t)m * K * sub;
        for (int c = 0; c < K; c++)
            memcpy(cent + (size_t)c * sub, data + (size_t)init_idx[(size_t)m * K + c] * dims + (size_t)m * sub,
                   (size_t)sub * sizeof(float));
        for (int64_t i = 0; i < n; i++) assign[i] = -1;
        int it = 0;
        for (; it < max_iter; it++) {
            memset(sums, 0, (size_t)K * sub * sizeof(float));
            memset(counts, 0, (size_t)K * sizeof(int32_t));
            int64_t changed = 0;
            for (int64_t i = 0; i < n; i++) {
                const float *vec = data + (size_t)i * dims + (size_t)m * sub;
                float best = FLT_MAX;
                int bc = -1;
                for (int c = 0; c < K; c++) {
                    float d = lbo_l2sq_f32(vec, cent + (size_t)c * sub, sub);
                    if (d < best) { best = d; bc = c; }
                }
                if (bc < 0) bc = 0; /* all distances NaN/Inf: Go would index -1 and panic; keep defined */
                if (assign[i] != bc) { changed++; assign[i] = bc; }
                counts[bc]++;
                float *cs = sums + (size_t)bc * sub;
                for (int j = 0; j < sub; j++) cs[j] += vec[j];
            }
            for (int c = 0; c < K; c++) {
                float *cc = cent + (size_t)c * sub;
                if (counts[c] > 0) {
                    float cnt = (float)counts[c];
                    for (int j = 0; j < sub; j++) cc[j] = sums[(size_t)c * sub + j] / cnt;
                } else {
                    int64_t idx = ((int64_t)c * 7919 + (int64_t)it * 104729) % n;
                    memcpy(cc, data + (size_t)idx * dims + (size_t)m * sub, (size_t)sub * sizeof(float));
                }
            }
            if (it > 0 && changed < n / 1000 + 1) { it++; break; }
        }
        if (iters_run) iters_run[m] = it;
    }
    free(assign); free(counts); free(sums);
    return 0;
}

/* ---------------------------------------------------------------------------------------------
 * HNSW layer search: ArrowHNSW.searchLayer (internal/store/arrow_hnsw.go:1108-1385) over the adjacency
 * layout of GraphData (internal/store/types/graph_data.go:605-670: per node `counts[id]` neighbours at
 * neighbors[id * max_degree ...]).
 *   candidates: min-heap of nodes to expand; result set: max-heap of the best <= ef nodes found so far.
 *   pop the closest candidate; stop when it is strictly worse than the worst result and the result set is full
 *   (arrow_hnsw.go:1322-1329); for every not-yet-visited neighbour (in list order) compute the distance and, if the
 *   result set is not full or the distance is strictly below its worst, push it on both heaps and evict the
 *   worst result when the set exceeds ef (:1349-1370).  Output ascending by distance (:1377-1383).
 * The reference's container/heap orders on Dist only, so the pop order among EQUAL distances is an artefact of
 * Go's heap algorithm (unpinned); here both heaps order on (distance, id).
 * -------------------------------------------------------------------------------------------- */
typedef struct { float d; uint32_t id; } hn_cand;
static inline int hn_less(hn_cand a, hn_cand b) { return a.d < b.d || (a.d == b.d && a.id < b.id); }

static void hn_push(hn_cand *h, int *n, hn_cand c, int maxheap) {
    int i = (*n)++;
    h[i] = c;
    while (i > 0) {
        int p = (i - 1) / 2;
        int up = maxheap ? hn_less(h[p], h[i]) : hn_less(h[i], h[p]);
        if (!up) break;
        hn_cand t = h[p]; h[p] = h[i]; h[i] = t;
        i = p;
    }
}
static hn_cand hn_pop(hn_cand *h, int *n, int maxheap) {
    hn_cand top = h[0];
    h[0] = h[--(*n)];
    int i = 0;
    for (;;) {
        int l = 2 * i + 1, r = l + 1, b = i;
        if (l < *n && (maxheap ? hn_less(h[b], h[l]) : hn_less(h[l], h[b]))) b = l;
        if (r < *n && (maxheap ? hn_less(h[b], h[r]) : hn_less(h[r], h[b]))) b = r;
        if (b == i) break;
        hn_cand t = h[b]; h[b] = h[i]; h[i] = t;
        i = b;
    }
    return top;
}

/* One query.  out_ids / out_d: [ef] ascending, padded with 0xffffffff / FLT_MAX.  Returns the number of results;
 * *n_visited (optional) = nodes whose distance was computed (the entry point included). */
LBO_API int lbo_hnsw_search_layer(int metric, int dtype, const void *db, int64_t n, int dim,
                                  const uint32_t *neighbors, const int32_t *counts, int max_degree,
                                  const void *query, uint32_t entry, int ef, uint32_t *out_ids, float *out_d,
                                  int64_t *n_visited) {
    if (ef <= 0 || n <= 0 || entry >= (uint64_t)n) return -1;
    const size_t es = elem_size(dtype);
    uint8_t *visited = (uint8_t *)calloc((size_t)(n + 7) / 8, 1);
    /* every push onto the candidate heap is also a result-set push, but evictions do not remove candidates:
     * the candidate heap can hold every node that was ever accepted */
    int cap = 1024, nc = 0, nr = 0;
    hn_cand *cand = (hn_cand *)malloc(sizeof(hn_cand) * cap);
    hn_cand *res = (hn_cand *)malloc(sizeof(hn_cand) * (ef + 2));
    int64_t nv = 0;
/* pair_distance: the dot metric is already negated (a distance, distance_resolvers.go:11-16) */
#define HN_DIST(id) pair_distance(metric, dtype, query, (const char *)db + (size_t)(id) * dim * es, dim)
    hn_cand ep = {HN_DIST(entry), entry};
    nv++;
    hn_push(cand, &nc, ep, 0);
    hn_push(res, &nr, ep, 1);
    visited[entry >> 3] |= (uint8_t)(1u << (entry & 7));
    while (nc > 0) {
        hn_cand cur = hn_pop(cand, &nc, 0);
        if (nr > 0 && cur.d > res[0].d && nr >= ef) break;
        const int cnt = counts ? counts[cur.id] : max_degree;
        for (int i = 0; i < cnt && i < max_degree; i++) {
            const uint32_t nb = neighbors[(size_t)cur.id * max_degree + i];
            if (nb >= (uint64_t)n) continue; /* padding / dangling id */
            if (visited[nb >> 3] & (1u << (nb & 7))) continue;
            visited[nb >> 3] |= (uint8_t)(1u << (nb & 7));
            hn_cand c = {HN_DIST(nb), nb};
            nv++;
            if (nr < ef || c.d < res[0].d) {
                if (nc + 1 >= cap) { cap *= 2; cand = (hn_cand *)realloc(cand, sizeof(hn_cand) * cap); }
                hn_push(cand, &nc, c, 0);
                hn_push(res, &nr, c, 1);
                if (nr > ef) hn_pop(res, &nr, 1);
            }
        }
    }
#undef HN_DIST
    const int count = nr;
    for (int i = count - 1; i >= 0; i--) {
        hn_cand c = hn_pop(res, &nr, 1);
        out_ids[i] = c.id;
        out_d[i] = c.d;
    }
    for (int i = count; i < ef; i++) { out_ids[i] = 0xffffffffu; out_d[i] = 3.402823466e+38f; }
    if (n_visited) *n_visited = nv;
    free(visited); free(cand); free(res);
    return count;
}

/* Batch of queries (OpenMP over queries), entry point per query. */
LBO_API int lbo_hnsw_search_layer_batch(int metric, int dtype, const void *db, int64_t n, int dim,
                                        const uint32_t *neighbors, const int32_t *counts, int max_degree,
                                        const void *queries, int64_t nq, const uint32_t *entries, int ef,
                                        uint32_t *out_ids, float *out_d, int64_t *n_visited) {
    const size_t qs = (size_t)dim * elem_size(dtype);
    int bad = 0;
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t q = 0; q < nq; q++) {
        int rc = lbo_hnsw_search_layer(metric, dtype, db, n, dim, neighbors, counts, max_degree,
                                       (const char *)queries + (size_t)q * qs, entries[q], ef,
                                       out_ids + (size_t)q * ef, out_d + (size_t)q * ef,
                                       n_visited ? n_visited + q : NULL);
        if (rc < 0) bad = 1;
    }
    return bad ? -1 : 0;
}

/* ---------------------------------------------------------------------------------------------
 * SQ8 scalar quantisation (internal/simd/sq8.go:70-104) and the inline de-quantising distance of the
 * HNSW computer (internal/store/arrow_hnsw.go:1176-1186).
 * -------------------------------------------------------------------------------------------- */
LBO_API void lbo_quantize_sq8(const float *src, int64_t n, float minv, float maxv, uint8_t *dst) {
    float scale = 255.0f / (maxv - minv);
    if (maxv == minv) scale = 0;
    for (int64_t i = 0; i < n; i++) {
        float val = (src[i] - minv) * scale;
        if (val < 0) val = 0;
        if (val > 255) val = 255;
        dst[i] = (uint8_t)val;
    }
}
LBO_API void lbo_compute_bounds(const float *v, int64_t n, float *mn, float *mx) {
    if (n == 0) { *mn = 0; *mx = 0; return; }
    float a = v[0], b = v[0];
    for (int64_t i = 1; i < n; i++) { if (v[i] < a) a = v[i]; if (v[i] > b) b = v[i]; }
    *mn = a; *mx = b;
}
LBO_API void lbo_dequantize_sq8(const uint8_t *src, int64_t n, float minv, float maxv, float *dst) {
    const float scale = (maxv - minv) / 255.0f;
    for (int64_t i = 0; i < n; i++) dst[i] = minv + (float)src[i] * scale;
}
LBO_API void lbo_sq8_dequant_distance_batch(const float *q, const uint8_t *rows, int64_t n, int dim, float minv,
                                            float maxv, float *out) {
    const float scale = (maxv - minv) / 255.0f;
    for (int64_t r = 0; r < n; r++) {
        float sum = 0;
        for (int i = 0; i < dim; i++) {
            float deq = minv + (float)rows[r * dim + i] * scale;
            float diff = q[i] - deq;
            sum += diff * diff;
        }
        out[r] = (float)sqrt((double)sum);
    }
}
/* simd.FindNearestCentroid (internal/simd/simd.go:278-326) */
LBO_API int lbo_find_nearest_centroid(const float *query, const float *cent, int sub, int k, float *out_d) {
    if (k <= 8) {
        float best = FLT_MAX; int bi = 0;
        for (int i = 0; i < k; i++) {
            float d = lbo_l2sq_f32(query, cent + (size_t)i * sub, sub);
            if (d < best) { best = d; bi = i; }
        }
        *out_d = best;
        return bi;
    }
    float best = lbo_euclid_f32(query, cent, sub); int bi = 0;
    for (int i = 1; i < k; i++) {
        float d = lbo_euclid_f32(query, cent + (size_t)i * sub, sub);
        if (d < best) { best = d; bi = i; }
    }
    *out_d = best;
    return bi;
}
