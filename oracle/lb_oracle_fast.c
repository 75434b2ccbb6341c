/*
 * lb_oracle_fast.c -- TEST / BENCH INFRASTRUCTURE ONLY ("O-fast").
 *
 * The timed CPU-SIMD baseline: a C restatement of the *structure* of the
 * reference's amd64 assembly kernels plus its per-query scan + size-k heap,
 * with OpenMP over all host cores.  It is a restatement (a "port"), not the Go
 * binary: Go is not installed here and the reference's amd64 simd package does
 * not link as shipped.  Results are checked against lb_oracle.c
 * within the reference's own SIMD-vs-scalar tolerance (1e-5 .. 1e-3 relative,
 * internal/simd/simd_test.go:105-122) -- FMA and wider lanes change the
 * summation order exactly as the reference's AVX kernels do.
 *
 * Structures followed:
 *   fp32 L2     internal/simd/distance_amd64.s:21-143   4 x ZMM accumulators, FMA
 *   fp32 cosine internal/simd/distance_amd64.s:149-296  dot, |a|^2, |b|^2 in one pass
 *   fp16        internal/simd/simd_amd64.s:303-530      VCVTPH2PS + FMA, fp32 accumulate
 *   int8 L2     internal/simd/simd_amd64.s:735-815      sign-extend, sub, madd -> int32
 *   ADC         internal/simd/pq_amd64.s:14-168         gather table[j*256+code], add in j order
 *   top-k       internal/store/adaptive_index.go:176-222 size-k max-heap, strict <
 *
 * Runtime dispatch: AVX-512 (F+BW+VL) if the host has it, else AVX2+F16C+FMA,
 * else scalar.
 */
#include <immintrin.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define LBF_API __attribute__((visibility("default")))
#define T512 __attribute__((target("avx512f,avx512bw,avx512vl,avx512dq,f16c,fma")))
#define T256 __attribute__((target("avx2,f16c,fma")))

enum { LBO_L2 = 0, LBO_COSINE = 1, LBO_DOT = 2 };
enum { LBO_F32 = 0, LBO_F16 = 1, LBO_I8 = 2 };

static int g_isa = -1; /* 2 = avx512, 1 = avx2, 0 = scalar */
static int isa(void) {
    if (g_isa < 0) {
        __builtin_cpu_init();
        if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") &&
            __builtin_cpu_supports("avx512vl") && __builtin_cpu_supports("avx512dq"))
            g_isa = 2;
        else if (__builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma") &&
                 __builtin_cpu_supports("f16c"))
            g_isa = 1;
        else
            g_isa = 0;
    }
    return g_isa;
}
LBF_API int lbf_isa(void) { return isa(); }
LBF_API int lbf_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1 to every rank; the timed CPU baseline sets its own thread count
 * (all host cores the process may run on) so the N>1 reference arm is not a 1-thread run. */
LBF_API void lbf_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* three partial sums of one pair: dot, |a|^2, |b|^2, or the squared diff */
typedef struct { float dot, na, nb, l2; } acc_t;

/* ------------------------------ AVX-512 --------------------------------- */
T512 static inline float hsum512(__m512 v) { return _mm512_reduce_add_ps(v); }

T512 static float l2sq_f32_512(const float *a, const float *b, int n) {
    __m512 s0 = _mm512_setzero_ps(), s1 = s0, s2 = s0, s3 = s0;
    int i = 0;
    for (; i + 64 <= n; i += 64) {
        __m512 d0 = _mm512_sub_ps(_mm512_loadu_ps(a + i), _mm512_loadu_ps(b + i));
        __m512 d1 = _mm512_sub_ps(_mm512_loadu_ps(a + i + 16), _mm512_loadu_ps(b + i + 16));
        __m512 d2 = _mm512_sub_ps(_mm512_loadu_ps(a + i + 32), _mm512_loadu_ps(b + i + 32));
        __m512 d3 = _mm512_sub_ps(_mm512_loadu_ps(a + i + 48), _mm512_loadu_ps(b + i + 48));
        s0 = _mm512_fmadd_ps(d0, d0, s0); s1 = _mm512_fmadd_ps(d1, d1, s1);
        s2 = _mm512_fmadd_ps(d2, d2, s2); s3 = _mm512_fmadd_ps(d3, d3, s3);
    }
    for (; i + 16 <= n; i += 16) {
        __m512 d = _mm512_sub_ps(_mm512_loadu_ps(a + i), _mm512_loadu_ps(b + i));
        s0 = _mm512_fmadd_ps(d, d, s0);
    }
    if (i < n) { /* masked tail, distance_amd64.s:110-128 */
        __mmask16 m = (__mmask16)((1u << (n - i)) - 1);
        __m512 d = _mm512_sub_ps(_mm512_maskz_loadu_ps(m, a + i), _mm512_maskz_loadu_ps(m, b + i));
        s0 = _mm512_fmadd_ps(d, d, s0);
    }
    return hsum512(_mm512_add_ps(_mm512_add_ps(s0, s1), _mm512_add_ps(s2, s3)));
}

T512 static void dot3_f32_512(const float *a, const float *b, int n, acc_t *o) {
    __m512 d = _mm512_setzero_ps(), na = d, nb = d;
    int i = 0;
    for (; i + 16 <= n; i += 16) {
        __m512 x = _mm512_loadu_ps(a + i), y = _mm512_loadu_ps(b + i);
        d = _mm512_fmadd_ps(x, y, d); na = _mm512_fmadd_ps(x, x, na); nb = _mm512_fmadd_ps(y, y, nb);
    }
    if (i < n) {
        __mmask16 m = (__mmask16)((1u << (n - i)) - 1);
        __m512 x = _mm512_maskz_loadu_ps(m, a + i), y = _mm512_maskz_loadu_ps(m, b + i);
        d = _mm512_fmadd_ps(x, y, d); na = _mm512_fmadd_ps(x, x, na); nb = _mm512_fmadd_ps(y, y, nb);
    }
    o->dot = hsum512(d); o->na = hsum512(na); o->nb = hsum512(nb);
}

T512 static float dot_f32_512(const float *a, const float *b, int n) {
    __m512 s0 = _mm512_setzero_ps(), s1 = s0;
    int i = 0;
    for (; i + 32 <= n; i += 32) {
        s0 = _mm512_fmadd_ps(_mm512_loadu_ps(a + i), _mm512_loadu_ps(b + i), s0);
        s1 = _mm512_fmadd_ps(_mm512_loadu_ps(a + i + 16), _mm512_loadu_ps(b + i + 16), s1);
    }
    for (; i + 16 <= n; i += 16)
        s0 = _mm512_fmadd_ps(_mm512_loadu_ps(a + i), _mm512_loadu_ps(b + i), s0);
    if (i < n) {
        __mmask16 m = (__mmask16)((1u << (n - i)) - 1);
        s0 = _mm512_fmadd_ps(_mm512_maskz_loadu_ps(m, a + i), _mm512_maskz_loadu_ps(m, b + i), s0);
    }
    return hsum512(_mm512_add_ps(s0, s1));
}

T512 static inline __m512 ldh512(const uint16_t *p) {
    return _mm512_cvtph_ps(_mm256_loadu_si256((const __m256i *)p));
}
T512 static inline __m512 ldh512_tail(const uint16_t *p, int r) {
    __mmask16 m = (__mmask16)((1u << r) - 1);
    return _mm512_cvtph_ps(_mm256_maskz_loadu_epi16(m, p));
}

T512 static float l2sq_f16_512(const uint16_t *a, const uint16_t *b, int n) {
    __m512 s0 = _mm512_setzero_ps(), s1 = s0;
    int i = 0;
    for (; i + 32 <= n; i += 32) {
        __m512 d0 = _mm512_sub_ps(ldh512(a + i), ldh512(b + i));
        __m512 d1 = _mm512_sub_ps(ldh512(a + i + 16), ldh512(b + i + 16));
        s0 = _mm512_fmadd_ps(d0, d0, s0); s1 = _mm512_fmadd_ps(d1, d1, s1);
    }
    for (; i + 16 <= n; i += 16) {
        __m512 d = _mm512_sub_ps(ldh512(a + i), ldh512(b + i));
        s0 = _mm512_fmadd_ps(d, d, s0);
    }
    if (i < n) {
        __m512 d = _mm512_sub_ps(ldh512_tail(a + i, n - i), ldh512_tail(b + i, n - i));
        s0 = _mm512_fmadd_ps(d, d, s0);
    }
    return hsum512(_mm512_add_ps(s0, s1));
}

T512 static void dot3_f16_512(const uint16_t *a, const uint16_t *b, int n, acc_t *o) {
    __m512 d = _mm512_setzero_ps(), na = d, nb = d;
    int i = 0;
    for (; i + 16 <= n; i += 16) {
        __m512 x = ldh512(a + i), y = ldh512(b + i);
        d = _mm512_fmadd_ps(x, y, d); na = _mm512_fmadd_ps(x, x, na); nb = _mm512_fmadd_ps(y, y, nb);
    }
    if (i < n) {
        __m512 x = ldh512_tail(a + i, n - i), y = ldh512_tail(b + i, n - i);
        d = _mm512_fmadd_ps(x, y, d); na = _mm512_fmadd_ps(x, x, na); nb = _mm512_fmadd_ps(y, y, nb);
    }
    o->dot = hsum512(d); o->na = hsum512(na); o->nb = hsum512(nb);
}

T512 static float dot_f16_512(const uint16_t *a, const uint16_t *b, int n) {
    __m512 s0 = _mm512_setzero_ps(), s1 = s0;
    int i = 0;
    for (; i + 32 <= n; i += 32) {
        s0 = _mm512_fmadd_ps(ldh512(a + i), ldh512(b + i), s0);
        s1 = _mm512_fmadd_ps(ldh512(a + i + 16), ldh512(b + i + 16), s1);
    }
    for (; i + 16 <= n; i += 16) s0 = _mm512_fmadd_ps(ldh512(a + i), ldh512(b + i), s0);
    if (i < n) s0 = _mm512_fmadd_ps(ldh512_tail(a + i, n - i), ldh512_tail(b + i, n - i), s0);
    return hsum512(_mm512_add_ps(s0, s1));
}

/* int8: sign-extend to int16, (sub,) madd to int32 -- simd_amd64.s:735-815 widened to 512 bit */
T512 static int32_t l2sq_i8_512(const int8_t *a, const int8_t *b, int n) {
    __m512i s = _mm512_setzero_si512();
    int i = 0;
    for (; i + 32 <= n; i += 32) {
        __m512i x = _mm512_cvtepi8_epi16(_mm256_loadu_si256((const __m256i *)(a + i)));
        __m512i y = _mm512_cvtepi8_epi16(_mm256_loadu_si256((const __m256i *)(b + i)));
        __m512i d = _mm512_sub_epi16(x, y);
        s = _mm512_add_epi32(s, _mm512_madd_epi16(d, d));
    }
    int32_t r = _mm512_reduce_add_epi32(s);
    for (; i < n; i++) { int32_t d = (int32_t)a[i] - (int32_t)b[i]; r += d * d; }
    return r;
}
T512 static int32_t dot_i8_512(const int8_t *a, const int8_t *b, int n) {
    __m512i s = _mm512_setzero_si512();
    int i = 0;
    for (; i + 32 <= n; i += 32) {
        __m512i x = _mm512_cvtepi8_epi16(_mm256_loadu_si256((const __m256i *)(a + i)));
        __m512i y = _mm512_cvtepi8_epi16(_mm256_loadu_si256((const __m256i *)(b + i)));
        s = _mm512_add_epi32(s, _mm512_madd_epi16(x, y));
    }
    int32_t r = _mm512_reduce_add_epi32(s);
    for (; i < n; i++) r += (int32_t)a[i] * (int32_t)b[i];
    return r;
}

/* ADC: 16 codes at a time, gather per subspace, add in j order, sqrt */
T512 static void adc_batch_512(const float *table, const uint8_t *codes, int M, int64_t n, float *out) {
    int64_t i = 0;
    __m512i rowoff = _mm512_mullo_epi32(_mm512_set_epi32(15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0),
                                        _mm512_set1_epi32(M));
    for (; i + 16 <= n; i += 16) {
        const uint8_t *base = codes + (size_t)i * M;
        __m512 sum = _mm512_setzero_ps();
        int j = 0;
        /* last group may read up to 3 bytes past row 15's byte j: keep the dword gather in bounds */
        for (; j < M; j++) {
            __m128i cb;
            uint8_t tmp[16];
            for (int r = 0; r < 16; r++) tmp[r] = base[(size_t)r * M + j];
            cb = _mm_loadu_si128((const __m128i *)tmp);
            __m512i idx = _mm512_cvtepu8_epi32(cb);
            __m512 v = _mm512_i32gather_ps(idx, table + (size_t)j * 256, 4);
            sum = _mm512_add_ps(sum, v);
        }
        (void)rowoff;
        _mm512_storeu_ps(out + i, _mm512_sqrt_ps(sum));
    }
    for (; i < n; i++) {
        float s = 0;
        const uint8_t *c = codes + (size_t)i * M;
        for (int j = 0; j < M; j++) s += table[j * 256 + c[j]];
        out[i] = sqrtf(s);
    }
}

/* ------------------------------ AVX2 ------------------------------------ */
T256 static inline float hsum256(__m256 v) {
    __m128 lo = _mm256_castps256_ps128(v), hi = _mm256_extractf128_ps(v, 1);
    lo = _mm_add_ps(lo, hi);
    lo = _mm_hadd_ps(lo, lo);
    lo = _mm_hadd_ps(lo, lo);
    return _mm_cvtss_f32(lo);
}
T256 static float l2sq_f32_256(const float *a, const float *b, int n) {
    __m256 s0 = _mm256_setzero_ps(), s1 = s0;
    int i = 0;
    for (; i + 16 <= n; i += 16) {
        __m256 d0 = _mm256_sub_ps(_mm256_loadu_ps(a + i), _mm256_loadu_ps(b + i));
        __m256 d1 = _mm256_sub_ps(_mm256_loadu_ps(a + i + 8), _mm256_loadu_ps(b + i + 8));
        s0 = _mm256_fmadd_ps(d0, d0, s0); s1 = _mm256_fmadd_ps(d1, d1, s1);
    }
    float r = hsum256(_mm256_add_ps(s0, s1));
    for (; i < n; i++) { float d = a[i] - b[i]; r += d * d; }
    return r;
}
T256 static void dot3_f32_256(const float *a, const float *b, int n, acc_t *o) {
    __m256 d = _mm256_setzero_ps(), na = d, nb = d;
    int i = 0;
    for (; i + 8 <= n; i += 8) {
        __m256 x = _mm256_loadu_ps(a + i), y = _mm256_loadu_ps(b + i);
        d = _mm256_fmadd_ps(x, y, d); na = _mm256_fmadd_ps(x, x, na); nb = _mm256_fmadd_ps(y, y, nb);
    }
    o->dot = hsum256(d); o->na = hsum256(na); o->nb = hsum256(nb);
    for (; i < n; i++) { o->dot += a[i] * b[i]; o->na += a[i] * a[i]; o->nb += b[i] * b[i]; }
}
T256 static inline __m256 ldh256(const uint16_t *p) {
    return _mm256_cvtph_ps(_mm_loadu_si128((const __m128i *)p));
}
T256 static void dot3_f16_256(const uint16_t *a, const uint16_t *b, int n, acc_t *o, int want_l2) {
    __m256 d = _mm256_setzero_ps(), na = d, nb = d, l2 = d;
    int i = 0;
    for (; i + 8 <= n; i += 8) {
        __m256 x = ldh256(a + i), y = ldh256(b + i);
        if (want_l2) { __m256 df = _mm256_sub_ps(x, y); l2 = _mm256_fmadd_ps(df, df, l2); }
        else { d = _mm256_fmadd_ps(x, y, d); na = _mm256_fmadd_ps(x, x, na); nb = _mm256_fmadd_ps(y, y, nb); }
    }
    o->dot = hsum256(d); o->na = hsum256(na); o->nb = hsum256(nb); o->l2 = hsum256(l2);
    for (; i < n; i++) {
        float x = _cvtsh_ss(a[i]), y = _cvtsh_ss(b[i]);
        float df = x - y;
        o->l2 += df * df; o->dot += x * y; o->na += x * x; o->nb += y * y;
    }
}

/* ------------------------------ scalar ---------------------------------- */
static float h2f_s(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16, exp = (h >> 10) & 0x1fu, man = h & 0x3ffu, bits;
    if (exp == 0) {
        if (!man) bits = sign;
        else { int e = -1; do { man <<= 1; e++; } while (!(man & 0x400u)); man &= 0x3ffu;
               bits = sign | ((uint32_t)(127 - 15 - e) << 23) | (man << 13); }
    } else if (exp == 31) bits = sign | 0x7f800000u | (man << 13);
    else bits = sign | ((exp + 112) << 23) | (man << 13);
    float f; memcpy(&f, &bits, 4); return f;
}

/* ------------------------------ dispatch -------------------------------- */
static float pair_fast(int metric, int dtype, const void *pa, const void *pb, int n) {
    int v = isa();
    acc_t o = {0, 0, 0, 0};
    if (dtype == LBO_F32) {
        const float *a = pa, *b = pb;
        if (metric == LBO_L2) {
            float s;
            if (v == 2) s = l2sq_f32_512(a, b, n);
            else if (v == 1) s = l2sq_f32_256(a, b, n);
            else { s = 0; for (int i = 0; i < n; i++) { float d = a[i] - b[i]; s += d * d; } }
            return sqrtf(s);
        }
        if (metric == LBO_DOT && v == 2) return -dot_f32_512(a, b, n);
        if (v == 2) dot3_f32_512(a, b, n, &o);
        else if (v == 1) dot3_f32_256(a, b, n, &o);
        else for (int i = 0; i < n; i++) { o.dot += a[i] * b[i]; o.na += a[i] * a[i]; o.nb += b[i] * b[i]; }
    } else if (dtype == LBO_F16) {
        const uint16_t *a = pa, *b = pb;
        if (metric == LBO_L2) {
            float s;
            if (v == 2) s = l2sq_f16_512(a, b, n);
            else if (v == 1) { dot3_f16_256(a, b, n, &o, 1); s = o.l2; }
            else { s = 0; for (int i = 0; i < n; i++) { float d = h2f_s(a[i]) - h2f_s(b[i]); s += d * d; } }
            return sqrtf(s);
        }
        if (metric == LBO_DOT && v == 2) return -dot_f16_512(a, b, n);
        if (v == 2) dot3_f16_512(a, b, n, &o);
        else if (v == 1) dot3_f16_256(a, b, n, &o, 0);
        else for (int i = 0; i < n; i++) { float x = h2f_s(a[i]), y = h2f_s(b[i]);
                                           o.dot += x * y; o.na += x * x; o.nb += y * y; }
    } else { /* int8 */
        const int8_t *a = pa, *b = pb;
        int32_t s = 0;
        if (metric == LBO_L2) {
            if (v == 2) s = l2sq_i8_512(a, b, n);
            else for (int i = 0; i < n; i++) { int32_t d = (int32_t)a[i] - b[i]; s += d * d; }
            return sqrtf((float)s);
        }
        if (v == 2) s = dot_i8_512(a, b, n);
        else for (int i = 0; i < n; i++) s += (int32_t)a[i] * b[i];
        return -(float)s;
    }
    if (metric == LBO_DOT) return -o.dot;
    if (o.na <= 0 || o.nb <= 0) return 1.0f; /* simd_amd64.go:763,782 */
    return 1.0f - o.dot / (float)sqrt((double)o.na * (double)o.nb);
}

static size_t esz(int dtype) { return dtype == LBO_F32 ? 4 : dtype == LBO_F16 ? 2 : 1; }

LBF_API float lbf_distance(int metric, int dtype, const void *a, const void *b, int n) {
    return pair_fast(metric, dtype, a, b, n);
}

LBF_API void lbf_batch_flat(int metric, int dtype, const void *q, const void *flat, int64_t n,
                            int dim, float *out) {
    size_t st = (size_t)dim * esz(dtype);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++)
        out[i] = pair_fast(metric, dtype, q, (const char *)flat + (size_t)i * st, dim);
}

/* size-k max-heap keyed on (d, id): adaptive_index.go:176-222, 327-350 */
typedef struct { float d; int64_t id; } hp_t;
static inline int hp_gt(hp_t a, hp_t b) { return a.d > b.d || (a.d == b.d && a.id > b.id); }
static inline void hp_down(hp_t *h, int n, int i) {
    for (;;) {
        int l = 2 * i + 1, r = l + 1, m = i;
        if (l < n && hp_gt(h[l], h[m])) m = l;
        if (r < n && hp_gt(h[r], h[m])) m = r;
        if (m == i) return;
        hp_t t = h[i]; h[i] = h[m]; h[m] = t; i = m;
    }
}
static inline void hp_up(hp_t *h, int i) {
    while (i > 0) { int p = (i - 1) / 2; if (!hp_gt(h[i], h[p])) return;
                    hp_t t = h[i]; h[i] = h[p]; h[p] = t; i = p; }
}
static inline void hp_offer(hp_t *h, int *cnt, int k, float d, int64_t id) {
    if (*cnt < k) { h[*cnt].d = d; h[*cnt].id = id; hp_up(h, (*cnt)++); }
    else if (d < h[0].d || (d == h[0].d && id < h[0].id)) { h[0].d = d; h[0].id = id; hp_down(h, k, 0); }
}
static int hp_cmp(const void *a, const void *b) {
    const hp_t *x = a, *y = b;
    if (x->d < y->d) return -1; if (x->d > y->d) return 1;
    return x->id < y->id ? -1 : x->id > y->id;
}
static void hp_emit(hp_t *h, int cnt, int k, int64_t id_base, float *od, int64_t *oi) {
    qsort(h, (size_t)cnt, sizeof(hp_t), hp_cmp);
    for (int j = 0; j < k; j++) { od[j] = j < cnt ? h[j].d : FLT_MAX; oi[j] = j < cnt ? h[j].id + id_base : -1; }
}
static inline int bt(const uint64_t *bm, int64_t i) { return (int)((bm[i >> 6] >> (i & 63)) & 1u); }

/* Brute-force k-NN, parallel over queries (the reference runs one goroutine per request). */
LBF_API int lbf_search(int metric, int dtype, const void *db, int64_t n, int dim, const void *queries,
                       int64_t nq, int k, const uint64_t *tomb, const uint64_t *allow,
                       int64_t id_base, float *out_d, int64_t *out_id) {
    if (k <= 0 || dim <= 0) return 1;
    size_t st = (size_t)dim * esz(dtype);
#pragma omp parallel
    {
        hp_t *h = (hp_t *)malloc(sizeof(hp_t) * (size_t)k);
#pragma omp for schedule(dynamic, 1)
        for (int64_t qi = 0; qi < nq; qi++) {
            int cnt = 0;
            const char *q = (const char *)queries + (size_t)qi * st;
            for (int64_t i = 0; i < n; i++) {
                if (tomb && bt(tomb, i)) continue;
                if (allow && !bt(allow, i)) continue;
                float d = pair_fast(metric, dtype, q, (const char *)db + (size_t)i * st, dim);
                if (!(d < INFINITY)) continue;
                hp_offer(h, &cnt, k, d, i);
            }
            hp_emit(h, cnt, k, id_base, out_d + qi * k, out_id + qi * k);
        }
        free(h);
    }
    return 0;
}

LBF_API void lbf_adc_batch(const float *table, const uint8_t *codes, int M, int64_t n, float *out) {
    const int64_t chunk = 4096;
    int64_t nchunks = (n + chunk - 1) / chunk;
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < nchunks; c++) {
        int64_t lo = c * chunk, hi = lo + chunk < n ? lo + chunk : n;
        if (isa() == 2) adc_batch_512(table, codes + (size_t)lo * M, M, hi - lo, out + lo);
        else for (int64_t i = lo; i < hi; i++) {
            float s = 0; const uint8_t *cc = codes + (size_t)i * M;
            for (int j = 0; j < M; j++) s += table[j * 256 + cc[j]];
            out[i] = sqrtf(s);
        }
    }
}

/* PQ search: per query build LUT, ADC-scan all codes keeping top-kprime, fp32 re-rank. */
LBF_API int lbf_pq_search(const float *codebooks, int M, int sub, const uint8_t *codes, int64_t n,
                          const float *raw, const float *queries, int64_t nq, int k, int kprime,
                          float *out_d, int64_t *out_id) {
    int dim = M * sub;
    int kk = raw ? (kprime > k ? kprime : k) : k;
#pragma omp parallel
    {
        float *table = (float *)malloc(sizeof(float) * (size_t)M * 256);
        float *dist = (float *)malloc(sizeof(float) * 4096);
        hp_t *h = (hp_t *)malloc(sizeof(hp_t) * (size_t)kk);
        hp_t *f = (hp_t *)malloc(sizeof(hp_t) * (size_t)k);
#pragma omp for schedule(dynamic, 1)
        for (int64_t qi = 0; qi < nq; qi++) {
            const float *q = queries + (size_t)qi * dim;
            for (int m = 0; m < M; m++)
                for (int c = 0; c < 256; c++) {
                    float s = pair_fast(LBO_L2, LBO_F32, q + m * sub, codebooks + ((size_t)m * 256 + c) * sub, sub);
                    table[m * 256 + c] = s * s;
                }
            int cnt = 0;
            for (int64_t lo = 0; lo < n; lo += 4096) {
                int64_t len = lo + 4096 < n ? 4096 : n - lo;
                if (isa() == 2) adc_batch_512(table, codes + (size_t)lo * M, M, len, dist);
                else for (int64_t i = 0; i < len; i++) {
                    float s = 0; const uint8_t *cc = codes + (size_t)(lo + i) * M;
                    for (int j = 0; j < M; j++) s += table[j * 256 + cc[j]];
                    dist[i] = sqrtf(s);
                }
                for (int64_t i = 0; i < len; i++) hp_offer(h, &cnt, kk, dist[i], lo + i);
            }
            if (raw) {
                int fc = 0;
                for (int j = 0; j < cnt; j++)
                    hp_offer(f, &fc, k, pair_fast(LBO_L2, LBO_F32, q, raw + (size_t)h[j].id * dim, dim), h[j].id);
                hp_emit(f, fc, k, 0, out_d + qi * k, out_id + qi * k);
            } else hp_emit(h, cnt, k, 0, out_d + qi * k, out_id + qi * k);
        }
        free(table); free(dist); free(h); free(f);
    }
    return 0;
}

/* Re-rank (parallel_search.go:147-365): gather candidate rows, batch distance, sort, top-k. */
LBF_API int lbf_rerank(int metric, int dtype, const void *db, int64_t n, int dim, const void *queries,
                       int64_t nq, const int64_t *cand, int c, int k, const uint64_t *tomb,
                       const uint64_t *allow, float *out_d, int64_t *out_id) {
    size_t st = (size_t)dim * esz(dtype);
#pragma omp parallel
    {
        hp_t *h = (hp_t *)malloc(sizeof(hp_t) * (size_t)k);
#pragma omp for schedule(dynamic, 16)
        for (int64_t qi = 0; qi < nq; qi++) {
            int cnt = 0;
            const char *q = (const char *)queries + (size_t)qi * st;
            const int64_t *ids = cand + (size_t)qi * c;
            for (int j = 0; j < c; j++) {
                int64_t id = ids[j];
                if (id < 0 || id >= n) continue;
                if (allow && !bt(allow, id)) continue;
                if (tomb && bt(tomb, id)) continue;
                float d = pair_fast(metric, dtype, q, (const char *)db + (size_t)id * st, dim);
                if (!(d < INFINITY)) continue;
                hp_offer(h, &cnt, k, d, id);
            }
            hp_emit(h, cnt, k, 0, out_d + qi * k, out_id + qi * k);
        }
        free(h);
    }
    return 0;
}
