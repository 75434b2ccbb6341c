// common.cuh -- shared device helpers for liblongbow_b200 (sm_100a only).
//
// * exact arithmetic: bit-for-bit the reference's portable Go kernels (Unrolled4x lane order,
//   separate mul/add roundings, sqrt through double) -- used wherever a value is RETURNED.
// * (key, id) packing: one u64 compare orders by (distance, id), which is the tie rule the
//   reference's brute-force heap implies (internal/store/adaptive_index.go:200-211).
// * warp / block bitonic sorts over shared memory.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace lb {

enum { METRIC_L2 = 0, METRIC_COSINE = 1, METRIC_DOT = 2 };
enum { DT_F32 = 0, DT_F16 = 1, DT_I8 = 2, DT_U8 = 3 };

constexpr uint64_t kInvalid = ~0ull;

// ---------------------------------------------------------------------------------------------
// Ordered key packing.  ord(f) is monotone in f over all non-NaN floats (negatives included:
// the dot metric is a negated similarity).
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t float_to_ordered(float f) {
#ifdef __CUDA_ARCH__
    uint32_t b = __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; uint32_t b = c.u;
#endif
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_float(uint32_t o) {
    uint32_t b = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    union { float f; uint32_t u; } c; c.u = b; return c.f;
#endif
}
__host__ __device__ __forceinline__ uint64_t pack_key(float key, uint32_t id) {
    return ((uint64_t)float_to_ordered(key) << 32) | id;
}
__host__ __device__ __forceinline__ float key_of(uint64_t p) { return ordered_to_float((uint32_t)(p >> 32)); }
__host__ __device__ __forceinline__ uint32_t id_of(uint64_t p) { return (uint32_t)p; }

__device__ __forceinline__ bool bit_set(const uint32_t* __restrict__ bm, uint32_t i) {
    return (__ldg(bm + (i >> 5)) >> (i & 31)) & 1u;
}

// ---------------------------------------------------------------------------------------------
// Element widening (exact for every supported type).
// ---------------------------------------------------------------------------------------------
template <typename T> struct Elem;
template <> struct Elem<float> {
    static constexpr int kDtype = DT_F32;
    static constexpr int kVec = 4;  // elements per 16-byte vector
    __device__ static __forceinline__ float widen(float v) { return v; }
};
template <> struct Elem<__half> {
    static constexpr int kDtype = DT_F16;
    static constexpr int kVec = 8;
    __device__ static __forceinline__ float widen(__half v) { return __half2float(v); }
};
template <> struct Elem<int8_t> {
    static constexpr int kDtype = DT_I8;
    static constexpr int kVec = 16;
    __device__ static __forceinline__ float widen(int8_t v) { return (float)v; }
};
template <> struct Elem<uint8_t> {
    static constexpr int kDtype = DT_U8;
    static constexpr int kVec = 16;
    __device__ static __forceinline__ float widen(uint8_t v) { return (float)v; }
};

// Unpack one 16-byte vector into kVec floats.
template <typename T> __device__ __forceinline__ void unpack16(const uint4& v, float* out);
template <> __device__ __forceinline__ void unpack16<float>(const uint4& v, float* out) {
    out[0] = __uint_as_float(v.x); out[1] = __uint_as_float(v.y);
    out[2] = __uint_as_float(v.z); out[3] = __uint_as_float(v.w);
}
template <> __device__ __forceinline__ void unpack16<__half>(const uint4& v, float* out) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
        float2 f = __half22float2(h);
        out[2 * i] = f.x; out[2 * i + 1] = f.y;
    }
}
template <> __device__ __forceinline__ void unpack16<int8_t>(const uint4& v, float* out) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int b = 0; b < 4; b++) out[4 * i + b] = (float)(int8_t)((w[i] >> (8 * b)) & 0xff);
}
template <> __device__ __forceinline__ void unpack16<uint8_t>(const uint4& v, float* out) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int b = 0; b < 4; b++) out[4 * i + b] = (float)((w[i] >> (8 * b)) & 0xff);
}

// ---------------------------------------------------------------------------------------------
// Exact (reference-order) accumulation.  State for one (query,row) pair.
//   L2:      s[l] += (a-b)*(a-b)           internal/simd/simd.go:365-396, :767-791, simd_baseline.go:13-35
//   DOT:     s[l] += a*b                    simd.go:453-479, :793-809, simd_baseline.go:37-54
//   COSINE:  dot[l], na[l], nb[l]           simd.go:399-450, :811-848
// Elements i = 4m+l go to lane l in increasing m; the remainder (dim % 4) goes to lane 0.
// __fmul_rn/__fadd_rn/__fsub_rn are never contracted into FMA.
// ---------------------------------------------------------------------------------------------
template <int METRIC> struct ExactAcc;

template <> struct ExactAcc<METRIC_L2> {
    float s[4];
    __device__ __forceinline__ void init() { s[0] = s[1] = s[2] = s[3] = 0.f; }
    __device__ __forceinline__ void add(int lane, float a, float b) {
        float d = __fsub_rn(a, b);
        s[lane] = __fadd_rn(s[lane], __fmul_rn(d, d));
    }
    // squared distance (L2SquaredFloat32, distance_functions.go:195-227)
    __device__ __forceinline__ float sum() const {
        return __fadd_rn(__fadd_rn(__fadd_rn(s[0], s[1]), s[2]), s[3]);
    }
    // float32(math.Sqrt(float64(sum))) == correctly rounded single sqrt (53 >= 2*24+2 bits)
    __device__ __forceinline__ float finish() const { return __fsqrt_rn(sum()); }
};
template <> struct ExactAcc<METRIC_DOT> {
    float s[4];
    __device__ __forceinline__ void init() { s[0] = s[1] = s[2] = s[3] = 0.f; }
    __device__ __forceinline__ void add(int lane, float a, float b) {
        s[lane] = __fadd_rn(s[lane], __fmul_rn(a, b));
    }
    // raw similarity; callers negate it when it is used as a distance
    __device__ __forceinline__ float finish() const {
        return __fadd_rn(__fadd_rn(__fadd_rn(s[0], s[1]), s[2]), s[3]);
    }
};
template <> struct ExactAcc<METRIC_COSINE> {
    float d[4], na[4], nb[4];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int i = 0; i < 4; i++) d[i] = na[i] = nb[i] = 0.f;
    }
    __device__ __forceinline__ void add(int lane, float a, float b) {
        d[lane] = __fadd_rn(d[lane], __fmul_rn(a, b));
        na[lane] = __fadd_rn(na[lane], __fmul_rn(a, a));
        nb[lane] = __fadd_rn(nb[lane], __fmul_rn(b, b));
    }
    __device__ __forceinline__ float finish() const {
        float dot = __fadd_rn(__fadd_rn(__fadd_rn(d[0], d[1]), d[2]), d[3]);
        float a = __fadd_rn(__fadd_rn(__fadd_rn(na[0], na[1]), na[2]), na[3]);
        float b = __fadd_rn(__fadd_rn(__fadd_rn(nb[0], nb[1]), nb[2]), nb[3]);
        if (a == 0.f || b == 0.f) return 1.0f;  // simd.go:446-448
        float den = (float)sqrt((double)a * (double)b);
        return __fsub_rn(1.0f, __fdiv_rn(dot, den));
    }
};

// Exact distance between a query held as floats (shared or global memory) and one row of T.
// `vec_ok` = row pointer 16-byte aligned and dim a multiple of Elem<T>::kVec.
template <typename T, int METRIC>
__device__ __forceinline__ float exact_pair(const float* __restrict__ q, const T* __restrict__ row, int dim,
                                            bool vec_ok) {
    ExactAcc<METRIC> acc;
    acc.init();
    constexpr int V = Elem<T>::kVec;
    int i = 0;
    if (vec_ok) {
        const uint4* rv = reinterpret_cast<const uint4*>(row);
        const int nv = dim / V;
        int v = 0;
        // eight 16-byte loads in flight per thread: a gathered row costs ~dim*elem/128 memory round
        // trips instead of one per vector (the accumulation order is unchanged)
        for (; v + 8 <= nv; v += 8) {
            uint4 raw[8];
#pragma unroll
            for (int u = 0; u < 8; u++) raw[u] = __ldg(rv + v + u);
#pragma unroll
            for (int u = 0; u < 8; u++) {
                float x[V];
                unpack16<T>(raw[u], x);
#pragma unroll
                for (int e = 0; e < V; e++) acc.add(e & 3, q[i + e], x[e]);
                i += V;
            }
        }
        for (; v < nv; v++) {
            uint4 raw = __ldg(rv + v);
            float x[V];
            unpack16<T>(raw, x);
#pragma unroll
            for (int e = 0; e < V; e++) acc.add(e & 3, q[i + e], x[e]);
            i += V;
        }
    } else {
        for (; i <= dim - 4; i += 4) {
#pragma unroll
            for (int e = 0; e < 4; e++) acc.add(e, q[i + e], Elem<T>::widen(row[i + e]));
        }
    }
    for (; i < dim; i++) acc.add(0, q[i], Elem<T>::widen(row[i]));
    return acc.finish();
}

// SQ8: squared L2 over uint8 as int32, returned as float32(int32) (internal/simd/sq8.go:45-66).
__device__ __forceinline__ float exact_sq8(const float* __restrict__ q, const uint8_t* __restrict__ row, int dim) {
    int32_t sum = 0;
    for (int i = 0; i < dim; i++) {
        int32_t d = (int32_t)q[i] - (int32_t)row[i];
        sum += d * d;
    }
    return (float)sum;
}

// ---------------------------------------------------------------------------------------------
// Bitonic sorts over u64 keys in shared memory (ascending).  n must be a power of two.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_bitonic_sort(uint64_t* a, int n, int lane) {
    for (int size = 2; size <= n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = lane; t < (n >> 1); t += 32) {
                const int lo = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));  // stride is a power of two
                const int hi = lo + stride;
                bool up = ((lo & size) == 0);
                uint64_t x = a[lo], y = a[hi];
                if ((x > y) == up) { a[lo] = y; a[hi] = x; }
            }
            __syncwarp();
        }
    }
}

template <typename K>
__device__ __forceinline__ void block_bitonic_sort_t(K* a, int n) {
    for (int size = 2; size <= n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
                const int lo = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));  // stride is a power of two
                const int hi = lo + stride;
                bool up = ((lo & size) == 0);
                K x = a[lo], y = a[hi];
                if ((x > y) == up) { a[lo] = y; a[hi] = x; }
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ void block_bitonic_sort(uint64_t* a, int n) {
    for (int size = 2; size <= n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
                const int lo = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));  // stride is a power of two
                const int hi = lo + stride;
                bool up = ((lo & size) == 0);
                uint64_t x = a[lo], y = a[hi];
                if ((x > y) == up) { a[lo] = y; a[hi] = x; }
            }
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Warp bitonic sort held entirely in registers: 32 lanes x R keys, blocked layout (element
// index = lane * R + r), ascending.  Strides < R are register-to-register compare-exchanges,
// strides >= R are one __shfl_xor per key.  ~10x faster than the shared-memory version; used by
// the tensor-core epilogue where list compaction sits on the critical path.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void ce_regs(uint64_t& a, uint64_t& b, bool asc) {
    const uint64_t lo = a < b ? a : b, hi = a < b ? b : a;
    a = asc ? lo : hi;
    b = asc ? hi : lo;
}

template <int R, int S>
__device__ __forceinline__ void sort_inreg_pass(uint64_t (&v)[R], int lane, int size) {
#pragma unroll
    for (int r = 0; r < R; r++) {
        if ((r & S) == 0) {
            const bool asc = (((lane * R + r) & size) == 0);
            ce_regs(v[r], v[r | S], asc);
        }
    }
}

template <int R>
__device__ __forceinline__ void warp_sort_regs(uint64_t (&v)[R], int lane) {
    constexpr int N = 32 * R;
#pragma unroll 1
    for (int size = 2; size <= N; size <<= 1) {
        // strides that cross lanes
#pragma unroll 1
        for (int s = size >> 1; s >= R; s >>= 1) {
            const int m = s / R;
            const bool lower = (lane & m) == 0;
            const bool asc = ((lane * R) & size) == 0;  // size > s >= R: the bit lives in the lane index
            const bool keep_min = (lower == asc);
#pragma unroll
            for (int r = 0; r < R; r++) {
                const uint64_t o = __shfl_xor_sync(0xffffffffu, v[r], m);
                const uint64_t lo = v[r] < o ? v[r] : o, hi = v[r] < o ? o : v[r];
                v[r] = keep_min ? lo : hi;
            }
        }
        // strides inside a lane's registers
        if (R > 16 && size > 16) sort_inreg_pass<R, (R > 16 ? 16 : 1)>(v, lane, size);
        if (R > 8 && size > 8) sort_inreg_pass<R, (R > 8 ? 8 : 1)>(v, lane, size);
        if (R > 4 && size > 4) sort_inreg_pass<R, (R > 4 ? 4 : 1)>(v, lane, size);
        if (R > 2 && size > 2) sort_inreg_pass<R, (R > 2 ? 2 : 1)>(v, lane, size);
        sort_inreg_pass<R, 1>(v, lane, size);
    }
}

// ---------------------------------------------------------------------------------------------
// Block-wide k-th smallest of packed (key,row) values held E per thread in registers.
// MSB-first bitwise search on the 32-bit key (skipping the prefix all keys share), then -- only if
// the k-th key is tied and not every tie is wanted -- on the row id.  Counting is a warp shuffle
// reduction plus one shared-memory exchange per bit; no atomics, so degenerate key distributions
// (all keys in one radix bin) cost the same as uniform ones.
// Returns kInvalid if fewer than k valid values exist.  s_red: shared int[64].  All threads call.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int block_sum_i(int x, int* s_red, int tid, int nwarps) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if ((tid & 31) == 0) s_red[tid >> 5] = x;
    __syncthreads();
    int t = 0;
    for (int w = 0; w < nwarps; w++) t += s_red[w];
    __syncthreads();
    return t;
}

template <int E>
__device__ __forceinline__ uint64_t block_kth_smallest(const uint64_t (&v)[E], int k, int* s_red, int tid, int nwarps) {
    int nv = 0;
    uint32_t a_and = 0xffffffffu, a_or = 0u;
#pragma unroll
    for (int e = 0; e < E; e++) {
        if (v[e] != kInvalid) { nv++; a_and &= (uint32_t)(v[e] >> 32); a_or |= (uint32_t)(v[e] >> 32); }
    }
    const int n_valid = block_sum_i(nv, s_red, tid, nwarps);
    if (n_valid < k) return kInvalid;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a_and &= __shfl_xor_sync(0xffffffffu, a_and, o);
        a_or |= __shfl_xor_sync(0xffffffffu, a_or, o);
    }
    if ((tid & 31) == 0) { s_red[tid >> 5] = (int)a_and; s_red[32 + (tid >> 5)] = (int)a_or; }
    __syncthreads();
    a_and = 0xffffffffu; a_or = 0u;
    for (int w = 0; w < nwarps; w++) { a_and &= (uint32_t)s_red[w]; a_or |= (uint32_t)s_red[32 + w]; }
    __syncthreads();
    const uint32_t diff = a_and ^ a_or;
    const int top = diff ? (31 - __clz(diff)) : -1;
    uint32_t T = (top < 0) ? a_and : (top >= 31 ? 0u : (a_and & ~((2u << top) - 1u)));
    for (int bit = top; bit >= 0; bit--) {
        const uint32_t test = T | (1u << bit);
        int c = 0;
#pragma unroll
        for (int e = 0; e < E; e++) c += (v[e] != kInvalid) && ((uint32_t)(v[e] >> 32) < test);
        if (block_sum_i(c, s_red, tid, nwarps) < k) T = test;
    }
    int c_less = 0, c_tie = 0;
#pragma unroll
    for (int e = 0; e < E; e++) {
        if (v[e] != kInvalid) {
            c_less += ((uint32_t)(v[e] >> 32) < T);
            c_tie += ((uint32_t)(v[e] >> 32) == T);
        }
    }
    const int n_less = block_sum_i(c_less, s_red, tid, nwarps);
    const int n_tie = block_sum_i(c_tie, s_red, tid, nwarps);
    const int need = k - n_less;  // 1..n_tie of the tied values, lowest row ids first
    if (need >= n_tie) return ((uint64_t)T << 32) | 0xffffffffu;
    uint32_t L = 0;
    for (int bit = 31; bit >= 0; bit--) {
        const uint32_t test = L | (1u << bit);
        int c = 0;
#pragma unroll
        for (int e = 0; e < E; e++)
            c += (v[e] != kInvalid) && ((uint32_t)(v[e] >> 32) == T) && ((uint32_t)v[e] < test);
        if (block_sum_i(c, s_red, tid, nwarps) < need) L = test;
    }
    return ((uint64_t)T << 32) | L;
}

// ---------------------------------------------------------------------------------------------
// Cooperative row gather (exact stage, re-rank, graph walk): see kernels_dense.cu S3.
// ---------------------------------------------------------------------------------------------
constexpr int RC_WARPS = 4;
constexpr int RC_CHUNK = 128;            // bytes of a row per stage
constexpr int RC_PITCH = RC_CHUNK + 16;  // shared-memory row pitch
constexpr int RC_NBUF = 3;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
                 "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// One trip of the cooperative gather: the warp's (up to ROWS) candidate rows are copied global -> shared in 128-byte
// chunks with NBUF chunk buffers in flight, and lane l accumulates candidate l's chunk out of shared memory in the
// reference order.  NBUF * ROWS == 96 (the staging space of one warp).
template <typename T, int ACC, int NBUF, int ROWS>
__device__ __forceinline__ void rc_gather(unsigned char* stage, const unsigned char* const (&src_row)[8],
                                          const bool (&src_ok)[8], size_t row_bytes, int n_chunks, int dim,
                                          const float* qf, bool ok, int lane, ExactAcc<ACC>& acc, int vlane = -1) {
    const int my_row = vlane < 0 ? lane : vlane;  // staging row this lane accumulates (default: its own lane index)
    constexpr int V = Elem<T>::kVec;
    constexpr int PIECES = RC_CHUNK / 16;
    auto issue = [&](int ch) {
        unsigned char* dstb = stage + (size_t)(ch % NBUF) * ROWS * RC_PITCH + (lane & 7) * 16;
        const size_t off = (size_t)ch * RC_CHUNK;
        const bool in_row = off + (size_t)(lane & 7) * 16 < row_bytes;  // tail chunk: pieces past the row end
#pragma unroll
        for (int i = 0; i < ROWS / 4; i++) {
            if (src_ok[i] && in_row) cp_async16(dstb + (size_t)(4 * i + (lane >> 3)) * RC_PITCH, src_row[i] + off);
        }
        cp_async_commit();
    };
#pragma unroll
    for (int p = 0; p < NBUF - 1; p++) {
        if (p < n_chunks) issue(p); else cp_async_commit();
    }
    for (int ch = 0; ch < n_chunks; ch++) {
        if (ch + NBUF - 1 < n_chunks) issue(ch + NBUF - 1); else cp_async_commit();
        cp_async_wait<NBUF - 1>();
        __syncwarp();
        if (ok) {
            const uint4* rowp = reinterpret_cast<const uint4*>(stage + (size_t)(ch % NBUF) * ROWS * RC_PITCH +
                                                               (size_t)my_row * RC_PITCH);
            const int e0 = ch * (RC_CHUNK / (int)sizeof(T));
#pragma unroll
            for (int p = 0; p < PIECES; p++) {
                const int i0 = e0 + p * V;
                if (i0 < dim) {  // dim % V == 0 on this path: a piece is inside the row entirely or not at all
                    float x[V];
                    unpack16<T>(rowp[p], x);
#pragma unroll
                    for (int e = 0; e < V; e++) acc.add(e & 3, qf[i0 + e], x[e]);
                }
            }
        }
        __syncwarp();
    }
}

// One of the four interleaved partial sums of ExactAcc (elements i = 4m + t, increasing m): what ONE lane of a quad
// accumulates when four lanes share a (query, row) pair.
template <int METRIC> struct QuadPart;
template <> struct QuadPart<METRIC_L2> {
    float s;
    __device__ __forceinline__ void init() { s = 0.f; }
    __device__ __forceinline__ void add(float a, float b) {
        const float d = __fsub_rn(a, b);
        s = __fadd_rn(s, __fmul_rn(d, d));
    }
    // gather the quad's four partial sums (lanes 4c .. 4c+3) into the reference accumulator
    __device__ __forceinline__ void collect(ExactAcc<METRIC_L2>& acc, int lane) const {
#pragma unroll
        for (int j = 0; j < 4; j++) acc.s[j] = __shfl_sync(0xffffffffu, s, (lane & ~3) + j);
    }
};
template <> struct QuadPart<METRIC_DOT> {
    float s;
    __device__ __forceinline__ void init() { s = 0.f; }
    __device__ __forceinline__ void add(float a, float b) { s = __fadd_rn(s, __fmul_rn(a, b)); }
    __device__ __forceinline__ void collect(ExactAcc<METRIC_DOT>& acc, int lane) const {
#pragma unroll
        for (int j = 0; j < 4; j++) acc.s[j] = __shfl_sync(0xffffffffu, s, (lane & ~3) + j);
    }
};
template <> struct QuadPart<METRIC_COSINE> {
    float d, na, nb;
    __device__ __forceinline__ void init() { d = na = nb = 0.f; }
    __device__ __forceinline__ void add(float a, float b) {
        d = __fadd_rn(d, __fmul_rn(a, b));
        na = __fadd_rn(na, __fmul_rn(a, a));
        nb = __fadd_rn(nb, __fmul_rn(b, b));
    }
    __device__ __forceinline__ void collect(ExactAcc<METRIC_COSINE>& acc, int lane) const {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            acc.d[j] = __shfl_sync(0xffffffffu, d, (lane & ~3) + j);
            acc.na[j] = __shfl_sync(0xffffffffu, na, (lane & ~3) + j);
            acc.nb[j] = __shfl_sync(0xffffffffu, nb, (lane & ~3) + j);
        }
    }
};

// Cooperative gather, QUAD form (fp32 rows): up to 8 rows per trip, lane 4c + t accumulates partial sum t of row c --
// the reference's four interleaved sums, one per lane, so all 32 lanes work when a trip holds few rows (a graph
// expansion brings ~7 new nodes on average: one lane per row left 24 lanes idle for 384 dependent steps).  Same
// staging and cp.async pipeline as rc_gather<.., 8>; `part` ends up with this lane's partial sum.
template <int ACC, int NBUF>
__device__ __forceinline__ void rc_gather_quad(unsigned char* stage, const unsigned char* const (&src_row)[8],
                                               const bool (&src_ok)[8], size_t row_bytes, int n_chunks, int dim,
                                               const float* qf, bool ok, int lane, QuadPart<ACC>& part) {
    constexpr int ROWS = 8;
    constexpr int PIECES = RC_CHUNK / 16;
    const int c = lane >> 2, t = lane & 3;
    auto issue = [&](int ch) {
        unsigned char* dstb = stage + (size_t)(ch % NBUF) * ROWS * RC_PITCH + (lane & 7) * 16;
        const size_t off = (size_t)ch * RC_CHUNK;
        const bool in_row = off + (size_t)(lane & 7) * 16 < row_bytes;
#pragma unroll
        for (int i = 0; i < ROWS / 4; i++) {
            if (src_ok[i] && in_row) cp_async16(dstb + (size_t)(4 * i + (lane >> 3)) * RC_PITCH, src_row[i] + off);
        }
        cp_async_commit();
    };
#pragma unroll
    for (int p = 0; p < NBUF - 1; p++) {
        if (p < n_chunks) issue(p); else cp_async_commit();
    }
    for (int ch = 0; ch < n_chunks; ch++) {
        if (ch + NBUF - 1 < n_chunks) issue(ch + NBUF - 1); else cp_async_commit();
        cp_async_wait<NBUF - 1>();
        __syncwarp();
        if (ok) {
            const float* rowp = reinterpret_cast<const float*>(stage + (size_t)(ch % NBUF) * ROWS * RC_PITCH +
                                                               (size_t)c * RC_PITCH) + t;
            const int e0 = ch * (RC_CHUNK / 4) + t;
            const float* qp = qf + e0;
            if ((ch + 1) * (RC_CHUNK / 4) <= dim) {   // chunk entirely inside the row: no per-piece test
#pragma unroll
                for (int p = 0; p < PIECES; p++) part.add(qp[4 * p], rowp[4 * p]);
            } else {
#pragma unroll
                for (int p = 0; p < PIECES; p++) {
                    if (e0 + 4 * p < dim) part.add(qp[4 * p], rowp[4 * p]);
                }
            }
        }
        __syncwarp();
    }
}

__host__ __device__ __forceinline__ int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

}  // namespace lb
