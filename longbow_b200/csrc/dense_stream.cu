// dense_stream.cu -- S1 coarse scan for SMALL query batches (1..8 queries): the HBM-bound case.
//
// With a handful of queries the contraction has no reuse to offer the tensor cores (AI = 2*nq/elem flop/B,
// far below the ridge), so the right kernel is a streaming one: read every database row exactly once with
// coalesced 128-bit loads, keep the queries on chip, and spend as few instructions per byte as possible.
// This is gpu.Index.Search as the reference calls it (one query per call, internal/gpu/faiss_gpu.go:125-131)
// and the "scan GB/s vs HBM roofline" half of the headline metric.
//
//   * a group of LPR lanes owns one row: lane s of the group loads 16-byte pieces s, s+LPR, ... of it
//     (LPR = 32 for rows >= 512 B, else 8), so one warp instruction reads 512 contiguous bytes;
//   * U row groups are in flight per warp (U loads per lane before any is consumed);
//   * queries sit in shared memory (fp32 for float rows, packed int8 + dp4a for int8 rows); a piece of a
//     query is read once and applied to all U rows in flight;
//   * partial dots are reduced across the LPR lanes with shuffles once per row, the group's first lane turns
//     them into ranking keys and applies the same threshold filter / candidate lists / shared progressive
//     threshold ladder as the tensor-core scan (dense_tc.cu), with lists and counters in shared memory.
// Output: partial[cta][q][kc] packed (key,row); merge + exact re-score follow (kernels_dense.cu).
#include "kernels.cuh"

namespace lb {

constexpr int ST_THREADS = 256;
constexpr int ST_WARPS = ST_THREADS / 32;
constexpr int ST_MAXQ = 8;

template <typename T> struct StreamQ;   // how a query piece multiplies a row piece
template <> struct StreamQ<__half> {
    using QT = float;                    // query element type in shared memory
    static constexpr int kV = 8;
    // q points at the piece's first float4; the second one lives `plane` floats further (two planes, so that the
    // 32 lanes of a warp read 32 consecutive float4: no bank conflicts -- the interleaved layout was 2-way)
    __device__ static __forceinline__ float dot16(const uint4& row, const float* q, int plane) {
        float x[8];
        unpack16<__half>(row, x);
        const float4 a = *reinterpret_cast<const float4*>(q), b = *reinterpret_cast<const float4*>(q + plane);
        float s = x[0] * a.x;
        s = fmaf(x[1], a.y, s); s = fmaf(x[2], a.z, s); s = fmaf(x[3], a.w, s);
        s = fmaf(x[4], b.x, s); s = fmaf(x[5], b.y, s); s = fmaf(x[6], b.z, s); s = fmaf(x[7], b.w, s);
        return s;
    }
};
template <> struct StreamQ<float> {
    using QT = float;
    static constexpr int kV = 4;
    __device__ static __forceinline__ float dot16(const uint4& row, const float* q, int) {
        const float4 a = *reinterpret_cast<const float4*>(q);
        float s = __uint_as_float(row.x) * a.x;
        s = fmaf(__uint_as_float(row.y), a.y, s);
        s = fmaf(__uint_as_float(row.z), a.z, s);
        s = fmaf(__uint_as_float(row.w), a.w, s);
        return s;
    }
};
template <> struct StreamQ<int8_t> {
    using QT = int8_t;
    static constexpr int kV = 16;
    __device__ static __forceinline__ float dot16(const uint4& row, const int8_t* q, int) {
        const uint4 a = *reinterpret_cast<const uint4*>(q);
        int s = __dp4a((int)row.x, (int)a.x, 0);
        s = __dp4a((int)row.y, (int)a.y, s);
        s = __dp4a((int)row.z, (int)a.z, s);
        s = __dp4a((int)row.w, (int)a.w, s);
        return (float)s;  // exact: |s| < 2^24 for 16 int8 products; the row total is summed in fp32 below
    }
};

struct StreamArgs {
    const void* db;
    const float* aux;
    uint32_t n_rows;
    int dim;
    const void* queries;
    int nq;
    const uint32_t* tomb;
    uint32_t tomb_bits;
    const uint32_t* allow;
    int kc, cap;
    uint64_t* partial;        // compact per-query output [nq][out_stride]; out_cnt[q] entries are valid
    uint32_t* out_cnt;        // [nq] running entry counts (zeroed by the caller)
    size_t out_stride;
    uint32_t row_begin;       // rows [row_begin, n_rows) are scanned
    const float* tau_init;    // [nq] or null
    const float* edges;       // [nq][LB_NEDGE] or null
    uint32_t* edge_cnt;       // [nq][LB_NEDGE]
    float* keys_out;          // dump mode: [nq][keys_ld] keys of rows [0, row_end_dump)
    int keys_ld;
    uint32_t dump_rows;
};

// int8 totals: per-piece integer dots are exact; summing <= 2^12 pieces of |value| < 2^22 in fp32 is exact only
// while the total stays below 2^24, which holds for dim <= 1024 (|dot| <= dim * 2^14).  Larger int8 rows
// still rank correctly up to fp32 rounding and are re-scored exactly afterwards.
template <typename T, int METRIC, int NQ, int LPR, int U>
__global__ void __launch_bounds__(ST_THREADS, (NQ == 1 ? 4 : 3))
dense_scan_stream(const StreamArgs a) {
    using SQ = StreamQ<T>;
    using QT = typename SQ::QT;
    constexpr int V = SQ::kV;
    constexpr int RPW = 32 / LPR;                       // rows per warp instruction
    constexpr int ROWS_PER_ITER = ST_WARPS * U * RPW;   // rows one CTA consumes per loop trip
    constexpr int CHK = (512 / ROWS_PER_ITER) > 0 ? (512 / ROWS_PER_ITER) : 1;
    static_assert(CHK * ROWS_PER_ITER <= 512, "candidate lists are sized for 512 rows between overflow checks");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    QT* qs = reinterpret_cast<QT*>(smem_raw);                                   // [NQ][dim]
    const size_t q_bytes = ((size_t)NQ * a.dim * sizeof(QT) + 15) & ~(size_t)15;
    uint64_t* lists = reinterpret_cast<uint64_t*>(smem_raw + q_bytes);          // [NQ][cap]
    float* s_edges = reinterpret_cast<float*>(lists + (size_t)NQ * a.cap);      // [NQ][LB_NEDGE]
    __shared__ float s_tau[ST_MAXQ];
    __shared__ int s_cnt[ST_MAXQ];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sub = lane % LPR, grp = lane / LPR;
    const int pieces = (int)((size_t)a.dim * sizeof(T) / 16);
    const bool dump = a.keys_out != nullptr;
    const bool shared_tau = a.edges != nullptr;

    // queries -> shared memory (widened to fp32 for float rows); queries past nq are zero
    for (int i = tid; i < NQ * a.dim; i += ST_THREADS) {
        const int q = i / a.dim, e = i - q * a.dim;
        QT v = (QT)0;
        if (q < a.nq) {
            if constexpr (sizeof(T) == 1) v = reinterpret_cast<const int8_t*>(a.queries)[(size_t)q * a.dim + e];
            else v = Elem<T>::widen(reinterpret_cast<const T*>(a.queries)[(size_t)q * a.dim + e]);
        }
        int di = e;
        if constexpr (sizeof(T) == 2) {  // fp16 rows: two planes of float4 per query (see StreamQ<__half>::dot16)
            const int pc = e >> 3, w = e & 7;
            di = (w < 4) ? (pc * 4 + w) : ((a.dim >> 1) + pc * 4 + (w - 4));
        }
        qs[(size_t)q * a.dim + di] = v;
    }
    if (tid < NQ) {
        s_cnt[tid] = 0;
        s_tau[tid] = (tid < a.nq) ? (a.tau_init ? a.tau_init[tid] : INFINITY) : -INFINITY;
    }
    for (int i = tid; i < NQ * LB_NEDGE; i += ST_THREADS)
        s_edges[i] = (shared_tau && i / LB_NEDGE < a.nq) ? a.edges[i] : INFINITY;
    __syncthreads();

    const uint32_t n_rows = dump ? a.dump_rows : a.n_rows;
    const uint4* dbv = reinterpret_cast<const uint4*>(a.db);
    uint32_t iter = 0;
    for (uint32_t base = a.row_begin + blockIdx.x * ROWS_PER_ITER; base < n_rows;
         base += gridDim.x * ROWS_PER_ITER, iter++) {
        // shared-threshold counters: fetched now (L2), folded into s_tau after this trip's rows are done
        // list-overflow check and threshold refresh need block barriers: only every CHK trips (<= 512 rows)
        const bool chk = ((iter + 1) % CHK) == 0;  // block-uniform
        const bool refresh = shared_tau && !dump && chk && tid < a.nq;
        uint4 gc[LB_NEDGE / 4];
        if (refresh) {
#pragma unroll
            for (int i = 0; i < LB_NEDGE / 4; i++)
                gc[i] = __ldcg(reinterpret_cast<const uint4*>(a.edge_cnt + (size_t)tid * LB_NEDGE) + i);
        }
        const uint32_t wrow = base + warp * (U * RPW) + grp;   // this lane's row in row group u: wrow + u * RPW
        float acc[U][NQ];
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
            for (int q = 0; q < NQ; q++) acc[u][q] = 0.f;
        for (int p = sub; p < pieces; p += LPR) {
            uint4 raw[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const uint32_t r = wrow + u * RPW;
                raw[u] = make_uint4(0, 0, 0, 0);
                if (r < n_rows) raw[u] = __ldcs(dbv + (size_t)r * pieces + p);  // streamed once: evict first
            }
#pragma unroll
            for (int q = 0; q < NQ; q++) {
                const QT* qp = qs + (size_t)q * a.dim + (size_t)p * (sizeof(T) == 2 ? 4 : V);
#pragma unroll
                for (int u = 0; u < U; u++) acc[u][q] += SQ::dot16(raw[u], qp, a.dim >> 1);
            }
        }
        // reduce the partial dots across the LPR lanes of each row
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1)
#pragma unroll
            for (int u = 0; u < U; u++)
#pragma unroll
                for (int q = 0; q < NQ; q++) acc[u][q] += __shfl_xor_sync(0xffffffffu, acc[u][q], o);
        if (sub == 0) {
#pragma unroll
            for (int u = 0; u < U; u++) {
                const uint32_t r = wrow + u * RPW;
                if (r >= n_rows) continue;
                float ax = 0.f;
                if constexpr (METRIC != METRIC_DOT) ax = __ldg(a.aux + r);
#pragma unroll
                for (int q = 0; q < NQ; q++) {
                    float key;
                    if constexpr (METRIC == METRIC_L2) key = fmaf(-2.f, acc[u][q], ax);
                    else if constexpr (METRIC == METRIC_COSINE) key = -acc[u][q] * ax;
                    else key = -acc[u][q];
                    if (dump) {
                        if (q < a.nq) a.keys_out[(size_t)q * a.keys_ld + r] = key;
                        continue;
                    }
                    if (key < s_tau[q]) {
                        bool ok = true;
                        if (a.tomb != nullptr && r < a.tomb_bits && bit_set(a.tomb, r)) ok = false;
                        if (ok && a.allow != nullptr && !bit_set(a.allow, r)) ok = false;
                        if (ok) {
                            const int pos = atomicAdd(&s_cnt[q], 1);
                            if (pos < a.cap) lists[(size_t)q * a.cap + pos] = pack_key(key, r);
                            if (shared_tau) {
                                int b = 0;
#pragma unroll
                                for (int i = 0; i < LB_NEDGE; i++) b += (s_edges[q * LB_NEDGE + i] <= key) ? 1 : 0;
                                if (b < LB_NEDGE) atomicAdd(a.edge_cnt + (size_t)q * LB_NEDGE + b, 1u);
                            }
                        }
                    }
                }
            }
        }
        if (dump || !chk) continue;
        __syncthreads();
        // compaction of any list that could overflow before the next check; shared-threshold refresh
        for (int q = 0; q < a.nq; q++) {
            const int c = min(s_cnt[q], a.cap);
            if (c > a.cap - CHK * ROWS_PER_ITER) {  // block-uniform
                uint64_t* buf = lists + (size_t)q * a.cap;
                const int n2 = next_pow2(c);
                for (int t = c + tid; t < n2; t += ST_THREADS) buf[t] = kInvalid;
                __syncthreads();
                block_bitonic_sort(buf, n2);
                if (tid == 0) {
                    s_cnt[q] = min(c, a.kc);
                    if (c >= a.kc) s_tau[q] = fminf(s_tau[q], nextafterf(key_of(buf[a.kc - 1]), INFINITY));
                }
                __syncthreads();
            }
        }
        if (refresh) {
            const uint32_t cv[LB_NEDGE] = {gc[0].x, gc[0].y, gc[0].z, gc[0].w, gc[1].x, gc[1].y, gc[1].z, gc[1].w,
                                           gc[2].x, gc[2].y, gc[2].z, gc[2].w, gc[3].x, gc[3].y, gc[3].z, gc[3].w};
            uint32_t cum = 0;
            float tg = INFINITY;
            bool found = false;
#pragma unroll
            for (int i = 0; i < LB_NEDGE; i++) {
                cum += cv[i];
                const bool hit = cum >= (uint32_t)a.kc;
                if (hit && !found) tg = s_edges[tid * LB_NEDGE + i];
                found = found || hit;
            }
            s_tau[tid] = fminf(s_tau[tid], tg);
        }
        __syncthreads();
    }
    if (dump) return;
    __syncthreads();
    // emit: this CTA's best kc per query, appended to the query's compact global list (one atomic per CTA)
    __shared__ uint32_t s_pos;
    for (int q = 0; q < a.nq; q++) {
        int c = min(s_cnt[q], a.cap);
        uint64_t* buf = lists + (size_t)q * a.cap;
        if (c > a.kc) {  // only then does the order decide what is kept
            const int n2 = next_pow2(c);
            for (int t = c + tid; t < n2; t += ST_THREADS) buf[t] = kInvalid;
            __syncthreads();
            block_bitonic_sort(buf, n2);
            c = a.kc;
        }
        if (tid == 0) s_pos = c ? atomicAdd(a.out_cnt + q, (uint32_t)c) : 0u;
        __syncthreads();
        uint64_t* out = a.partial + (size_t)q * a.out_stride + s_pos;
        for (int t = tid; t < c; t += ST_THREADS) out[t] = buf[t];
        __syncthreads();
    }
}

bool dense_stream_eligible(int dtype, int dim, const void* db, int nq, int kc) {
    const size_t elem = dtype == DT_F16 ? 2 : dtype == DT_I8 ? 1 : dtype == DT_F32 ? 4 : 0;
    if (elem == 0 || nq < 1 || nq > ST_MAXQ) return false;
    const size_t row_bytes = (size_t)dim * elem;
    if (row_bytes % 16 != 0) return false;  // whole 16-byte pieces
    if (reinterpret_cast<uintptr_t>(db) & 15) return false;
    if (kc > 256) return false;
    return true;
}

// One wave exactly: the grid-stride loop gives every CTA the same share of rows, so the grid must equal the number
// of CTAs that are resident at once (round 1 launched 4 per SM with 3 resident: a second, one-third-full wave
// cost 15 % of the kernel, ncu "SM Active Cycles" 498k of 584k).  Single-query kernels fit 4 per SM (64 registers).
int dense_stream_grid(int sm_count, int nq) { return (nq == 1 ? 4 : 3) * sm_count; }

template <typename T, int METRIC, int NQ>
static cudaError_t launch_stream_nq(const StreamArgs& a, int grid, cudaStream_t st) {
    const size_t row_bytes = (size_t)a.dim * sizeof(T);
    const size_t q_bytes = ((size_t)NQ * a.dim * sizeof(typename StreamQ<T>::QT) + 15) & ~(size_t)15;
    const size_t smem = q_bytes + (size_t)NQ * a.cap * 8 + (size_t)NQ * LB_NEDGE * 4;
#define LB_ST(LPR_, U_)                                                                                      \
    {                                                                                                        \
        auto kern = dense_scan_stream<T, METRIC, NQ, LPR_, U_>;                                              \
        LB_SMEM_OPTIN(kern);                                                                                 \
        kern<<<grid, ST_THREADS, smem, st>>>(a);                                                             \
    }
    if (row_bytes >= 384) LB_ST(32, (NQ == 1 ? 8 : NQ == 2 ? 4 : 2)) else LB_ST(8, (NQ <= 2 ? 8 : 4))
#undef LB_ST
    count_launch();
    return cudaGetLastError();
}

template <typename T, int METRIC>
static cudaError_t launch_stream_metric(const StreamArgs& a, int grid, cudaStream_t st) {
    if (a.nq <= 1) return launch_stream_nq<T, METRIC, 1>(a, grid, st);
    if (a.nq <= 2) return launch_stream_nq<T, METRIC, 2>(a, grid, st);
    if (a.nq <= 4) return launch_stream_nq<T, METRIC, 4>(a, grid, st);
    return launch_stream_nq<T, METRIC, 8>(a, grid, st);
}

template <typename T>
static cudaError_t launch_stream_dtype(const StreamArgs& a, int metric, int grid, cudaStream_t st) {
    switch (metric) {
        case METRIC_L2: return launch_stream_metric<T, METRIC_L2>(a, grid, st);
        case METRIC_COSINE: return launch_stream_metric<T, METRIC_COSINE>(a, grid, st);
        default: return launch_stream_metric<T, METRIC_DOT>(a, grid, st);
    }
}

// rows_per_iter upper bound used to size the candidate lists (cap >= kc + rows one CTA adds per trip)
int dense_stream_cap(int kc) { return next_pow2(kc + 512); }  // LPR=8, U=16: 512 rows per trip

cudaError_t launch_dense_scan_stream(const ScanArgs& s, int grid, uint32_t row_begin, uint32_t dump_rows,
                                     uint32_t* out_cnt, size_t out_stride, cudaStream_t st) {
    StreamArgs a;
    a.db = s.db; a.aux = s.aux; a.n_rows = s.n_rows; a.dim = s.dim; a.queries = s.queries; a.nq = s.nq;
    a.tomb = s.tomb; a.tomb_bits = s.tomb_bits; a.allow = s.allow;
    a.kc = s.kc; a.cap = dense_stream_cap(s.kc);
    a.partial = s.partial; a.out_cnt = out_cnt; a.out_stride = out_stride;
    a.row_begin = row_begin; a.tau_init = s.tau_init;
    a.edges = s.edges; a.edge_cnt = s.edge_cnt;
    a.keys_out = s.keys_out; a.keys_ld = s.keys_ld; a.dump_rows = dump_rows;
    switch (s.dtype) {
        case DT_F32: return launch_stream_dtype<float>(a, s.metric, grid, st);
        case DT_F16: return launch_stream_dtype<__half>(a, s.metric, grid, st);
        case DT_I8:
            if (s.metric == METRIC_COSINE) return cudaErrorInvalidValue;
            return launch_stream_dtype<int8_t>(a, s.metric, grid, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace lb
