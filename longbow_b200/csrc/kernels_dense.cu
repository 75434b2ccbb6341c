// kernels_dense.cu -- dense (uncompressed) distance + fused top-k kernels, SIMT path.
//
// Stages of a dense search:
//   S1 coarse scan   : tile of queries x stream of DB rows -> approximate ranking key per pair,
//                      fused per-query top-kc selection in shared memory (no distance matrix
//                      in HBM) -> partial[part][q][kc] packed (key,id)
//   S2 merge         : per query, parts*kc -> kc candidates
//   S3 exact re-score: candidates re-evaluated with the reference's own arithmetic
//                      (common.cuh ExactAcc), sorted by (distance,id) -> top-k
// S1 here is the general SIMT kernel (any dim, any dtype, unaligned rows).  The tcgen05
// kernels in dense_tc.cu produce the same partial[] format for the aligned fp16 / int8 cases.
#include "kernels.cuh"

namespace lb {

bool g_rescore_legacy = false;
bool g_rescore_block = false;   // lb_set_option("rescore_block"): A/B the block-per-query cooperative kernel  // lb_set_option("rescore_legacy"): A/B the thread-per-candidate kernel

// ---------------------------------------------------------------------------------------------
// Row auxiliaries used by coarse keys: aux[r] = |x_r|^2 (L2) or 1/|x_r| (cosine; 0 for a
// zero row, which makes the coarse key 0 == cosine distance 1.0, simd.go:446-448).
// One warp per row.  Accuracy only affects the coarse ranking, never a returned value.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void row_aux_kernel(const T* __restrict__ db, int64_t n, int dim, int metric,
                               float* __restrict__ aux, int64_t row0) {
    int64_t r = row0 + (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (r >= n) return;
    const T* row = db + r * dim;
    float s = 0.f;
    for (int i = lane; i < dim; i += 32) {
        float v = Elem<T>::widen(row[i]);
        s = fmaf(v, v, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) aux[r] = (metric == METRIC_COSINE) ? (s > 0.f ? rsqrtf(s) : 0.f) : s;
}

// Exact |x_r|^2 in the reference's 4-lane order (simd.go:399-450 accumulates normB this way inside
// the cosine loop).  Stored once per row at add time so the cosine re-score only has to run the dot
// chain; the value is bit-identical to what the reference recomputes for every pair.
template <typename T>
__global__ void row_norm_exact_kernel(const T* __restrict__ db, int64_t n, int dim, float* __restrict__ nrm,
                                      int64_t row0) {
    int64_t r = row0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const T* row = db + r * dim;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int i = 0;
    for (; i <= dim - 4; i += 4) {
        const float x0 = Elem<T>::widen(row[i]), x1 = Elem<T>::widen(row[i + 1]);
        const float x2 = Elem<T>::widen(row[i + 2]), x3 = Elem<T>::widen(row[i + 3]);
        s0 = __fadd_rn(s0, __fmul_rn(x0, x0)); s1 = __fadd_rn(s1, __fmul_rn(x1, x1));
        s2 = __fadd_rn(s2, __fmul_rn(x2, x2)); s3 = __fadd_rn(s3, __fmul_rn(x3, x3));
    }
    for (; i < dim; i++) { const float x = Elem<T>::widen(row[i]); s0 = __fadd_rn(s0, __fmul_rn(x, x)); }
    nrm[r] = __fadd_rn(__fadd_rn(__fadd_rn(s0, s1), s2), s3);
}

cudaError_t launch_row_norm_exact(int dtype, const void* db, int64_t n, int dim, float* nrm, int64_t row0,
                                  cudaStream_t st) {
    if (n <= row0) return cudaSuccess;
    unsigned blocks = (unsigned)((n - row0 + 127) / 128);
    switch (dtype) {
        case DT_F32: row_norm_exact_kernel<float><<<blocks, 128, 0, st>>>((const float*)db, n, dim, nrm, row0); break;
        case DT_F16: row_norm_exact_kernel<__half><<<blocks, 128, 0, st>>>((const __half*)db, n, dim, nrm, row0); break;
        case DT_I8: row_norm_exact_kernel<int8_t><<<blocks, 128, 0, st>>>((const int8_t*)db, n, dim, nrm, row0); break;
        default: return cudaErrorInvalidValue;
    }
    count_launch();
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// S1: SIMT coarse scan with fused selection.
//   grid  = (parts, ceil(nq / TQ)); block = 256 threads
//   tile  = TQ queries x TN rows, K chunks of KC elements staged (widened to fp32) in smem
//   thread micro-tile = (TQ/16) queries x 8 rows
// ---------------------------------------------------------------------------------------------
constexpr int TN = 128;
constexpr int KC = 32;
constexpr int NT = 256;

template <typename T, int METRIC, int TQ>
__global__ void __launch_bounds__(NT)
dense_scan_simt(const T* __restrict__ db, const float* __restrict__ aux, uint32_t n_rows, int dim,
                const T* __restrict__ queries, int nq, const uint32_t* __restrict__ tomb,
                uint32_t tomb_bits, const uint32_t* __restrict__ allow, int kc, int cap,
                uint32_t rows_per_part, uint64_t* __restrict__ partial) {
    constexpr int MQ = TQ / 16;  // queries per thread
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* As = reinterpret_cast<float*>(smem_raw);  // [KC][TQ]
    float* Bs = As + KC * TQ;                        // [KC][TN]
    float* tau = Bs + KC * TN;                       // [TQ]
    int* cnt = reinterpret_cast<int*>(tau + TQ);     // [TQ]
    uint64_t* cand = reinterpret_cast<uint64_t*>(cnt + TQ);  // [TQ][cap]

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.y * TQ;
    const uint32_t row_begin = blockIdx.x * rows_per_part;
    const uint32_t row_end = min(n_rows, row_begin + rows_per_part);

    for (int i = tid; i < TQ; i += NT) { tau[i] = INFINITY; cnt[i] = 0; }

    constexpr int V = Elem<T>::kVec;
    const bool vec_ok = (dim % V == 0) && ((reinterpret_cast<uintptr_t>(db) & 15) == 0) &&
                        ((reinterpret_cast<uintptr_t>(queries) & 15) == 0);
    __syncthreads();

    for (uint32_t r0 = row_begin; r0 < row_end; r0 += TN) {
        float acc[MQ][8];
#pragma unroll
        for (int i = 0; i < MQ; i++)
#pragma unroll
            for (int j = 0; j < 8; j++) acc[i][j] = 0.f;

        for (int k0 = 0; k0 < dim; k0 += KC) {
            // ---- stage A (queries) and B (rows) chunks, widened, transposed to [kk][row]
            if (vec_ok) {
                constexpr int VPC = KC / V;  // 16-byte vectors per row chunk
                for (int idx = tid; idx < TQ * VPC; idx += NT) {
                    int q = idx % TQ, kv = idx / TQ;
                    int kk = kv * V;
                    float x[V];
                    if (q0 + q < nq && k0 + kk < dim) {
                        uint4 raw = __ldg(reinterpret_cast<const uint4*>(queries + (size_t)(q0 + q) * dim + k0 + kk));
                        unpack16<T>(raw, x);
                    } else {
#pragma unroll
                        for (int e = 0; e < V; e++) x[e] = 0.f;
                    }
#pragma unroll
                    for (int e = 0; e < V; e++) As[(kk + e) * TQ + q] = x[e];
                }
                for (int idx = tid; idx < TN * VPC; idx += NT) {
                    int r = idx % TN, kv = idx / TN;
                    int kk = kv * V;
                    float x[V];
                    if (r0 + r < row_end && k0 + kk < dim) {
                        uint4 raw = __ldg(reinterpret_cast<const uint4*>(db + (size_t)(r0 + r) * dim + k0 + kk));
                        unpack16<T>(raw, x);
                    } else {
#pragma unroll
                        for (int e = 0; e < V; e++) x[e] = 0.f;
                    }
#pragma unroll
                    for (int e = 0; e < V; e++) Bs[(kk + e) * TN + r] = x[e];
                }
            } else {
                for (int idx = tid; idx < TQ * KC; idx += NT) {
                    int kk = idx % KC, q = idx / KC;
                    float v = 0.f;
                    if (q0 + q < nq && k0 + kk < dim) v = Elem<T>::widen(queries[(size_t)(q0 + q) * dim + k0 + kk]);
                    As[kk * TQ + q] = v;
                }
                for (int idx = tid; idx < TN * KC; idx += NT) {
                    int kk = idx % KC, r = idx / KC;
                    float v = 0.f;
                    if (r0 + r < row_end && k0 + kk < dim) v = Elem<T>::widen(db[(size_t)(r0 + r) * dim + k0 + kk]);
                    Bs[kk * TN + r] = v;
                }
            }
            __syncthreads();
#pragma unroll 8
            for (int kk = 0; kk < KC; kk++) {
                float a[MQ], b[8];
                if constexpr (MQ == 4) {
                    float4 av = *reinterpret_cast<const float4*>(&As[kk * TQ + ty * 4]);
                    a[0] = av.x; a[1] = av.y; a[2] = av.z; a[3] = av.w;
                } else {
#pragma unroll
                    for (int i = 0; i < MQ; i++) a[i] = As[kk * TQ + ty * MQ + i];
                }
                float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk * TN + tx * 4]);
                float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk * TN + 64 + tx * 4]);
                b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
                b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
                for (int i = 0; i < MQ; i++)
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        if constexpr (METRIC == METRIC_L2) {
                            float d = a[i] - b[j];  // difference form: no cancellation near d = 0
                            acc[i][j] = fmaf(d, d, acc[i][j]);
                        } else {
                            acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
                        }
                    }
            }
            __syncthreads();
        }

        // ---- fused selection: threshold filter, rare append to the query's candidate buffer
#pragma unroll
        for (int i = 0; i < MQ; i++) {
            const int ql = ty * MQ + i;
            if (q0 + ql >= nq) continue;
            const float t = tau[ql];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const uint32_t row = r0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
                if (row >= row_end) continue;
                float key;
                if constexpr (METRIC == METRIC_L2) key = acc[i][j];
                else if constexpr (METRIC == METRIC_COSINE) key = -acc[i][j] * __ldg(aux + row);
                else key = -acc[i][j];
                if (key < t) {
                    if (tomb != nullptr && row < tomb_bits && bit_set(tomb, row)) continue;
                    if (allow != nullptr && !bit_set(allow, row)) continue;
                    int pos = atomicAdd(&cnt[ql], 1);
                    if (pos < cap) cand[(size_t)ql * cap + pos] = pack_key(key, row);
                }
            }
        }
        __syncthreads();
        // ---- compaction: keep the kc best, tighten the threshold
        for (int ql = warp; ql < TQ; ql += NT / 32) {
            int c = min(cnt[ql], cap);
            if (c > cap - TN) {
                uint64_t* buf = cand + (size_t)ql * cap;
                int n2 = next_pow2(c);
                for (int t2 = c + lane; t2 < n2; t2 += 32) buf[t2] = kInvalid;
                __syncwarp();
                warp_bitonic_sort(buf, n2, lane);
                if (lane == 0) {
                    int keep = min(c, kc);
                    cnt[ql] = keep;
                    tau[ql] = (c >= kc) ? key_of(buf[kc - 1]) : INFINITY;
                }
            }
        }
        __syncthreads();
    }

    // ---- final compaction + write-out
    for (int ql = warp; ql < TQ; ql += NT / 32) {
        if (q0 + ql >= nq) continue;
        int c = min(cnt[ql], cap);
        uint64_t* buf = cand + (size_t)ql * cap;
        int n2 = next_pow2(max(c, 2));
        for (int t2 = c + lane; t2 < n2; t2 += 32) buf[t2] = kInvalid;
        __syncwarp();
        warp_bitonic_sort(buf, n2, lane);
        uint64_t* out = partial + ((size_t)blockIdx.x * nq + (q0 + ql)) * kc;
        for (int t2 = lane; t2 < kc; t2 += 32) out[t2] = (t2 < c) ? buf[t2] : kInvalid;
    }
}

size_t dense_scan_simt_smem(int tq, int cap) {
    return (size_t)(KC * tq + KC * TN + 2 * tq) * 4 + (size_t)tq * cap * 8;
}

template <typename T, int METRIC, int TQ>
static cudaError_t launch_scan_tq(const ScanArgs& a, cudaStream_t st) {
    auto kern = dense_scan_simt<T, METRIC, TQ>;
    size_t smem = dense_scan_simt_smem(TQ, a.cap);
    LB_SMEM_OPTIN(kern);
    dim3 grid(a.parts, (a.nq + TQ - 1) / TQ);
    kern<<<grid, NT, smem, st>>>((const T*)a.db, a.aux, a.n_rows, a.dim, (const T*)a.queries, a.nq, a.tomb,
                                 a.tomb_bits, a.allow, a.kc, a.cap, a.rows_per_part, a.partial);
    count_launch();
    return cudaGetLastError();
}

template <typename T, int METRIC>
static cudaError_t launch_scan_metric(const ScanArgs& a, cudaStream_t st) {
    switch (a.tq) {
        case 64: return launch_scan_tq<T, METRIC, 64>(a, st);
        case 32: return launch_scan_tq<T, METRIC, 32>(a, st);
        default: return launch_scan_tq<T, METRIC, 16>(a, st);
    }
}

template <typename T>
static cudaError_t launch_scan_dtype(const ScanArgs& a, cudaStream_t st) {
    switch (a.metric) {
        case METRIC_L2: return launch_scan_metric<T, METRIC_L2>(a, st);
        case METRIC_COSINE: return launch_scan_metric<T, METRIC_COSINE>(a, st);
        default: return launch_scan_metric<T, METRIC_DOT>(a, st);
    }
}

cudaError_t launch_dense_scan_simt(const ScanArgs& a, cudaStream_t st) {
    switch (a.dtype) {
        case DT_F32: return launch_scan_dtype<float>(a, st);
        case DT_F16: return launch_scan_dtype<__half>(a, st);
        case DT_I8: return launch_scan_dtype<int8_t>(a, st);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_row_aux(int dtype, const void* db, int64_t n, int dim, int metric, float* aux, int64_t row0,
                           cudaStream_t st) {
    if (n <= row0) return cudaSuccess;
    int64_t rows = n - row0;
    unsigned blocks = (unsigned)((rows + 7) / 8);
    switch (dtype) {
        case DT_F32: row_aux_kernel<float><<<blocks, 256, 0, st>>>((const float*)db, n, dim, metric, aux, row0); break;
        case DT_F16: row_aux_kernel<__half><<<blocks, 256, 0, st>>>((const __half*)db, n, dim, metric, aux, row0); break;
        case DT_I8: row_aux_kernel<int8_t><<<blocks, 256, 0, st>>>((const int8_t*)db, n, dim, metric, aux, row0); break;
        default: return cudaErrorInvalidValue;
    }
    count_launch();
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// S2: merge partial lists.  grid = nq, block = 256.  Streams parts*kc entries through a
// 4096-entry shared buffer, keeping the best kc after each refill.
// ---------------------------------------------------------------------------------------------
constexpr int MERGE_BUF = 4096;

__global__ void __launch_bounds__(256)
merge_partials_kernel(const uint64_t* __restrict__ partial, int parts, int nq, int kc,
                      uint64_t* __restrict__ merged) {
    __shared__ uint64_t buf[MERGE_BUF];
    const int q = blockIdx.x;
    const int total = parts * kc;
    int have = 0;  // entries [0,have) hold the best so far (sorted)
    int next = 0;  // next flat entry index to read
    while (next < total || have == 0) {
        int room = MERGE_BUF - have;
        int take = min(room, total - next);
        for (int t = threadIdx.x; t < take; t += blockDim.x) {
            int e = next + t;
            int p = e / kc, j = e % kc;
            buf[have + t] = partial[((size_t)p * nq + q) * kc + j];
        }
        int filled = have + take;
        int n2 = next_pow2(max(filled, 2));
        for (int t = filled + threadIdx.x; t < n2; t += blockDim.x) buf[t] = kInvalid;
        __syncthreads();
        block_bitonic_sort(buf, n2);
        have = min(filled, kc);
        next += take;
        if (take == 0) break;
    }
    for (int t = threadIdx.x; t < kc; t += blockDim.x)
        merged[(size_t)q * kc + t] = (t < have) ? buf[t] : kInvalid;
}

// ---------------------------------------------------------------------------------------------
// S2 (selection form): per query, the kc smallest (key,row) of parts*kc packed entries.
// Entries are first filtered by the final shared threshold of the scan (when there is one) and compacted
// into a 4096-entry shared buffer; the buffer is sorted whenever it fills (keeping the best kc) and once
// at the end.  With tight scan thresholds the lists are nearly empty and a single short sort remains; in
// the worst case (no threshold, every slot valid) this degrades to the streaming merge above.
// Output is sorted ascending.
// ---------------------------------------------------------------------------------------------
constexpr int MSEL_T = 256;
constexpr int MSEL_BATCH = 4;
// 16 KiB of keys and <= 32 registers: 8 CTAs per SM = 1184 resident, so a 1024-query batch is ONE wave (with a 32 KiB
// buffer and 38 registers it was 888 resident = 1.15 waves, i.e. the time of two).  The final threshold of the scan
// usually leaves a few hundred entries per query, far below the capacity.
constexpr int MSEL_CAP = 2048;

__global__ void __launch_bounds__(MSEL_T, 8)
merge_select_kernel(const uint64_t* __restrict__ partial, int parts, int nq, int kc, uint64_t* __restrict__ merged,
                    uint64_t* __restrict__ kth, const float* __restrict__ edges, const uint32_t* __restrict__ edge_cnt,
                    const uint64_t* __restrict__ compact, const uint32_t* __restrict__ counts, size_t stride) {
    __shared__ int s_n;
    __shared__ uint32_t s_thr;
    __shared__ uint64_t s_buf[MSEL_CAP];
    const int q = blockIdx.x, tid = threadIdx.x;
    // list form: partial[part][q][kc].  compact form: `parts` (0 or 1) head lists of kc entries at
    // partial[q*kc], then counts[q] entries at compact[q*stride].
    const int head = parts * kc;
    const int total = (compact != nullptr) ? head + (int)min((size_t)counts[q], stride) : head;
    // Final shared threshold of the scan (dense_tc.cu / dense_stream.cu): the lowest ladder edge whose
    // counters reach kc.  At least kc live rows lie below it, every one of the global kc best among them,
    // and list compaction only ever dropped rows that are not among those -- so everything at or above the
    // edge can be discarded before sorting.
    if (tid == 0) {
        uint32_t thr = 0xffffffffu;
        if (edges != nullptr) {
            uint32_t cum = 0;
            for (int i = 0; i < LB_NEDGE; i++) {
                cum += edge_cnt[(size_t)q * LB_NEDGE + i];
                if (cum >= (uint32_t)kc) {
                    const float e = edges[(size_t)q * LB_NEDGE + i];
                    if (e < INFINITY) thr = float_to_ordered(e);
                    break;
                }
            }
        }
        s_thr = thr;
        s_n = 0;
    }
    __syncthreads();
    const uint32_t thr = s_thr;
    int have = 0;  // s_buf[0, have) = best so far (sorted) after a flush
    auto fetch = [&](int i) -> uint64_t {
        if (i >= total) return kInvalid;
        uint64_t x;
        if (compact == nullptr) x = partial[((size_t)(i / kc) * nq + q) * kc + (i % kc)];
        else x = (i < head) ? partial[(size_t)q * kc + i] : compact[(size_t)q * stride + (i - head)];
        return ((x != kInvalid) && ((uint32_t)(x >> 32) < thr)) ? x : kInvalid;
    };
    // Slots are fetched MSEL_BATCH per thread with all loads in flight (one memory round trip per 1024 slots, not
    // per 256); the buffer is flushed (sorted, best kc kept) before a batch that could overflow it.
    constexpr int BATCH = MSEL_BATCH;
    for (int base = 0; base < total; base += MSEL_T * BATCH) {
        uint64_t x[BATCH];
#pragma unroll
        for (int b = 0; b < BATCH; b++) x[b] = fetch(base + b * MSEL_T + tid);
        int mine = 0;
#pragma unroll
        for (int b = 0; b < BATCH; b++) mine += (x[b] != kInvalid);
        // worst case every slot of the batch survives: flush first if that could overflow
        const int filled = have + s_n;
        __syncthreads();  // everyone has read s_n before anyone appends to it again
        if (filled > MSEL_CAP - MSEL_T * BATCH) {  // block-uniform
            const int n = filled;
            const int n2 = next_pow2(max(n, 2));
            for (int t = n + tid; t < n2; t += MSEL_T) s_buf[t] = kInvalid;
            __syncthreads();
            block_bitonic_sort(s_buf, n2);
            have = min(n, kc);
            if (tid == 0) s_n = 0;
            __syncthreads();
        }
        if (mine) {
            int pos = have + atomicAdd(&s_n, mine);
#pragma unroll
            for (int b = 0; b < BATCH; b++)
                if (x[b] != kInvalid) s_buf[pos++] = x[b];
        }
        __syncthreads();
    }
    const int n = have + s_n;
    const int n2 = next_pow2(max(n, 2));
    for (int t = n + tid; t < n2; t += MSEL_T) s_buf[t] = kInvalid;
    __syncthreads();
    block_bitonic_sort(s_buf, n2);
    for (int t = tid; t < kc; t += MSEL_T) merged[(size_t)q * kc + t] = (t < n) ? s_buf[t] : kInvalid;
    if (kth != nullptr && tid == 0) kth[q] = (n >= kc) ? s_buf[kc - 1] : kInvalid;
}

cudaError_t launch_merge_select(const uint64_t* partial, int parts, int nq, int kc, uint64_t* merged, uint64_t* kth,
                                const float* edges, const uint32_t* edge_cnt, cudaStream_t st) {
    if (nq <= 0) return cudaSuccess;
    if (kc > MSEL_CAP - MSEL_BATCH * MSEL_T) return cudaErrorInvalidValue;
    merge_select_kernel<<<nq, MSEL_T, 0, st>>>(partial, parts, nq, kc, merged, kth, edges, edge_cnt, nullptr, nullptr, 0);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_merge_select_compact(const uint64_t* head, const uint64_t* compact, const uint32_t* counts,
                                        size_t stride, int nq, int kc, uint64_t* merged, const float* edges,
                                        const uint32_t* edge_cnt, cudaStream_t st) {
    if (nq <= 0) return cudaSuccess;
    if (kc > MSEL_CAP - MSEL_BATCH * MSEL_T) return cudaErrorInvalidValue;
    merge_select_kernel<<<nq, MSEL_T, 0, st>>>(head, head ? 1 : 0, nq, kc, merged, nullptr, edges, edge_cnt, compact,
                                               counts, stride);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_merge_partials(const uint64_t* partial, int parts, int nq, int kc, uint64_t* merged,
                                  cudaStream_t st) {
    if (nq <= 0) return cudaSuccess;
    merge_partials_kernel<<<nq, 256, 0, st>>>(partial, parts, nq, kc, merged);
    count_launch();
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Certification of the coarse stage (thread 0 of a re-score block, after the exact sort).
// Every row outside the candidate set has a coarse key >= the kc-th coarse key c_last; its exact key is at
// least c_last - eps.  If that still exceeds the exact key of the k-th result (+ eps for the mapping), no
// outside row can belong to the true top-k.  Exact keys are mapped into coarse-key space:
//   L2  key = |x|^2 - 2 q.x = d^2 - |q|^2      cosine  key = -q.x/|x| = (d - 1) |q|      dot  key = -q.x = d
// ---------------------------------------------------------------------------------------------
template <int METRIC>
__device__ __forceinline__ void certify(const CertArgs& ca, int q, const uint64_t* packed_q, int c, const uint64_t* keys,
                                        int k, float qn2) {
    bool cert = true;
    const uint64_t last = packed_q[c - 1];
    const uint64_t kth = (k - 1 < c) ? keys[k - 1] : kInvalid;
    if (last != kInvalid && kth != kInvalid) {  // fewer candidates than kc: every live row was a candidate
        const float c_last = key_of(last), dk = key_of(kth);
        const float qn = sqrtf(qn2), xn2 = *ca.max_norm2, xn = sqrtf(xn2);
        float ek, eps;
        if (METRIC == METRIC_L2 && ca.key_space == 1) {
            // SIMT scan: keys are |q - x|^2 accumulated in fp32 (its own order): error relative to the value itself
            ek = dk * dk;
            eps = (ca.beta + 4e-6f) * fmaxf(c_last, ek);
        } else if (METRIC == METRIC_L2) { ek = dk * dk - qn2; eps = 2.f * ca.beta * qn * xn + 4e-6f * (xn2 + qn2); }
        else if (METRIC == METRIC_COSINE) { ek = (dk - 1.f) * qn; eps = ca.beta * qn + 2e-6f * qn; }
        else { ek = dk; eps = ca.beta * qn * xn; }
        cert = (c_last - ek) > 2.f * eps;
    }
    ca.flags[q] = cert ? 0u : 1u;
    if (!cert) atomicAdd(ca.count, 1u);
}

// |q|^2 of the block's query (fp32, any order: only used for the certification bound)
__device__ __forceinline__ float block_qnorm2(const float* qf, int dim, float* s_red8) {
    float s = 0.f;
    for (int i = threadIdx.x; i < dim; i += blockDim.x) s = fmaf(qf[i], qf[i], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) s_red8[threadIdx.x >> 5] = s;
    __syncthreads();
    float t = 0.f;
    for (int w = 0; w < (int)((blockDim.x + 31) >> 5); w++) t += s_red8[w];
    __syncthreads();
    return t;
}

// ---------------------------------------------------------------------------------------------
// S3: exact re-score of candidate ids + final (distance,id) sort -> top-k.
//   grid = nq; block = next_pow2(c) threads (>= 32, <= 1024); one thread per candidate.
// Candidate sources:
//   packed   : u64 (key,id) list from S2 (kInvalid = empty)
//   ids32    : uint32 VectorIDs from the host graph walk (re-rank path); out-of-range,
//              tombstoned or predicate-failing ids are dropped here (parallel_search.go:183-229)
// Also evaluates the certification test for the coarse path: the k-th exact
// distance must be below the coarse bound of everything that was NOT a candidate.
// ---------------------------------------------------------------------------------------------
template <typename T, int METRIC>
__global__ void rescore_kernel(const T* __restrict__ db, uint32_t n_rows, int dim, const T* __restrict__ queries,
                               int nq, const uint64_t* __restrict__ packed, const uint32_t* __restrict__ ids32,
                               int c, int k, const uint32_t* __restrict__ tomb, uint32_t tomb_bits,
                               const uint32_t* __restrict__ allow, int64_t id_base, float* __restrict__ out_d,
                               int64_t* __restrict__ out_l, int negate_dot, const float* __restrict__ nrm,
                               const CertArgs ca) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ float s_red32[32];
    const int n2 = blockDim.x;
    uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);  // [n2]
    float* qf = reinterpret_cast<float*>(keys + n2);          // [dim]
    __shared__ float s_qn[4];
    const int q = blockIdx.x;
    const T* qrow = queries + (size_t)q * dim;
    for (int i = threadIdx.x; i < dim; i += blockDim.x) qf[i] = Elem<T>::widen(qrow[i]);
    __syncthreads();
    // cosine with stored row norms: the query's |q|^2 once per block, reference lane order
    // (lane l sums elements 4m+l; the dim % 4 tail goes to lane 0), simd.go:399-450
    const bool use_nrm = (METRIC == METRIC_COSINE) && (nrm != nullptr);
    if (use_nrm) {
        if (threadIdx.x < 4) {
            const int l = threadIdx.x;
            float sacc = 0.f;
            const int main_end = dim - (dim & 3);
            for (int i = l; i < main_end; i += 4) sacc = __fadd_rn(sacc, __fmul_rn(qf[i], qf[i]));
            if (l == 0) for (int i = main_end; i < dim; i++) sacc = __fadd_rn(sacc, __fmul_rn(qf[i], qf[i]));
            s_qn[l] = sacc;
        }
        __syncthreads();
    }

    uint64_t mine = kInvalid;
    if ((int)threadIdx.x < c) {
        uint32_t id = 0xffffffffu;
        bool ok = false;
        if (packed != nullptr) {
            uint64_t p = packed[(size_t)q * c + threadIdx.x];
            if (p != kInvalid) { id = id_of(p); ok = id < n_rows; }
        } else {
            id = ids32[(size_t)q * c + threadIdx.x];
            ok = id < n_rows;
            if (ok && allow != nullptr && !bit_set(allow, id)) ok = false;
            if (ok && tomb != nullptr && id < tomb_bits && bit_set(tomb, id)) ok = false;
        }
        if (ok) {
            const T* row = db + (size_t)id * dim;
            const bool vec_ok = (dim % Elem<T>::kVec == 0) && ((reinterpret_cast<uintptr_t>(row) & 15) == 0);
            float d;
            if (use_nrm) {
                const float dot = exact_pair<T, METRIC_DOT>(qf, row, dim, vec_ok);
                const float na = __fadd_rn(__fadd_rn(__fadd_rn(s_qn[0], s_qn[1]), s_qn[2]), s_qn[3]);
                const float nb = __ldg(nrm + id);
                if (na == 0.f || nb == 0.f) d = 1.0f;
                else d = __fsub_rn(1.0f, __fdiv_rn(dot, (float)sqrt((double)na * (double)nb)));
            } else {
                d = exact_pair<T, METRIC>(qf, row, dim, vec_ok);
            }
            if (METRIC == METRIC_DOT && negate_dot) d = -d;
            if (d < INFINITY) mine = pack_key(d, id);  // NaN / +Inf are never returned
        }
    }
    keys[threadIdx.x] = mine;
    __syncthreads();
    block_bitonic_sort(keys, n2);
    if (ca.flags != nullptr && packed != nullptr) {  // block-uniform
        const float qn2 = block_qnorm2(qf, dim, s_red32);
        if (threadIdx.x == 0) certify<METRIC>(ca, q, packed + (size_t)q * c, c, keys, k, qn2);
    }
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
        uint64_t p = (j < n2) ? keys[j] : kInvalid;
        bool valid = p != kInvalid;
        out_d[(size_t)q * k + j] = valid ? key_of(p) : 3.402823466e+38f;
        out_l[(size_t)q * k + j] = valid ? (int64_t)id_of(p) + id_base : -1;
    }
}

// ---------------------------------------------------------------------------------------------
// S3, cooperative-gather form (rows 16-byte aligned, row bytes a multiple of 16).
// The thread-per-candidate kernel above makes every warp-level load touch 32 different cache lines
// (one 16-byte piece of 32 different rows).  Here a warp gathers its 32 candidate rows together:
// 128-byte chunks of each row are copied global -> shared with cp.async (each warp instruction
// covers 4 rows x 128 contiguous bytes = 4 full lines), triple buffered, and lane l then runs the
// reference-order accumulation of candidate l out of shared memory (row pitch 144 B: conflict free).
// Arithmetic and order are identical to exact_pair(), so results are bit-identical to the kernel above.
// ---------------------------------------------------------------------------------------------
// (RC_* constants, cp.async helpers and rc_gather live in common.cuh: the graph walk uses the same gather)

template <typename T, int METRIC>
__global__ void __launch_bounds__(RC_WARPS * 32)
rescore_coop_kernel(const T* __restrict__ db, uint32_t n_rows, int dim, const T* __restrict__ queries, int nq,
                    const uint64_t* __restrict__ packed, const uint32_t* __restrict__ ids32, int c, int k, int n2,
                    const uint32_t* __restrict__ tomb, uint32_t tomb_bits, const uint32_t* __restrict__ allow,
                    int64_t id_base, float* __restrict__ out_d, int64_t* __restrict__ out_l, int negate_dot,
                    const float* __restrict__ nrm, const CertArgs ca) {
    __shared__ float s_red32[32];
    constexpr int ACC = (METRIC == METRIC_COSINE) ? METRIC_DOT : METRIC;  // cosine: stored |x|^2, dot chain only
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);                      // [n2]
    float* qf = reinterpret_cast<float*>(keys + n2);                              // [dim]
    unsigned char* stage_all = reinterpret_cast<unsigned char*>(qf + ((dim + 3) & ~3));  // [warps][NBUF][32][PITCH]
    __shared__ float s_qn[4];
    const int q = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const T* qrow = queries + (size_t)q * dim;
    for (int i = tid; i < dim; i += blockDim.x) qf[i] = Elem<T>::widen(qrow[i]);
    for (int i = tid; i < n2; i += blockDim.x) keys[i] = kInvalid;
    __syncthreads();
    if (METRIC == METRIC_COSINE) {
        // |q|^2 once per block in the reference lane order (simd.go:399-450)
        if (tid < 4) {
            float sacc = 0.f;
            const int main_end = dim - (dim & 3);
            for (int i = tid; i < main_end; i += 4) sacc = __fadd_rn(sacc, __fmul_rn(qf[i], qf[i]));
            if (tid == 0) for (int i = main_end; i < dim; i++) sacc = __fadd_rn(sacc, __fmul_rn(qf[i], qf[i]));
            s_qn[tid] = sacc;
        }
        __syncthreads();
    }
    const size_t row_bytes = (size_t)dim * sizeof(T);
    const int n_chunks = (int)((row_bytes + RC_CHUNK - 1) / RC_CHUNK);
    unsigned char* stage = stage_all + (size_t)warp * RC_NBUF * 32 * RC_PITCH;

    // Re-rank mode: drop out-of-range / tombstoned / filtered ids FIRST and compact the survivors, so every
    // warp gathers 32 live rows per trip (with a 30% predicate and 5% tombstones three lanes in four would
    // otherwise idle through the whole gather).  Order does not matter: the result is sorted by (distance, id).
    uint32_t* live = reinterpret_cast<uint32_t*>(stage_all + (size_t)RC_WARPS * RC_NBUF * 32 * RC_PITCH);  // [n2]
    __shared__ int s_live;
    int n_cand = c;
    if (packed == nullptr) {
        if (tid == 0) s_live = 0;
        __syncthreads();
        for (int ci = tid; ci < ((c + 31) & ~31); ci += blockDim.x) {
            uint32_t id = 0xffffffffu;
            bool ok = false;
            if (ci < c) {
                id = ids32[(size_t)q * c + ci];
                ok = id < n_rows;
                if (ok && allow != nullptr && !bit_set(allow, id)) ok = false;
                if (ok && tomb != nullptr && id < tomb_bits && bit_set(tomb, id)) ok = false;
            }
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            int b = 0;
            if (lane == 0 && m) b = atomicAdd(&s_live, __popc(m));
            b = __shfl_sync(0xffffffffu, b, 0);
            if (ok) live[b + __popc(m & ((1u << lane) - 1u))] = id;
        }
        __syncthreads();
        n_cand = s_live;
    }

    // Candidates are dealt to the warps in equal contiguous shares (not 32 at a time): with bitmaps in force only
    // ~36 of 128 ids survive, and dealing by 32 left three of the four warps idle through the whole gather.
    const int per_warp = (n_cand + RC_WARPS - 1) / RC_WARPS;
    const int w_begin = warp * per_warp, w_end = min(n_cand, w_begin + per_warp);
    for (int base = w_begin; base < w_end; base += 32) {
        // candidate of this lane
        const int ci = base + lane;
        uint32_t id = 0xffffffffu;
        bool ok = false;
        if (ci < w_end) {
            if (packed != nullptr) {
                const uint64_t p = packed[(size_t)q * c + ci];
                if (p != kInvalid) { id = id_of(p); ok = id < n_rows; }
            } else {
                id = live[ci];
                ok = true;
            }
        }
        // rows this lane copies: copy instruction i moves piece (lane % 8) of row 4i + lane / 8
        const unsigned char* src_row[8];
        bool src_ok[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int r = 4 * i + (lane >> 3);
            const uint32_t rid = __shfl_sync(0xffffffffu, id, r);
            src_ok[i] = __shfl_sync(0xffffffffu, ok ? 1 : 0, r) != 0;
            src_row[i] = reinterpret_cast<const unsigned char*>(db) + (size_t)rid * row_bytes + (lane & 7) * 16;
        }
        // The staging space holds 3 chunks of 32 rows.  A trip with <= 16 (<= 8) rows lays the same space out as
        // 6 (12) chunk buffers of 16 (8) rows: a short candidate list (bitmaps in force) gets a deeper pipeline
        // instead of idle buffer rows -- 12 buffers cover a whole 1.5 KB row, i.e. ONE memory round trip.
        const int rows_here = min(32, w_end - base);
        ExactAcc<ACC> acc;
        acc.init();
        if (rows_here <= 8) rc_gather<T, ACC, 12, 8>(stage, src_row, src_ok, row_bytes, n_chunks, dim, qf, ok, lane, acc);
        else if (rows_here <= 16) rc_gather<T, ACC, 6, 16>(stage, src_row, src_ok, row_bytes, n_chunks, dim, qf, ok, lane, acc);
        else rc_gather<T, ACC, 3, 32>(stage, src_row, src_ok, row_bytes, n_chunks, dim, qf, ok, lane, acc);
        if (ok) {
            float d;
            if (METRIC == METRIC_COSINE) {
                const float dot = acc.finish();
                const float na = __fadd_rn(__fadd_rn(__fadd_rn(s_qn[0], s_qn[1]), s_qn[2]), s_qn[3]);
                const float nb = __ldg(nrm + id);
                if (na == 0.f || nb == 0.f) d = 1.0f;
                else d = __fsub_rn(1.0f, __fdiv_rn(dot, (float)sqrt((double)na * (double)nb)));
            } else {
                d = acc.finish();
            }
            if (METRIC == METRIC_DOT && negate_dot) d = -d;
            if (d < INFINITY) keys[ci] = pack_key(d, id);  // NaN / +Inf are never returned
        }
    }
    __syncthreads();
    block_bitonic_sort(keys, n2);
    if (ca.flags != nullptr && packed != nullptr) {  // block-uniform
        const float qn2 = block_qnorm2(qf, dim, s_red32);
        if (tid == 0) certify<METRIC>(ca, q, packed + (size_t)q * c, c, keys, k, qn2);
    }
    for (int j = tid; j < k; j += blockDim.x) {
        const uint64_t p = (j < n2) ? keys[j] : kInvalid;
        const bool valid = p != kInvalid;
        out_d[(size_t)q * k + j] = valid ? key_of(p) : 3.402823466e+38f;
        out_l[(size_t)q * k + j] = valid ? (int64_t)id_of(p) + id_base : -1;
    }
}

// ---------------------------------------------------------------------------------------------
// S3, warp-per-query form.  Same gather and the same arithmetic as rescore_coop_kernel, but one WARP owns a query
// end to end (ids -> bitmap tests -> gather -> exact distances -> sort -> certification -> output) and never
// meets a block barrier: the four warps of a CTA are four independent queries.  The block form spent most of a
// short-list call (bitmaps in force: ~36 live ids of 128) waiting on its own serial latency chain -- ids, bitmap
// words, rows, sort, each behind a __syncthreads -- with three CTAs per SM; here twelve to twenty queries per SM
// are in flight at different stages.  SR = staging row-buffers per warp (96: trips of 32 rows; 48: trips of 16,
// half the shared memory, for lists that bitmaps are going to thin out).
// ---------------------------------------------------------------------------------------------
template <typename T, int METRIC, int SR>
__global__ void __launch_bounds__(RC_WARPS * 32)
rescore_warp_kernel(const T* __restrict__ db, uint32_t n_rows, int dim, const T* __restrict__ queries, int nq,
                    const uint64_t* __restrict__ packed, const uint32_t* __restrict__ ids32, int c, int k, int n2,
                    const uint32_t* __restrict__ tomb, uint32_t tomb_bits, const uint32_t* __restrict__ allow,
                    int64_t id_base, float* __restrict__ out_d, int64_t* __restrict__ out_l, int negate_dot,
                    const float* __restrict__ nrm, const CertArgs ca) {
    constexpr int ACC = (METRIC == METRIC_COSINE) ? METRIC_DOT : METRIC;  // cosine: stored |x|^2, dot chain only
    constexpr int TRIP = (SR >= 96) ? 32 : 16;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.x * RC_WARPS + warp;
    if (q >= nq) return;
    const size_t per_warp = (size_t)n2 * 8 + (size_t)n2 * 4 + (size_t)((dim + 3) & ~3) * 4 + (size_t)SR * RC_PITCH;
    unsigned char* base_p = smem_raw + (size_t)warp * per_warp;
    uint64_t* keys = reinterpret_cast<uint64_t*>(base_p);                 // [n2]
    uint32_t* live = reinterpret_cast<uint32_t*>(keys + n2);              // [n2]
    float* qf = reinterpret_cast<float*>(live + n2);                      // [dim]
    unsigned char* stage = reinterpret_cast<unsigned char*>(qf + ((dim + 3) & ~3));
    const T* qrow = queries + (size_t)q * dim;
    for (int i = lane; i < dim; i += 32) qf[i] = Elem<T>::widen(qrow[i]);
    for (int i = lane; i < n2; i += 32) keys[i] = kInvalid;
    __syncwarp();
    float qn_exact = 0.f;  // cosine: |q|^2 in the reference lane order (simd.go:399-450), lanes 0..3 hold the lanes
    if (METRIC == METRIC_COSINE) {
        float sacc = 0.f;
        if (lane < 4) {
            const int main_end = dim - (dim & 3);
            for (int i = lane; i < main_end; i += 4) sacc = __fadd_rn(sacc, __fmul_rn(qf[i], qf[i]));
            if (lane == 0) for (int i = main_end; i < dim; i++) sacc = __fadd_rn(sacc, __fmul_rn(qf[i], qf[i]));
        }
        const float s0 = __shfl_sync(0xffffffffu, sacc, 0), s1 = __shfl_sync(0xffffffffu, sacc, 1);
        const float s2 = __shfl_sync(0xffffffffu, sacc, 2), s3 = __shfl_sync(0xffffffffu, sacc, 3);
        qn_exact = __fadd_rn(__fadd_rn(__fadd_rn(s0, s1), s2), s3);
    }
    const size_t row_bytes = (size_t)dim * sizeof(T);
    const int n_chunks = (int)((row_bytes + RC_CHUNK - 1) / RC_CHUNK);
    // live candidate ids, compacted (order is irrelevant: the result is sorted by (distance, id))
    int n_cand = 0;
    for (int c0 = 0; c0 < c; c0 += 32) {
        const int ci = c0 + lane;
        uint32_t id = 0xffffffffu;
        bool ok = false;
        if (ci < c) {
            if (packed != nullptr) {
                const uint64_t p = packed[(size_t)q * c + ci];
                if (p != kInvalid) { id = id_of(p); ok = id < n_rows; }
            } else {
                id = ids32[(size_t)q * c + ci];
                ok = id < n_rows;
                if (ok && allow != nullptr && !bit_set(allow, id)) ok = false;
                if (ok && tomb != nullptr && id < tomb_bits && bit_set(tomb, id)) ok = false;
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (ok) live[n_cand + __popc(m & ((1u << lane) - 1u))] = id;
        n_cand += __popc(m);
    }
    __syncwarp();
    for (int base = 0; base < n_cand; base += TRIP) {
        const int ci = base + lane;
        const bool ok = lane < TRIP && ci < n_cand;
        const uint32_t id = ok ? live[ci] : 0xffffffffu;
        const unsigned char* src_row[8];
        bool src_ok[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int r = 4 * i + (lane >> 3);
            const uint32_t rid = __shfl_sync(0xffffffffu, id, r);
            src_ok[i] = __shfl_sync(0xffffffffu, ok ? 1 : 0, r) != 0;
            src_row[i] = reinterpret_cast<const unsigned char*>(db) + (size_t)rid * row_bytes + (lane & 7) * 16;
        }
        const int rows_here = min(TRIP, n_cand - base);
        ExactAcc<ACC> acc;
        acc.init();
        if (rows_here <= 8) rc_gather<T, ACC, SR / 8, 8>(stage, src_row, src_ok, row_bytes, n_chunks, dim, qf, ok, lane, acc);
        else if (rows_here <= 16 || SR < 96) rc_gather<T, ACC, SR / 16, 16>(stage, src_row, src_ok, row_bytes, n_chunks, dim, qf, ok, lane, acc);
        else rc_gather<T, ACC, (SR >= 96 ? SR / 32 : 3), 32>(stage, src_row, src_ok, row_bytes, n_chunks, dim, qf, ok, lane, acc);
        if (ok) {
            float d;
            if (METRIC == METRIC_COSINE) {
                const float dot = acc.finish();
                const float nb = __ldg(nrm + id);
                if (qn_exact == 0.f || nb == 0.f) d = 1.0f;
                else d = __fsub_rn(1.0f, __fdiv_rn(dot, (float)sqrt((double)qn_exact * (double)nb)));
            } else {
                d = acc.finish();
            }
            if (METRIC == METRIC_DOT && negate_dot) d = -d;
            if (d < INFINITY) keys[ci] = pack_key(d, id);  // NaN / +Inf are never returned
        }
    }
    __syncwarp();
    warp_bitonic_sort(keys, next_pow2(max(n_cand, 2)), lane);
    if (ca.flags != nullptr && packed != nullptr) {
        float s = 0.f;
        for (int i = lane; i < dim; i += 32) s = fmaf(qf[i], qf[i], s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) certify<METRIC>(ca, q, packed + (size_t)q * c, c, keys, k, s);
    }
    for (int j = lane; j < k; j += 32) {
        const uint64_t p = (j < n2) ? keys[j] : kInvalid;
        const bool valid = p != kInvalid;
        out_d[(size_t)q * k + j] = valid ? key_of(p) : 3.402823466e+38f;
        out_l[(size_t)q * k + j] = valid ? (int64_t)id_of(p) + id_base : -1;
    }
}

template <typename T>
static cudaError_t launch_rescore_t(const RescoreArgs& a, cudaStream_t st) {
    int n2 = next_pow2(max(a.c, 32));
    if (n2 > 1024) return cudaErrorInvalidValue;
    const CertArgs ca{a.cert_flags, a.cert_count, a.max_norm2, a.beta, a.key_space};
    // cooperative-gather kernel whenever rows are 16-byte aligned multiples of 16 bytes
    const size_t row_bytes = (size_t)a.dim * sizeof(T);
    const bool coop = (row_bytes % 16 == 0) && ((reinterpret_cast<uintptr_t>(a.db) & 15) == 0) &&
                      !(a.metric == METRIC_COSINE && a.nrm == nullptr) && !g_rescore_legacy;
    // few queries: the block-per-query kernel spreads one query's candidates over four warps (latency); many
    // queries: one warp per query keeps every SM full of independent queries (throughput)
    if (coop && !g_rescore_block && a.nq >= 256) {
        // warp-per-query kernel; lists that bitmaps will thin out get the half-size staging (more queries per SM)
        const bool thin = a.packed == nullptr && (a.allow != nullptr || a.tomb != nullptr);
        const int SRv = thin ? 48 : 96;
        const size_t per_warp = (size_t)n2 * 12 + (size_t)((a.dim + 3) & ~3) * 4 + (size_t)SRv * RC_PITCH;
        const size_t smem = per_warp * RC_WARPS;
        if (smem <= 200 * 1024) {
            const int grid = (a.nq + RC_WARPS - 1) / RC_WARPS;
#define LB_RW(M, S)                                                                                         \
    {                                                                                                       \
        auto kern = rescore_warp_kernel<T, M, S>;                                                           \
        LB_SMEM_OPTIN(kern);                                                                                \
        kern<<<grid, RC_WARPS * 32, smem, st>>>((const T*)a.db, a.n_rows, a.dim, (const T*)a.queries, a.nq, \
                                                a.packed, a.ids32, a.c, a.k, n2, a.tomb, a.tomb_bits, a.allow, \
                                                a.id_base, a.out_d, a.out_l, a.negate_dot, a.nrm, ca);      \
    }
            switch (a.metric) {
                case METRIC_L2: if (thin) LB_RW(METRIC_L2, 48) else LB_RW(METRIC_L2, 96) break;
                case METRIC_COSINE: if (thin) LB_RW(METRIC_COSINE, 48) else LB_RW(METRIC_COSINE, 96) break;
                default: if (thin) LB_RW(METRIC_DOT, 48) else LB_RW(METRIC_DOT, 96) break;
            }
#undef LB_RW
            count_launch();
            return cudaGetLastError();
        }
    }
    if (coop) {
        const size_t smem = (size_t)n2 * 8 + (size_t)((a.dim + 3) & ~3) * 4 + (size_t)RC_WARPS * RC_NBUF * 32 * RC_PITCH +
                            (size_t)n2 * 4;
#define LB_RC(M)                                                                                            \
    {                                                                                                       \
        auto kern = rescore_coop_kernel<T, M>;                                                              \
        LB_SMEM_OPTIN(kern);                                                                                \
        kern<<<a.nq, RC_WARPS * 32, smem, st>>>((const T*)a.db, a.n_rows, a.dim, (const T*)a.queries, a.nq, \
                                                a.packed, a.ids32, a.c, a.k, n2, a.tomb, a.tomb_bits, a.allow, \
                                                a.id_base, a.out_d, a.out_l, a.negate_dot, a.nrm, ca);      \
    }
        switch (a.metric) {
            case METRIC_L2: LB_RC(METRIC_L2) break;
            case METRIC_COSINE: LB_RC(METRIC_COSINE) break;
            default: LB_RC(METRIC_DOT) break;
        }
#undef LB_RC
        count_launch();
        return cudaGetLastError();
    }
    size_t smem = (size_t)n2 * 8 + (size_t)a.dim * 4;
#define LB_RS(M)                                                                                            \
    {                                                                                                       \
        auto kern = rescore_kernel<T, M>;                                                                   \
        LB_SMEM_OPTIN(kern);                                                                                \
        kern<<<a.nq, n2, smem, st>>>((const T*)a.db, a.n_rows, a.dim, (const T*)a.queries, a.nq, a.packed,  \
                                     a.ids32, a.c, a.k, a.tomb, a.tomb_bits, a.allow, a.id_base, a.out_d,   \
                                     a.out_l, a.negate_dot, a.nrm, ca);                                     \
    }
    switch (a.metric) {
        case METRIC_L2: LB_RS(METRIC_L2) break;
        case METRIC_COSINE: LB_RS(METRIC_COSINE) break;
        default: LB_RS(METRIC_DOT) break;
    }
#undef LB_RS
    count_launch();
    return cudaGetLastError();
}

// max |x|^2 over rows [row0, n): one warp per row, atomicMax on the float bits (non-negative floats order as uints)
template <typename T>
__global__ void row_maxnorm_kernel(const T* __restrict__ db, int64_t n, int dim, int64_t row0, float* __restrict__ out) {
    const int64_t r = row0 + (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= n) return;
    const T* row = db + r * dim;
    float s = 0.f;
    for (int i = lane; i < dim; i += 32) {
        const float v = Elem<T>::widen(row[i]);
        s = fmaf(v, v, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0 && s == s) atomicMax(reinterpret_cast<unsigned int*>(out), __float_as_uint(fabsf(s)));
}

cudaError_t launch_row_maxnorm(int dtype, const void* db, int64_t n, int dim, int64_t row0, float* max_norm2,
                               cudaStream_t st) {
    if (n <= row0) return cudaSuccess;
    const unsigned blocks = (unsigned)((n - row0 + 7) / 8);
    switch (dtype) {
        case DT_F32: row_maxnorm_kernel<float><<<blocks, 256, 0, st>>>((const float*)db, n, dim, row0, max_norm2); break;
        case DT_F16: row_maxnorm_kernel<__half><<<blocks, 256, 0, st>>>((const __half*)db, n, dim, row0, max_norm2); break;
        case DT_I8: row_maxnorm_kernel<int8_t><<<blocks, 256, 0, st>>>((const int8_t*)db, n, dim, row0, max_norm2); break;
        default: return cudaErrorInvalidValue;
    }
    count_launch();
    return cudaGetLastError();
}

// exact fallback: rows excluded by the bitmaps get NaN (select_k ignores NaN)
__global__ void mask_rows_kernel(float* __restrict__ dist, int64_t n, const uint32_t* __restrict__ tomb, uint32_t tomb_bits,
                                 const uint32_t* __restrict__ allow) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool ok = true;
    if (tomb != nullptr && (uint32_t)i < tomb_bits && bit_set(tomb, (uint32_t)i)) ok = false;
    if (ok && allow != nullptr && !bit_set(allow, (uint32_t)i)) ok = false;
    if (!ok) dist[i] = __int_as_float(0x7fc00000);
}

cudaError_t launch_mask_rows(float* dist, int64_t n, const uint32_t* tomb, uint32_t tomb_bits, const uint32_t* allow,
                             cudaStream_t st) {
    if (n <= 0 || (tomb == nullptr && allow == nullptr)) return cudaSuccess;
    mask_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dist, n, tomb, tomb_bits, allow);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_rescore(const RescoreArgs& a, cudaStream_t st) {
    if (a.nq <= 0) return cudaSuccess;
    switch (a.dtype) {
        case DT_F32: return launch_rescore_t<float>(a, st);
        case DT_F16: return launch_rescore_t<__half>(a, st);
        case DT_I8: return launch_rescore_t<int8_t>(a, st);
        default: return cudaErrorInvalidValue;
    }
}

// ---------------------------------------------------------------------------------------------
// One query x n rows -> n exact distances (simd.*DistanceBatch* semantics: dot is RAW here).
// One thread per row; the query sits in shared memory.
// ---------------------------------------------------------------------------------------------
template <typename T, int METRIC>
__global__ void batch_flat_kernel(const T* __restrict__ db, int64_t n, int dim, const T* __restrict__ query,
                                  float* __restrict__ out, int negate_dot) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* qf = reinterpret_cast<float*>(smem_raw);
    for (int i = threadIdx.x; i < dim; i += blockDim.x) qf[i] = Elem<T>::widen(query[i]);
    __syncthreads();
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const T* row = db + r * dim;
    const bool vec_ok = (dim % Elem<T>::kVec == 0) && ((reinterpret_cast<uintptr_t>(row) & 15) == 0);
    float d = exact_pair<T, METRIC>(qf, row, dim, vec_ok);
    if (METRIC == METRIC_DOT && negate_dot) d = -d;
    out[r] = d;
}

__global__ void batch_flat_sq8_kernel(const uint8_t* __restrict__ db, int64_t n, int dim,
                                      const uint8_t* __restrict__ query, float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* qf = reinterpret_cast<float*>(smem_raw);
    for (int i = threadIdx.x; i < dim; i += blockDim.x) qf[i] = (float)query[i];
    __syncthreads();
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    out[r] = exact_sq8(qf, db + r * dim, dim);
}

cudaError_t launch_batch_flat(int metric, int dtype, const void* db, int64_t n, int dim, const void* query,
                              float* out, int negate_dot, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    unsigned blocks = (unsigned)((n + 127) / 128);
    size_t smem = (size_t)dim * 4;
#define LB_BF(T, M) batch_flat_kernel<T, M><<<blocks, 128, smem, st>>>((const T*)db, n, dim, (const T*)query, out, negate_dot)
    if (dtype == DT_U8) {
        if (metric != METRIC_L2) return cudaErrorInvalidValue;
        batch_flat_sq8_kernel<<<blocks, 128, smem, st>>>((const uint8_t*)db, n, dim, (const uint8_t*)query, out);
    } else if (dtype == DT_F32) {
        if (metric == METRIC_L2) LB_BF(float, METRIC_L2); else if (metric == METRIC_COSINE) LB_BF(float, METRIC_COSINE); else LB_BF(float, METRIC_DOT);
    } else if (dtype == DT_F16) {
        if (metric == METRIC_L2) LB_BF(__half, METRIC_L2); else if (metric == METRIC_COSINE) LB_BF(__half, METRIC_COSINE); else LB_BF(__half, METRIC_DOT);
    } else if (dtype == DT_I8) {
        if (metric == METRIC_L2) LB_BF(int8_t, METRIC_L2); else if (metric == METRIC_DOT) LB_BF(int8_t, METRIC_DOT); else return cudaErrorInvalidValue;
    } else {
        return cudaErrorInvalidValue;
    }
#undef LB_BF
    count_launch();
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// select_k over one distance array (Arrow compute "select_k_neighbors"), and the shard merge.
// Both reduce to: build packed (distance, index) keys, block-sort chunks, keep k.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
select_k_chunks_kernel(const float* __restrict__ d, int64_t n, int kc, uint64_t* __restrict__ partial) {
    __shared__ uint64_t buf[MERGE_BUF];
    int64_t base = (int64_t)blockIdx.x * MERGE_BUF;
    for (int t = threadIdx.x; t < MERGE_BUF; t += blockDim.x) {
        int64_t i = base + t;
        uint64_t p = kInvalid;
        if (i < n) {
            float v = d[i];
            if (v == v) p = pack_key(v, (uint32_t)i);
        }
        buf[t] = p;
    }
    __syncthreads();
    block_bitonic_sort(buf, MERGE_BUF);
    for (int t = threadIdx.x; t < kc; t += blockDim.x) partial[(size_t)blockIdx.x * kc + t] = buf[t];
}

__global__ void unpack_topk_kernel(const uint64_t* __restrict__ merged, int nq, int kc, int k, int64_t id_base,
                                   float* __restrict__ out_d, int64_t* __restrict__ out_l) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq * k) return;
    int q = i / k, j = i % k;
    uint64_t p = (j < kc) ? merged[(size_t)q * kc + j] : kInvalid;
    bool valid = p != kInvalid;
    if (out_d) out_d[i] = valid ? key_of(p) : 3.402823466e+38f;
    out_l[i] = valid ? (int64_t)id_of(p) + id_base : -1;
}

// Chunks of 4096 distances are sorted by one block each (the k best kept), then merged in two levels: `groups` blocks
// merge `per` chunk lists each, one block merges the group lists.  (One block streaming every chunk list through its
// buffer took chunks * k / (4096 - k) sorts in sequence -- 245 x 2048 entries, 10.9 ms, for k = 2048 over 1 M rows.)
static void select_k_plan(int64_t n, int* groups, int* per) {
    int chunks = (int)((n + MERGE_BUF - 1) / MERGE_BUF);
    if (chunks < 1) chunks = 1;
    if (chunks <= 8) { *groups = 1; *per = chunks; return; }
    int g = 1;
    while (g * g < chunks) g++;
    *groups = g;
    *per = (chunks + g - 1) / g;
}

size_t select_k_scratch_entries(int64_t n, int k) {
    int groups, per;
    select_k_plan(n, &groups, &per);
    return ((size_t)groups * per + (groups > 1 ? groups : 0)) * (size_t)k;
}

cudaError_t launch_select_k(const float* d, int64_t n, int k, uint64_t* scratch_partial, uint64_t* scratch_merged,
                            int64_t* out_idx, float* out_d, int64_t id_base, cudaStream_t st) {
    if (k > MERGE_BUF / 2) return cudaErrorInvalidValue;  // the streaming merge keeps k entries and needs room to refill
    int kc = k;
    int groups, per;
    select_k_plan(n, &groups, &per);
    // groups * per >= chunks: the blocks past the data emit empty (all-invalid) lists, so every group has `per` lists
    select_k_chunks_kernel<<<groups * per, 256, 0, st>>>(d, n, kc, scratch_partial);
    count_launch();
    if (groups > 1) {
        uint64_t* level1 = scratch_partial + (size_t)groups * per * kc;
        // group q merges chunk lists q, q + groups, q + 2 groups, ... (the [part][query][kc] layout of the kernel)
        merge_partials_kernel<<<groups, 256, 0, st>>>(scratch_partial, per, groups, kc, level1);
        count_launch();
        merge_partials_kernel<<<1, 256, 0, st>>>(level1, groups, 1, kc, scratch_merged);
    } else {
        merge_partials_kernel<<<1, 256, 0, st>>>(scratch_partial, per, 1, kc, scratch_merged);
    }
    count_launch();
    unpack_topk_kernel<<<(k + 127) / 128, 128, 0, st>>>(scratch_merged, 1, kc, k, id_base, out_d, out_idx);
    count_launch();
    return cudaGetLastError();
}

// Shard merge: (distance,label) lists with int64 labels.  Lists are short (parts * k_in), so one
// block per query sorts (distance, slot) keys and then maps slots back to labels.
// Shard merge: [parts] sorted (distance, int64 label) lists per query -> the k best by (distance, label).
// Part pp's lists start at in_d_base + pp * stride_d (distances, [nq][k_in] f32) and in_l_base + pp * stride_l
// (labels, [nq][k_in] i64), strides in bytes: two separate [parts][nq][k_in] arrays and the packed per-rank
// exchange records ([distances | labels] per rank: one all-gather, or one peer push) are the same kernel with
// different strides.
// Order: sort by (distance, slot), then every run of equal distances that starts inside the first k outputs is
// re-ranked by label -- O(run) label reads per entry (the earlier version re-derived the whole ranking per
// output thread, O(run^2) each).
__global__ void __launch_bounds__(256)
merge_topk_kernel(const char* __restrict__ in_d_base, size_t stride_d, const char* __restrict__ in_l_base,
                  size_t stride_l, int parts, int nq, int k_in, int k, float* __restrict__ out_d,
                  int64_t* __restrict__ out_l, const uint32_t* wait_flags, uint32_t wait_seq, uint32_t* err) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (wait_flags != nullptr) {
        // exchange.cu: the records of this batch are pushed by the peers; wait until every source's flag has
        // reached the batch's sequence number (acquire at system scope), at most ~10 s
        if (threadIdx.x == 0) {
            unsigned long long t0, t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
            for (int p = 0; p < parts; p++) {
                for (;;) {
                    uint32_t v;
                    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(wait_flags + p) : "memory");
                    if ((int32_t)(v - wait_seq) >= 0) break;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                    if (t1 - t0 > 10000000000ull) { atomicExch(err, 1u); break; }
                    __nanosleep(200);
                }
            }
        }
        __syncthreads();
    }
    const int total = parts * k_in;
    const int n2 = next_pow2(max(total, 2));
    uint64_t* hi = reinterpret_cast<uint64_t*>(smem_raw);  // [n2] (ordered distance << 32 | slot)
    const int q = blockIdx.x;
    auto dist_at = [&](int t) {
        const int pp = t / k_in, j = t - pp * k_in;
        return __ldcg(reinterpret_cast<const float*>(in_d_base + (size_t)pp * stride_d) + (size_t)q * k_in + j);
    };
    auto label_at = [&](int t) {
        const int pp = t / k_in, j = t - pp * k_in;
        return __ldcg(reinterpret_cast<const long long*>(in_l_base + (size_t)pp * stride_l) + (size_t)q * k_in + j);
    };
    for (int t = threadIdx.x; t < n2; t += blockDim.x) {
        uint64_t p = kInvalid;
        if (t < total && label_at(t) >= 0) p = ((uint64_t)float_to_ordered(dist_at(t)) << 32) | (uint32_t)t;
        hi[t] = p;
    }
    __syncthreads();
    block_bitonic_sort(hi, n2);
    for (int j = threadIdx.x; j < k; j += blockDim.x) {  // padding slots (disjoint from the valid ones below)
        if (j >= n2 || hi[j] == kInvalid) {
            out_d[(size_t)q * k + j] = 3.402823466e+38f;
            out_l[(size_t)q * k + j] = -1;
        }
    }
    for (int t = threadIdx.x; t < n2; t += blockDim.x) {
        const uint64_t p = hi[t];
        if (p == kInvalid) continue;
        const uint32_t dkey = (uint32_t)(p >> 32);
        const bool tie_prev = t > 0 && (uint32_t)(hi[t - 1] >> 32) == dkey;
        const bool tie_next = t + 1 < n2 && hi[t + 1] != kInvalid && (uint32_t)(hi[t + 1] >> 32) == dkey;
        int pos = t;
        const int64_t my = label_at((int)(uint32_t)p);
        if (tie_prev || tie_next) {
            int lo = t;
            while (lo > 0 && (uint32_t)(hi[lo - 1] >> 32) == dkey) lo--;
            if (lo >= k) continue;  // the whole run lies beyond the outputs
            int rank = 0;
            for (int b = lo; b < n2; b++) {
                const uint64_t pb = hi[b];
                if (pb == kInvalid || (uint32_t)(pb >> 32) != dkey) break;
                if (b == t) continue;
                const int64_t lb_ = label_at((int)(uint32_t)pb);
                rank += (lb_ < my) || (lb_ == my && b < t);
            }
            pos = lo + rank;
        }
        if (pos < k) {
            out_d[(size_t)q * k + pos] = ordered_to_float(dkey);
            out_l[(size_t)q * k + pos] = my;
        }
    }
}

cudaError_t launch_merge_topk_strided(const void* in_d_base, size_t stride_d, const void* in_l_base, size_t stride_l,
                                      int parts, int nq, int k_in, int k, float* out_d, int64_t* out_l,
                                      cudaStream_t st) {
    return launch_merge_topk_wait(in_d_base, stride_d, in_l_base, stride_l, parts, nq, k_in, k, out_d, out_l, nullptr, 0,
                                  0, nullptr, st);
}

cudaError_t launch_merge_topk_wait(const void* in_d_base, size_t stride_d, const void* in_l_base, size_t stride_l,
                                   int parts, int nq, int k_in, int k, float* out_d, int64_t* out_l,
                                   const uint32_t* wait_flags, uint32_t wait_seq, int /*rank*/, uint32_t* err,
                                   cudaStream_t st) {
    if (nq <= 0) return cudaSuccess;
    int total = parts * k_in;
    int n2 = next_pow2(total < 2 ? 2 : total);
    size_t smem = (size_t)n2 * 8;
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    LB_SMEM_OPTIN(merge_topk_kernel);
    merge_topk_kernel<<<nq, 256, smem, st>>>((const char*)in_d_base, stride_d, (const char*)in_l_base, stride_l, parts,
                                             nq, k_in, k, out_d, out_l, wait_flags, wait_seq, err);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_merge_topk(const float* in_d, const int64_t* in_l, int parts, int nq, int k_in, int k,
                              float* out_d, int64_t* out_l, cudaStream_t st) {
    return launch_merge_topk_strided(in_d, (size_t)nq * k_in * 4, in_l, (size_t)nq * k_in * 8, parts, nq, k_in, k,
                                     out_d, out_l, st);
}

}  // namespace lb
