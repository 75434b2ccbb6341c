// kernels_pq.cu -- product-quantisation kernels: ADC look-up table build, ADC code scan with
// fused top-k', stateless ADC batch, PQ encode.
//
// Bit-exactness contract (north star): every ADC distance equals the reference's
//   float32(sqrt(float64( ((t[0]+t[1])+t[2])+...+t[M-1] )))   internal/simd/simd.go:345-355
// so the sum is accumulated sequentially in j with plain fp32 adds (no FMA, no tree), and the
// table entries equal L2SquaredFloat32 (4-lane order) of internal/pq/adc_table.go:15-51.
#include "kernels.cuh"

namespace lb {

// ---------------------------------------------------------------------------------------------
// LUT build: grid (nq, M), block K(=256) threads; thread c -> table[q][m*K + c].
// ---------------------------------------------------------------------------------------------
__global__ void adc_lut_kernel(const float* __restrict__ codebooks, int M, int K, int sub,
                               const float* __restrict__ queries, float* __restrict__ luts) {
    extern __shared__ float qs[];  // [sub]
    const int q = blockIdx.x, m = blockIdx.y, c = threadIdx.x;
    const int dim = M * sub;
    for (int i = threadIdx.x; i < sub; i += blockDim.x) qs[i] = queries[(size_t)q * dim + m * sub + i];
    __syncthreads();
    if (c >= K) return;
    const float* cent = codebooks + ((size_t)m * K + c) * sub;
    ExactAcc<METRIC_L2> acc;
    acc.init();
    int i = 0;
    for (; i <= sub - 4; i += 4) {
#pragma unroll
        for (int e = 0; e < 4; e++) acc.add(e, qs[i + e], __ldg(cent + i + e));
    }
    for (; i < sub; i++) acc.add(0, qs[i], __ldg(cent + i));
    luts[((size_t)q * M + m) * K + c] = acc.sum();
}

cudaError_t launch_adc_lut(const float* codebooks, int M, int K, int sub, const float* queries, int nq, float* luts,
                           cudaStream_t st) {
    if (nq <= 0) return cudaSuccess;
    dim3 grid(nq, M);
    int threads = ((K + 31) / 32) * 32;
    adc_lut_kernel<<<grid, threads, sub * sizeof(float), st>>>(codebooks, M, K, sub, queries, luts);
    count_launch();
    return cudaGetLastError();
}

// sequential ADC sum of one code row held in registers / memory
template <bool VEC>
__device__ __forceinline__ float adc_row_sum(const float* __restrict__ lut, const uint8_t* __restrict__ row, int M) {
    float sum = 0.f;
    if (VEC) {
        const uint4* rv = reinterpret_cast<const uint4*>(row);
        const int nv = M >> 4;
        for (int v = 0; v < nv; v++) {
            uint4 raw = __ldg(rv + v);
            const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
            const float* base = lut + (v << 4) * 256;
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int b = 0; b < 4; b++)
                    sum = __fadd_rn(sum, base[(i * 4 + b) * 256 + ((w[i] >> (8 * b)) & 0xff)]);
        }
    } else {
        for (int j = 0; j < M; j++) sum = __fadd_rn(sum, lut[j * 256 + row[j]]);
    }
    return sum;
}

// ---------------------------------------------------------------------------------------------
// ADC scan with fused selection.  grid = (nq, parts): consecutive blocks scan the SAME code
// range for different queries, so the codes are served from L2 after the first touch.
// block = 256 threads, one code row per thread per step; the query's LUT (M*1 KiB) sits in smem.
// ---------------------------------------------------------------------------------------------
constexpr int PQ_NT = 256;

__global__ void __launch_bounds__(PQ_NT)
adc_scan_kernel(const uint8_t* __restrict__ codes, uint32_t n_rows, int M, const float* __restrict__ luts,
                const uint32_t* __restrict__ tomb, uint32_t tomb_bits, const uint32_t* __restrict__ allow,
                int kc, int cap, uint32_t rows_per_part, int nq, uint64_t* __restrict__ partial) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* lut = reinterpret_cast<float*>(smem_raw);                   // [M*256]
    uint64_t* cand = reinterpret_cast<uint64_t*>(lut + (size_t)M * 256);  // [cap]
    __shared__ int s_cnt;
    __shared__ float s_tau;
    const int q = blockIdx.x;
    const uint32_t row_begin = blockIdx.y * rows_per_part;
    const uint32_t row_end = min(n_rows, row_begin + rows_per_part);
    {
        const float4* src = reinterpret_cast<const float4*>(luts + (size_t)q * M * 256);
        float4* dst = reinterpret_cast<float4*>(lut);
        for (int i = threadIdx.x; i < M * 64; i += PQ_NT) dst[i] = __ldg(src + i);
    }
    if (threadIdx.x == 0) { s_cnt = 0; s_tau = INFINITY; }
    __syncthreads();
    const bool vec = (M % 16 == 0) && ((reinterpret_cast<uintptr_t>(codes) & 15) == 0);

    for (uint32_t r0 = row_begin; r0 < row_end; r0 += PQ_NT) {
        const uint32_t row = r0 + threadIdx.x;
        if (row < row_end) {
            const uint8_t* rp = codes + (size_t)row * M;
            float sum = vec ? adc_row_sum<true>(lut, rp, M) : adc_row_sum<false>(lut, rp, M);
            float d = __fsqrt_rn(sum);
            if (d < s_tau) {
                bool ok = true;
                if (tomb != nullptr && row < tomb_bits && bit_set(tomb, row)) ok = false;
                if (ok && allow != nullptr && !bit_set(allow, row)) ok = false;
                if (ok) {
                    int pos = atomicAdd(&s_cnt, 1);
                    if (pos < cap) cand[pos] = pack_key(d, row);
                }
            }
        }
        __syncthreads();
        int c = min(s_cnt, cap);
        if (c > cap - PQ_NT) {  // uniform across the block
            int n2 = next_pow2(c);
            for (int t = c + threadIdx.x; t < n2; t += PQ_NT) cand[t] = kInvalid;
            __syncthreads();
            block_bitonic_sort(cand, n2);
            if (threadIdx.x == 0) {
                s_cnt = min(c, kc);
                s_tau = (c >= kc) ? key_of(cand[kc - 1]) : INFINITY;
            }
            __syncthreads();
        }
    }
    int c = min(s_cnt, cap);
    int n2 = next_pow2(max(c, 2));
    for (int t = c + threadIdx.x; t < n2; t += PQ_NT) cand[t] = kInvalid;
    __syncthreads();
    block_bitonic_sort(cand, n2);
    uint64_t* out = partial + ((size_t)blockIdx.y * nq + q) * kc;
    for (int t = threadIdx.x; t < kc; t += PQ_NT) out[t] = (t < c) ? cand[t] : kInvalid;
}

cudaError_t launch_adc_scan(const PqScanArgs& a, cudaStream_t st) {
    if (a.nq <= 0) return cudaSuccess;
    size_t smem = (size_t)a.M * 1024 + (size_t)a.cap * 8;
    LB_SMEM_OPTIN(adc_scan_kernel);
    dim3 grid(a.nq, a.parts);
    adc_scan_kernel<<<grid, PQ_NT, smem, st>>>(a.codes, a.n_rows, a.M, a.luts, a.tomb, a.tomb_bits, a.allow, a.kc,
                                               a.cap, a.rows_per_part, a.nq, a.partial);
    count_launch();
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Stateless ADC batch (simd.ADCDistanceBatch): table [M*256] -> out[n].
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PQ_NT)
adc_batch_kernel(const float* __restrict__ table, const uint8_t* __restrict__ codes, int M, int64_t n,
                 float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* lut = reinterpret_cast<float*>(smem_raw);
    for (int i = threadIdx.x; i < M * 256; i += PQ_NT) lut[i] = __ldg(table + i);
    __syncthreads();
    const bool vec = (M % 16 == 0) && ((reinterpret_cast<uintptr_t>(codes) & 15) == 0);
    for (int64_t row = (int64_t)blockIdx.x * PQ_NT + threadIdx.x; row < n; row += (int64_t)gridDim.x * PQ_NT) {
        const uint8_t* rp = codes + (size_t)row * M;
        float sum = vec ? adc_row_sum<true>(lut, rp, M) : adc_row_sum<false>(lut, rp, M);
        out[row] = __fsqrt_rn(sum);
    }
}

cudaError_t launch_adc_batch(const float* table, const uint8_t* codes, int M, int64_t n, float* out, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    size_t smem = (size_t)M * 1024;
    LB_SMEM_OPTIN(adc_batch_kernel);
    int64_t blocks = (n + PQ_NT - 1) / PQ_NT;
    if (blocks > 148 * 8) blocks = 148 * 8;
    adc_batch_kernel<<<(unsigned)blocks, PQ_NT, smem, st>>>(table, codes, M, n, out);
    count_launch();
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// PQ encode (internal/pq/encoder.go:76-136, internal/simd/simd.go:278-326): per subspace the first
// centroid with the strictly smallest distance; K <= 16 compares squared distances, larger K
// compares the sqrt'd values (EuclideanDistanceBatchFlat results).
// grid (ceil(n/128), M), block 128: codebook m in smem, one vector per thread.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
pq_encode_kernel(const float* __restrict__ codebooks, int M, int K, int sub, const float* __restrict__ vecs,
                 int64_t n, uint8_t* __restrict__ codes) {
    extern __shared__ float cb[];  // [K*sub]
    const int m = blockIdx.y;
    for (int i = threadIdx.x; i < K * sub; i += blockDim.x) cb[i] = codebooks[(size_t)m * K * sub + i];
    __syncthreads();
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const float* x = vecs + (size_t)v * M * sub + (size_t)m * sub;
    int best = 0;
    float bd = 3.402823466e+38f;
    for (int c = 0; c < K; c++) {
        ExactAcc<METRIC_L2> acc;
        acc.init();
        const float* cent = cb + c * sub;
        int i = 0;
        for (; i <= sub - 4; i += 4) {
#pragma unroll
            for (int e = 0; e < 4; e++) acc.add(e, __ldg(x + i + e), cent[i + e]);
        }
        for (; i < sub; i++) acc.add(0, __ldg(x + i), cent[i]);
        float d = (K <= 16) ? acc.sum() : acc.finish();
        if (c == 0 && K > 16) { bd = d; best = 0; }
        else if (d < bd) { bd = d; best = c; }
    }
    codes[(size_t)v * M + m] = (uint8_t)best;
}

cudaError_t launch_pq_encode(const float* codebooks, int M, int K, int sub, const float* vecs, int64_t n,
                             uint8_t* codes, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    size_t smem = (size_t)K * sub * 4;
    LB_SMEM_OPTIN(pq_encode_kernel);
    dim3 grid((unsigned)((n + 127) / 128), M);
    pq_encode_kernel<<<grid, 128, smem, st>>>(codebooks, M, K, sub, vecs, n, codes);
    count_launch();
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// PQ training: Lloyd's k-means per subspace, TrainKMeans (internal/pq/kmeans.go:64-151) as
// PQEncoder.Train drives it (internal/pq/encoder.go:39-73).  Bit-exact with the reference's arithmetic
// given the same initial rows: the E-step uses L2SquaredFloat32's 4-lane order and a strict '<' argmin,
// and the M-step adds the member vectors of a centroid IN ROW ORDER with plain fp32 adds (one warp per
// (subspace, centroid) walks the assignment array, ballots its members and adds them one at a time,
// lanes = dimensions) before the fp32 division.  All M subspaces advance together; a subspace that meets
// the reference's early-stop rule is frozen by its active[] flag (no host round trips).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
kmeans_assign_kernel(const float* __restrict__ data, int64_t n, int dims, int sub, int K, const float* __restrict__ cent,
                     int32_t* __restrict__ assign, uint32_t* __restrict__ changed, const int32_t* __restrict__ active) {
    extern __shared__ float cb[];  // [K][sub]
    const int m = blockIdx.y;
    if (!active[m]) return;
    for (int i = threadIdx.x; i < K * sub; i += blockDim.x) cb[i] = cent[(size_t)m * K * sub + i];
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* x = data + (size_t)i * dims + (size_t)m * sub;
    int best = -1;
    float bd = 3.402823466e+38f;
    for (int c = 0; c < K; c++) {
        ExactAcc<METRIC_L2> acc;
        acc.init();
        const float* cc = cb + c * sub;
        int j = 0;
        for (; j <= sub - 4; j += 4) {
#pragma unroll
            for (int e = 0; e < 4; e++) acc.add(e, __ldg(x + j + e), cc[j + e]);
        }
        for (; j < sub; j++) acc.add(0, __ldg(x + j), cc[j]);
        const float d = acc.sum();
        if (d < bd) { bd = d; best = c; }
    }
    if (best < 0) best = 0;
    int32_t* a = assign + (size_t)m * n + i;
    if (*a != best) { atomicAdd(changed + m, 1u); *a = best; }
}

constexpr int KM_MAXJ = 4;  // sub <= 128
__global__ void __launch_bounds__(32)
kmeans_update_kernel(const float* __restrict__ data, int64_t n, int dims, int sub, int K, float* __restrict__ cent,
                     const int32_t* __restrict__ assign, const int32_t* __restrict__ active, int iter) {
    const int c = blockIdx.x, m = blockIdx.y, lane = threadIdx.x;
    if (!active[m]) return;
    const int32_t* a = assign + (size_t)m * n;
    float sum[KM_MAXJ] = {0.f, 0.f, 0.f, 0.f};
    int count = 0;
    for (int64_t base = 0; base < n; base += 32) {
        const int64_t i = base + lane;
        unsigned mask = __ballot_sync(0xffffffffu, i < n && a[i] == c);
        count += __popc(mask);
        while (mask) {  // members in increasing row order
            const int b = __ffs(mask) - 1;
            mask &= mask - 1;
            const float* v = data + (size_t)(base + b) * dims + (size_t)m * sub;
#pragma unroll
            for (int t = 0; t < KM_MAXJ; t++) {
                const int j = lane + 32 * t;
                if (j < sub) sum[t] = __fadd_rn(sum[t], __ldg(v + j));
            }
        }
    }
    float* cc = cent + ((size_t)m * K + c) * sub;
#pragma unroll
    for (int t = 0; t < KM_MAXJ; t++) {
        const int j = lane + 32 * t;
        if (j >= sub) continue;
        if (count > 0) {
            cc[j] = __fdiv_rn(sum[t], (float)count);
        } else {  // empty cluster: re-seed (stand-in for the reference's rand.Intn, see oracle/lb_oracle.c)
            const int64_t idx = ((int64_t)c * 7919 + (int64_t)iter * 104729) % n;
            cc[j] = data[(size_t)idx * dims + (size_t)m * sub + j];
        }
    }
}

__global__ void kmeans_finish_iter_kernel(int M, int64_t n, int iter, uint32_t* __restrict__ changed,
                                          int32_t* __restrict__ active, int32_t* __restrict__ iters) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M || !active[m]) return;
    iters[m] = iter + 1;
    if (iter > 0 && (int64_t)changed[m] < n / 1000 + 1) active[m] = 0;  // kmeans.go:146-148
    changed[m] = 0;
}

__global__ void kmeans_init_kernel(const float* __restrict__ data, int dims, int sub, int K, int M,
                                   const int32_t* __restrict__ init_idx, float* __restrict__ cent) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)M * K * sub) return;
    const int j = (int)(t % sub);
    const int c = (int)((t / sub) % K);
    const int m = (int)(t / ((int64_t)sub * K));
    cent[t] = data[(size_t)init_idx[(size_t)m * K + c] * dims + (size_t)m * sub + j];
}

cudaError_t launch_pq_train(const float* d_data, int64_t n, int dims, int M, int K, int max_iter,
                            const int32_t* d_init_idx, float* d_cent, int32_t* d_assign, uint32_t* d_changed,
                            int32_t* d_active, int32_t* d_iters, cudaStream_t st) {
    const int sub = dims / M;
    if (sub > 32 * KM_MAXJ) return cudaErrorInvalidValue;
    const int64_t total = (int64_t)M * K * sub;
    kmeans_init_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_data, dims, sub, K, M, d_init_idx, d_cent);
    count_launch();
    const size_t smem = (size_t)K * sub * 4;
    LB_SMEM_OPTIN(kmeans_assign_kernel);
    for (int it = 0; it < max_iter; it++) {
        dim3 ga((unsigned)((n + 127) / 128), M);
        kmeans_assign_kernel<<<ga, 128, smem, st>>>(d_data, n, dims, sub, K, d_cent, d_assign, d_changed, d_active);
        dim3 gu(K, M);
        kmeans_update_kernel<<<gu, 32, 0, st>>>(d_data, n, dims, sub, K, d_cent, d_assign, d_active, it);
        kmeans_finish_iter_kernel<<<(M + 127) / 128, 128, 0, st>>>(M, n, it, d_changed, d_active, d_iters);
        count_launch(); count_launch(); count_launch();
    }
    return cudaGetLastError();
}

__global__ void unpack_topk_kernel2(const uint64_t* __restrict__ merged, int nq, int kc, int k, int64_t id_base,
                                    float* __restrict__ out_d, int64_t* __restrict__ out_l) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq * k) return;
    int q = i / k, j = i % k;
    uint64_t p = (j < kc) ? merged[(size_t)q * kc + j] : kInvalid;
    bool valid = p != kInvalid;
    out_d[i] = valid ? key_of(p) : 3.402823466e+38f;
    out_l[i] = valid ? (int64_t)id_of(p) + id_base : -1;
}

cudaError_t launch_unpack_topk(const uint64_t* merged, int nq, int kc, int k, int64_t id_base, float* out_d,
                               int64_t* out_l, cudaStream_t st) {
    if (nq <= 0) return cudaSuccess;
    int total = nq * k;
    unpack_topk_kernel2<<<(total + 255) / 256, 256, 0, st>>>(merged, nq, kc, k, id_base, out_d, out_l);
    count_launch();
    return cudaGetLastError();
}

}  // namespace lb
