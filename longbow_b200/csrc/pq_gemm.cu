// pq_gemm.cu -- batched PQ search: coarse stage on the tensor cores.
//
// A batch of Q queries against N codes costs N*M table look-ups per query on the ADC path (internal/pq/adc_table.go,
// internal/simd/simd.go:350 ADCDistanceBatch): at Q = 256, N = 10 M, M = 96 that is 2.5e11 shared-memory look-ups, and
// the look-up kernel (pq_scan.cu) is bound by the shared-memory pipe (64 two-byte entries per clock and SM), not by
// HBM -- arithmetic intensity per code byte grows with Q.  From ~64 queries up it is cheaper to DECODE a slab of rows
// once (codes -> fp16 vectors, |x|^2), run the dense tensor-core scan (dense_tc.cu, L2 keys |x|^2 - 2 q.x) over the
// slab for the whole batch, and keep only the coarse candidates: the exact stage (adc_exact_kernel: the reference's
// sequential fp32 table sums) and the certification decide the answer exactly as on the look-up path, so results stay
// bit-identical to the reference.  The decoded slab is scratch: the index still stores one byte per sub-quantiser.
//
// Error of the coarse key (what the certification must cover): codebook and query elements are rounded to fp16
// (relative 2^-11 each, absolute 2^-25 below the normal range), so with d = q - x and d_h its rounded version
// | |d_h| - |d| | <= 2^-11 (|q| + |x|) + sqrt(dims) 2^-24; the tensor core's truncating fp32 accumulation adds
// beta |q_h||x_h| to the dot product (beta as for the dense fp16 scan).
#include <cuda_fp16.h>

#include "kernels.cuh"

namespace lb {

// fp32 codebooks [M][256][sub] -> fp16, once per handle
__global__ void pq_codebook16_kernel(const float* __restrict__ cb, __half* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __float2half_rn(cb[i]);
}

cudaError_t launch_pq_codebook16(const float* cb, void* out, size_t n, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    pq_codebook16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(cb, (__half*)out, n);
    count_launch();
    return cudaGetLastError();
}

// queries fp32 [nq][dims] -> fp16, plus per query { |q_h|^2 (fp32 sum of the rounded values), |q| }
__global__ void __launch_bounds__(128)
pq_q16_kernel(const float* __restrict__ q, int dims, __half* __restrict__ q16, float2* __restrict__ qn) {
    __shared__ float s_a[4], s_b[4];
    const int qi = blockIdx.x, tid = threadIdx.x;
    float a = 0.f, b = 0.f;
    for (int j = tid; j < dims; j += blockDim.x) {
        const float v = q[(size_t)qi * dims + j];
        const __half h = __float2half_rn(v);
        q16[(size_t)qi * dims + j] = h;
        const float hv = __half2float(h);
        a = fmaf(hv, hv, a);
        b = fmaf(v, v, b);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if ((tid & 31) == 0) { s_a[tid >> 5] = a; s_b[tid >> 5] = b; }
    __syncthreads();
    if (tid == 0) {
        float ta = 0.f, tb = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) { ta += s_a[w]; tb += s_b[w]; }
        qn[qi] = make_float2(ta, sqrtf(tb));
    }
}

cudaError_t launch_pq_q16(const float* q, int nq, int dims, void* q16, float* qn, cudaStream_t st) {
    if (nq <= 0) return cudaSuccess;
    pq_q16_kernel<<<nq, 128, 0, st>>>(q, dims, (__half*)q16, (float2*)qn);
    count_launch();
    return cudaGetLastError();
}

// Per-row |x_h|^2 of the decoded vectors, computed ONCE when codes are added (4 bytes per row next to the M code
// bytes): sum over j of n2[j][code_j] with n2[j][c] = |fp16 centroid (j, c)|^2 (fp32).  One warp per row, lanes
// over j, a fixed shuffle tree: deterministic, so the coarse keys do not vary from run to run.  xmax2 collects the
// largest value (bits of a non-negative float order as unsigned integers).
__global__ void pq_centroid_norm_kernel(const __half* __restrict__ cb16, int sub, float* __restrict__ n2, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float a = 0.f;
    for (int u = 0; u < sub; u++) {
        const float f = __half2float(cb16[i * sub + u]);
        a = fmaf(f, f, a);
    }
    n2[i] = a;
}

cudaError_t launch_pq_centroid_norms(const void* cb16, int M, int sub, float* n2, cudaStream_t st) {
    const size_t n = (size_t)M * 256;
    pq_centroid_norm_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const __half*)cb16, sub, n2, n);
    count_launch();
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256)
pq_row_norm_kernel(const uint8_t* __restrict__ codes, const float* __restrict__ n2, int M, int64_t r0, int64_t n,
                   float* __restrict__ xn2, uint32_t* __restrict__ xmax2) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;   // whole warps
    const uint8_t* c = codes + (size_t)(r0 + row) * M;
    float a = 0.f;
    for (int j = lane; j < M; j += 32) a += __ldg(n2 + (size_t)j * 256 + c[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) {
        xn2[r0 + row] = a;
        atomicMax(xmax2, __float_as_uint(a));
    }
}

cudaError_t launch_pq_row_norms(const uint8_t* codes, const float* n2, int M, int64_t r0, int64_t n, float* xn2,
                                uint32_t* xmax2, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    pq_row_norm_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(codes, n2, M, r0, n, xn2, xmax2);
    count_launch();
    return cudaGetLastError();
}

// Decode rows [r0, r0 + n) of the row-major codes into fp16 vectors x[n][dims].
//
// The codebooks (M * 256 * sub halves: 393 KB at M = 96, sub = 8) do not fit one SM's shared memory, so the
// sub-quantisers are split into G groups of mg whose tables do (<= 192 KiB): CTA (b, g) keeps group g's tables in
// shared memory for its whole life and writes the mg * sub halves of its rows' vectors -- one contiguous piece per
// row.  Gathering from shared memory instead of L2 (30 GB of 32-byte sectors per 10 M rows) leaves the kernel with
// its real work: one code byte in, sub fp16 values out, a pure stream with PQD_UNROLL independent tasks per thread in
// flight and no barrier in the loop.
constexpr int PQD_THREADS = 1024;
constexpr int PQD_UNROLL = 4;
constexpr int PQD_TABLE_BYTES = 192 * 1024;

template <int SUB>
__global__ void __launch_bounds__(PQD_THREADS, 1)
pq_decode_kernel(const uint8_t* __restrict__ codes, const __half* __restrict__ cb16, int M, int mg, uint32_t r0, uint32_t n,
                 __half* __restrict__ x) {
    extern __shared__ __align__(16) uint8_t pqd_smem[];
    __half* cbs = reinterpret_cast<__half*>(pqd_smem);
    const int g = blockIdx.y, j0 = g * mg;
    const int mgc = min(mg, M - j0);
    const int tid = threadIdx.x;
    {
        const uint4* src = reinterpret_cast<const uint4*>(cb16 + (size_t)j0 * 256 * SUB);
        uint4* dst = reinterpret_cast<uint4*>(cbs);
        const int n16 = mgc * 256 * SUB * 2 / 16;
        for (int i = tid; i < n16; i += PQD_THREADS) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    const size_t dims = (size_t)M * SUB;
    // this CTA's rows: an even split of [0, n) over gridDim.x
    const uint32_t per = (n + gridDim.x - 1) / gridDim.x;
    const uint32_t rb = blockIdx.x * per;
    const uint32_t re = min(n, rb + per);
    if (rb >= re) return;
    const uint32_t tasks = (re - rb) * (uint32_t)mgc;
    for (uint32_t t0 = tid; t0 < tasks; t0 += PQD_THREADS * PQD_UNROLL) {
        uint32_t row[PQD_UNROLL], jj[PQD_UNROLL], code[PQD_UNROLL];
#pragma unroll
        for (int u = 0; u < PQD_UNROLL; u++) {
            const uint32_t t = t0 + u * PQD_THREADS;
            const uint32_t lr = t / (uint32_t)mgc;
            row[u] = rb + lr;
            jj[u] = t - lr * (uint32_t)mgc;
            code[u] = (t < tasks) ? __ldg(codes + (size_t)(r0 + row[u]) * M + j0 + jj[u]) : 0u;
        }
#pragma unroll
        for (int u = 0; u < PQD_UNROLL; u++) {
            if (t0 + u * PQD_THREADS < tasks) {
                const __half* src = cbs + ((size_t)jj[u] * 256 + code[u]) * SUB;
                __half* dst = x + (size_t)row[u] * dims + (size_t)(j0 + jj[u]) * SUB;
                if constexpr (SUB % 8 == 0) {
#pragma unroll
                    for (int w = 0; w < SUB; w += 8)
                        __stcs(reinterpret_cast<uint4*>(dst + w), *reinterpret_cast<const uint4*>(src + w));
                } else if constexpr (SUB % 4 == 0) {
                    __stcs(reinterpret_cast<uint2*>(dst), *reinterpret_cast<const uint2*>(src));
                } else {
                    *reinterpret_cast<__half2*>(dst) = *reinterpret_cast<const __half2*>(src);
                }
            }
        }
    }
}

bool pq_gemm_eligible(int dims, int M, int sub) {
    if (M * sub != dims) return false;
    if ((dims * 2) % 16 != 0) return false;                 // TMA row pitch of the decoded slab
    return sub == 2 || sub == 4 || sub == 8 || sub == 16 || sub == 32;
}

cudaError_t launch_pq_decode(const uint8_t* codes, const void* cb16, int M, int sub, uint32_t r0, uint32_t n, void* x,
                             int sm_count, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    int mg = PQD_TABLE_BYTES / (256 * sub * 2);
    if (mg > M) mg = M;
    if (mg < 1) mg = 1;
    const int G = (M + mg - 1) / mg;
    int bx = sm_count / G;
    if (bx < 1) bx = 1;
    const uint32_t max_bx = (n + 63) / 64;
    if ((uint32_t)bx > max_bx) bx = (int)max_bx;
    const dim3 grid(bx, G);
    const size_t smem = (size_t)mg * 256 * sub * 2;
#define LB_PQD(S_)                                                                                       \
    {                                                                                                    \
        auto kern = pq_decode_kernel<S_>;                                                                \
        LB_SMEM_OPTIN(kern);                                                                             \
        kern<<<grid, PQD_THREADS, smem, st>>>(codes, (const __half*)cb16, M, mg, r0, n, (__half*)x);     \
    }
    if (sub == 2) LB_PQD(2) else if (sub == 4) LB_PQD(4) else if (sub == 8) LB_PQD(8)
    else if (sub == 16) LB_PQD(16) else if (sub == 32) LB_PQD(32) else return cudaErrorInvalidValue;
#undef LB_PQD
    count_launch();
    return cudaGetLastError();
}

// packed (key, local row) -> (key, local row + r0); invalid entries stay invalid
__global__ void pq_offset_rows_kernel(uint64_t* __restrict__ p, size_t n, uint32_t r0) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const uint64_t v = p[i];
        if (v != kInvalid) p[i] = v + r0;
    }
}

cudaError_t launch_pq_offset_rows(uint64_t* p, size_t n, uint32_t r0, cudaStream_t st) {
    if (n == 0 || r0 == 0) return cudaSuccess;
    pq_offset_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, n, r0);
    count_launch();
    return cudaGetLastError();
}

}  // namespace lb
