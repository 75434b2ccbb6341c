// kernels_misc.cu -- the small arithmetic helpers either side of the hot path (SURVEY.md 8 f1 / f2 remainder):
// SQ8 scalar quantisation (internal/simd/sq8.go:70-104), the HNSW computer's inline SQ8 de-quantising distance
// (internal/store/arrow_hnsw.go:1176-1186), FindNearestCentroid (internal/simd/simd.go:278-326) and the
// row -> VectorID scatter of GenerateFilterBitset (internal/store/dataset.go:247-300).
#include <algorithm>
#include <cfloat>

#pragma GCC visibility push(default)
#include "../../include/longbow_b200.h"
#pragma GCC visibility pop
#include "kernels.cuh"

namespace lb {
int api_fail(int code, const char* what);
int api_fail_cuda(cudaError_t e, const char* where);
int api_use_device(int device);

// QuantizeSQ8: scale = 255 / (max - min) (0 if equal), val = (v - min) * scale clamped to [0, 255], byte(val)
// truncates.  Two separate fp32 roundings, as the Go compiler emits on amd64.
__global__ void quantize_sq8_kernel(const float* __restrict__ src, int64_t n, float minv, float maxv,
                                    uint8_t* __restrict__ dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float scale = (maxv == minv) ? 0.f : __fdiv_rn(255.0f, __fsub_rn(maxv, minv));
    float val = __fmul_rn(__fsub_rn(src[i], minv), scale);
    if (val < 0.f) val = 0.f;
    if (val > 255.f) val = 255.f;
    dst[i] = (val == val) ? (uint8_t)val : (uint8_t)0;
}

// arrow_hnsw.go:1177-1181: deq = min + float32(code) * scale with scale = (max - min) / 255
__global__ void dequantize_sq8_kernel(const uint8_t* __restrict__ src, int64_t n, float minv, float maxv,
                                      float* __restrict__ dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float scale = __fdiv_rn(__fsub_rn(maxv, minv), 255.0f);
    dst[i] = __fadd_rn(minv, __fmul_rn((float)src[i], scale));
}

// ComputeBounds (sq8.go:88-104): min / max of a vector; NaN never wins a '<' or '>' comparison, and the first
// element seeds both, exactly as the sequential loop does -- except that a NaN first element would stick there
// and not here (documented: NaN inputs are outside the contract).
__global__ void bounds_kernel(const float* __restrict__ v, int64_t n, float* __restrict__ out /*[2]: as ordered uints*/) {
    float mn = FLT_MAX, mx = -FLT_MAX;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float x = v[i];
        mn = fminf(mn, x);
        mx = fmaxf(mx, x);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(reinterpret_cast<unsigned int*>(out), float_to_ordered(mn));
        atomicMax(reinterpret_cast<unsigned int*>(out) + 1, float_to_ordered(mx));
    }
}

// One fp32 query against n SQ8 rows de-quantised inline (arrow_hnsw.go:1176-1186): single sequential fp32
// accumulator, diff = q - (min + code * scale), sqrt through double.
__global__ void sq8_dequant_distance_kernel(const float* __restrict__ q, const uint8_t* __restrict__ rows, int64_t n,
                                            int dim, float minv, float maxv, float* __restrict__ out) {
    extern __shared__ float qs[];
    for (int i = threadIdx.x; i < dim; i += blockDim.x) qs[i] = q[i];
    __syncthreads();
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const float scale = __fdiv_rn(__fsub_rn(maxv, minv), 255.0f);
    const uint8_t* row = rows + (size_t)r * dim;
    float sum = 0.f;
    for (int i = 0; i < dim; i++) {
        const float deq = __fadd_rn(minv, __fmul_rn((float)row[i], scale));
        const float diff = __fsub_rn(qs[i], deq);
        sum = __fadd_rn(sum, __fmul_rn(diff, diff));
    }
    out[r] = __fsqrt_rn(sum);
}

// FindNearestCentroid (simd.go:278-326): k <= 8 compares SQUARED distances starting from MaxFloat32, larger k
// compares the sqrt'd EuclideanDistanceBatchFlat results starting from results[0]; strict '<' keeps the first
// minimum.  One block; thread c computes centroid c (looping when k > blockDim), thread 0 scans in index order.
__global__ void find_nearest_centroid_kernel(const float* __restrict__ query, const float* __restrict__ cent, int sub,
                                             int k, int* __restrict__ out_idx, float* __restrict__ out_d) {
    extern __shared__ float sm[];  // [sub] query, then [k] distances
    float* d = sm + sub;
    for (int i = threadIdx.x; i < sub; i += blockDim.x) sm[i] = query[i];
    __syncthreads();
    for (int c = threadIdx.x; c < k; c += blockDim.x) {
        ExactAcc<METRIC_L2> acc;
        acc.init();
        const float* cc = cent + (size_t)c * sub;
        int i = 0;
        for (; i <= sub - 4; i += 4) {
#pragma unroll
            for (int e = 0; e < 4; e++) acc.add(e, sm[i + e], cc[i + e]);
        }
        for (; i < sub; i++) acc.add(0, sm[i], cc[i]);
        d[c] = (k <= 8) ? acc.sum() : acc.finish();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float best = (k <= 8) ? FLT_MAX : d[0];
        int bi = 0;
        for (int c = (k <= 8 ? 0 : 1); c < k; c++)
            if (d[c] < best) { best = d[c]; bi = c; }
        *out_idx = bi;
        *out_d = best;
    }
}

// GenerateFilterBitset's scatter (dataset.go:268-277): every matching row i of a record batch sets the bit of its
// VectorID -- ids[i] when the index has one for the row (0xffffffff = GetVectorID miss), or vid_base + i for a
// batch whose rows were indexed contiguously.
__global__ void filter_scatter_kernel(const uint32_t* __restrict__ batch_bitmap, int64_t n_rows,
                                      const uint32_t* __restrict__ ids, uint32_t vid_base, uint32_t n_vids,
                                      uint32_t* __restrict__ global_bitmap) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    if (!((batch_bitmap[i >> 5] >> (i & 31)) & 1u)) return;
    const uint32_t vid = ids ? ids[i] : vid_base + (uint32_t)i;
    if (vid >= n_vids) return;
    atomicOr(global_bitmap + (vid >> 5), 1u << (vid & 31));
}

}  // namespace lb
using namespace lb;

#define MCK(call)                                                  \
    do {                                                           \
        cudaError_t e__ = (call);                                  \
        if (e__ != cudaSuccess) return api_fail_cuda(e__, #call);  \
    } while (0)

struct MScratch {
    cudaStream_t st;
    void* p[8];
    int n = 0;
    explicit MScratch(cudaStream_t s) : st(s) {}
    cudaError_t get(void** out, size_t bytes) {
        cudaError_t e = cudaMallocAsync(out, bytes ? bytes : 16, st);
        if (e == cudaSuccess && n < 8) p[n++] = *out;
        return e;
    }
    ~MScratch() { for (int i = 0; i < n; i++) cudaFreeAsync(p[i], st); }
};

extern "C" {

int lb_simd_quantize_sq8(int device, const float* src, int64_t n, float min_val, float max_val, uint8_t* dst) {
    if (n < 0) return api_fail(LB_ERR_INVALID, "n < 0");
    if (n == 0) return LB_OK;
    if (!src || !dst) return api_fail(LB_ERR_INVALID, "NULL buffer");
    int rc = api_use_device(device);
    if (rc) return rc;
    cudaStream_t st = cudaStreamPerThread;
    MScratch scr(st);
    float* d_s; uint8_t* d_d;
    MCK(scr.get((void**)&d_s, (size_t)n * 4));
    MCK(scr.get((void**)&d_d, (size_t)n));
    MCK(cudaMemcpyAsync(d_s, src, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    quantize_sq8_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_s, n, min_val, max_val, d_d);
    count_launch();
    MCK(cudaGetLastError());
    MCK(cudaMemcpyAsync(dst, d_d, (size_t)n, cudaMemcpyDeviceToHost, st));
    MCK(cudaStreamSynchronize(st));
    return LB_OK;
}

int lb_simd_dequantize_sq8(int device, const uint8_t* src, int64_t n, float min_val, float max_val, float* dst) {
    if (n < 0) return api_fail(LB_ERR_INVALID, "n < 0");
    if (n == 0) return LB_OK;
    if (!src || !dst) return api_fail(LB_ERR_INVALID, "NULL buffer");
    int rc = api_use_device(device);
    if (rc) return rc;
    cudaStream_t st = cudaStreamPerThread;
    MScratch scr(st);
    uint8_t* d_s; float* d_d;
    MCK(scr.get((void**)&d_s, (size_t)n));
    MCK(scr.get((void**)&d_d, (size_t)n * 4));
    MCK(cudaMemcpyAsync(d_s, src, (size_t)n, cudaMemcpyHostToDevice, st));
    dequantize_sq8_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_s, n, min_val, max_val, d_d);
    count_launch();
    MCK(cudaGetLastError());
    MCK(cudaMemcpyAsync(dst, d_d, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    MCK(cudaStreamSynchronize(st));
    return LB_OK;
}

int lb_simd_compute_bounds(int device, const float* vec, int64_t n, float* min_val, float* max_val) {
    if (n < 0 || !min_val || !max_val) return api_fail(LB_ERR_INVALID, "bad argument");
    if (n == 0) { *min_val = 0.f; *max_val = 0.f; return LB_OK; }  // sq8.go:89-91
    if (!vec) return api_fail(LB_ERR_INVALID, "NULL buffer");
    int rc = api_use_device(device);
    if (rc) return rc;
    cudaStream_t st = cudaStreamPerThread;
    MScratch scr(st);
    float* d_v; uint32_t* d_o;
    MCK(scr.get((void**)&d_v, (size_t)n * 4));
    MCK(scr.get((void**)&d_o, 8));
    MCK(cudaMemcpyAsync(d_v, vec, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    const uint32_t init[2] = {0xffffffffu, 0u};
    MCK(cudaMemcpyAsync(d_o, init, 8, cudaMemcpyHostToDevice, st));
    int blocks = (int)std::min<int64_t>((n + 255) / 256, 1184);
    bounds_kernel<<<blocks, 256, 0, st>>>(d_v, n, reinterpret_cast<float*>(d_o));
    count_launch();
    MCK(cudaGetLastError());
    uint32_t o[2];
    MCK(cudaMemcpyAsync(o, d_o, 8, cudaMemcpyDeviceToHost, st));
    MCK(cudaStreamSynchronize(st));
    *min_val = ordered_to_float(o[0]);
    *max_val = ordered_to_float(o[1]);
    return LB_OK;
}

int lb_simd_sq8_dequant_distance_batch(int device, const float* query, const uint8_t* rows, int64_t n, int dim,
                                       float min_val, float max_val, float* results) {
    if (n < 0 || dim <= 0) return api_fail(LB_ERR_INVALID, "bad size");
    if (n == 0) return LB_OK;
    if (!query || !rows || !results) return api_fail(LB_ERR_INVALID, "NULL buffer");
    if (dim > 8192) return api_fail(LB_ERR_UNSUPPORTED, "dim > 8192");
    int rc = api_use_device(device);
    if (rc) return rc;
    cudaStream_t st = cudaStreamPerThread;
    MScratch scr(st);
    float *d_q, *d_o; uint8_t* d_r;
    MCK(scr.get((void**)&d_q, (size_t)dim * 4));
    MCK(scr.get((void**)&d_r, (size_t)n * dim));
    MCK(scr.get((void**)&d_o, (size_t)n * 4));
    MCK(cudaMemcpyAsync(d_q, query, (size_t)dim * 4, cudaMemcpyHostToDevice, st));
    MCK(cudaMemcpyAsync(d_r, rows, (size_t)n * dim, cudaMemcpyHostToDevice, st));
    sq8_dequant_distance_kernel<<<(unsigned)((n + 127) / 128), 128, (size_t)dim * 4, st>>>(d_q, d_r, n, dim, min_val,
                                                                                           max_val, d_o);
    count_launch();
    MCK(cudaGetLastError());
    MCK(cudaMemcpyAsync(results, d_o, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    MCK(cudaStreamSynchronize(st));
    return LB_OK;
}

int lb_simd_find_nearest_centroid(int device, const float* query, const float* centroids, int sub_dim, int k,
                                  int* out_index, float* out_distance) {
    if (!query || !centroids || !out_index || !out_distance || sub_dim <= 0 || k <= 0)
        return api_fail(LB_ERR_INVALID, "bad argument");
    if ((size_t)(sub_dim + k) * 4 > 200 * 1024) return api_fail(LB_ERR_UNSUPPORTED, "sub_dim + k too large");
    int rc = api_use_device(device);
    if (rc) return rc;
    cudaStream_t st = cudaStreamPerThread;
    MScratch scr(st);
    float *d_q, *d_c, *d_d; int* d_i;
    MCK(scr.get((void**)&d_q, (size_t)sub_dim * 4));
    MCK(scr.get((void**)&d_c, (size_t)k * sub_dim * 4));
    MCK(scr.get((void**)&d_d, 4));
    MCK(scr.get((void**)&d_i, 4));
    MCK(cudaMemcpyAsync(d_q, query, (size_t)sub_dim * 4, cudaMemcpyHostToDevice, st));
    MCK(cudaMemcpyAsync(d_c, centroids, (size_t)k * sub_dim * 4, cudaMemcpyHostToDevice, st));
    LB_SMEM_OPTIN(find_nearest_centroid_kernel);
    find_nearest_centroid_kernel<<<1, 256, (size_t)(sub_dim + k) * 4, st>>>(d_q, d_c, sub_dim, k, d_i, d_d);
    count_launch();
    MCK(cudaGetLastError());
    MCK(cudaMemcpyAsync(out_index, d_i, 4, cudaMemcpyDeviceToHost, st));
    MCK(cudaMemcpyAsync(out_distance, d_d, 4, cudaMemcpyDeviceToHost, st));
    MCK(cudaStreamSynchronize(st));
    return LB_OK;
}

int lb_filter_scatter_device(int device, const uint64_t* d_batch_bitmap, int64_t n_rows, const uint32_t* d_vector_ids,
                             uint32_t vid_base, int64_t n_vector_ids, uint64_t* d_global_bitmap, void* stream) {
    if (n_rows < 0 || n_vector_ids < 0) return api_fail(LB_ERR_INVALID, "bad size");
    if (n_rows == 0) return LB_OK;
    if (!d_batch_bitmap || !d_global_bitmap) return api_fail(LB_ERR_INVALID, "NULL buffer");
    int rc = api_use_device(device);
    if (rc) return rc;
    const uint32_t nv = (uint32_t)std::min<int64_t>(n_vector_ids, 0xffffffffll);
    filter_scatter_kernel<<<(unsigned)((n_rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint32_t*>(d_batch_bitmap), n_rows, d_vector_ids, vid_base, nv,
        reinterpret_cast<uint32_t*>(d_global_bitmap));
    count_launch();
    MCK(cudaGetLastError());
    return LB_OK;
}

}  // extern "C"
